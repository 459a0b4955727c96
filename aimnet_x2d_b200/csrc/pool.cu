// a7 / a8 -- graph pooling (models/pooling.py).
//
// Attention pooling is ONE kernel per direction: a CTA owns a molecule, stages its rows of x in shared
// memory with a bulk async copy, and does score -> per-(head,molecule) softmax -> weighted sum from
// there, so x is read from HBM exactly once and the reference's [heads, N, F] temporary
// (pooling.py:150-154) never exists.  Molecules with more rows than the shared-memory chunk are
// processed in chunks (second pass re-reads x, served by L2).
// Parameter gradients are reduced in two passes with a fixed order (no atomics).
#include "common.cuh"

namespace ax2d {

constexpr int kPoolThreads = 128;
constexpr int kMaxHeads = 8;
// per-CTA partial record of the backward: [heads*F] gw, [heads] gb, [1] gT, padded to whole float4s
__host__ __device__ constexpr int pool_partial_stride(int heads, int F) { return (heads * F + heads + 1 + 3) & ~3; }

struct ChunkLoader {
  uint64_t* bar;
  uint32_t phase;
  __device__ void init(uint64_t* b) {
    bar = b;
    phase = 0;
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
  }
  // all threads call; thread 0 issues.  Caller must have passed a __syncthreads() since the last reads of dst.
  __device__ void load(float* dst, const float* src, uint32_t bytes) {
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, bytes);
      const uint32_t chunk = 32768;
      for (uint32_t off = 0; off < bytes; off += chunk) {
        const uint32_t n = bytes - off < chunk ? bytes - off : chunk;
        bulk_g2s(reinterpret_cast<unsigned char*>(dst) + off, reinterpret_cast<const unsigned char*>(src) + off, n,
                 bar);
      }
    }
  }
  __device__ void wait() {
    mbar_wait(bar, phase);
    phase ^= 1;
  }
};

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kPoolThreads) attn_pool_fwd_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ seg_ptr, int64_t N, int F, int heads,
    const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ temperature,
    float* __restrict__ pooled, float* __restrict__ attn, float* __restrict__ zbuf, int CH) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  float* xs = reinterpret_cast<float*>(smem_raw);                 // [CH, F]
  float* ws = xs + static_cast<size_t>(CH) * F;                   // [heads, F]
  float* as = ws + static_cast<size_t>(heads) * F;                // [heads, CH]  scores then weights
  float* abar = as + static_cast<size_t>(heads) * CH;             // [CH]

  const int g = blockIdx.x;
  const int n0 = seg_ptr[g], n1 = seg_ptr[g + 1];
  const int n = n1 - n0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F4 = F >> 2;
  const bool single = n <= CH;
  float4* pooled4 = reinterpret_cast<float4*>(pooled + static_cast<int64_t>(g) * F);
  if (n == 0) {   // torch_scatter leaves untouched segments at 0
    for (int c = tid; c < F4; c += kPoolThreads) pooled4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  ChunkLoader ld;
  ld.init(&bar);
  ld.load(xs, x + static_cast<int64_t>(n0) * F, static_cast<uint32_t>((single ? n : CH) * F * 4));
  for (int i = tid; i < heads * F4; i += kPoolThreads)
    reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  const float T = __ldg(temperature);
  __syncthreads();

  // ---- pass 1: scores z[h,i] = (w_h . x_i + b_h) / T            (pooling.py:134-140)
  for (int c0 = 0; c0 < n; c0 += CH) {
    const int rows = n - c0 < CH ? n - c0 : CH;
    if (c0 > 0) {
      __syncthreads();
      ld.load(xs, x + static_cast<int64_t>(n0 + c0) * F, static_cast<uint32_t>(rows * F * 4));
    }
    ld.wait();
    for (int i = warp; i < rows; i += kPoolThreads / 32) {
      float dot[kMaxHeads];
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h) dot[h] = 0.f;
      const float4* xr = reinterpret_cast<const float4*>(xs + static_cast<size_t>(i) * F);
      for (int c = lane; c < F4; c += 32) {
        const float4 xv = xr[c];
#pragma unroll
        for (int h = 0; h < kMaxHeads; ++h)
          if (h < heads) {
            const float4 wv = reinterpret_cast<const float4*>(ws + static_cast<size_t>(h) * F)[c];
            dot[h] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
          }
      }
#pragma unroll
      for (int h = 0; h < kMaxHeads; ++h)
        if (h < heads) {
          const float s = warp_sum(dot[h]);
          if (lane == 0) {
            const float z = (s + __ldg(b + h)) / T;
            zbuf[static_cast<int64_t>(h) * N + n0 + c0 + i] = z;
            if (single) as[h * CH + i] = z;
          }
        }
    }
  }
  __syncthreads();

  // ---- softmax per (head, molecule)                              (pooling.py:144-145, scatter_softmax)
  for (int h = warp; h < heads; h += kPoolThreads / 32) {
    const float* zrow = single ? (as + h * CH) : (zbuf + static_cast<int64_t>(h) * N + n0);
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, zrow[i]);
    m = warp_max(m);
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += expf(zrow[i] - m);
    s = warp_sum(s);
    float* arow = attn + static_cast<int64_t>(h) * N + n0;
    for (int i = lane; i < n; i += 32) {
      const float a = expf(zrow[i] - m) / s;
      arow[i] = a;
      if (single) as[h * CH + i] = a;
    }
  }
  __syncthreads();

  // ---- pass 2: pooled = mean_h sum_i a[h,i] x_i                  (pooling.py:150-161)
  float4 acc[4];   // F <= 4 * 128 * 4 = 2048 columns
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c0 = 0; c0 < n; c0 += CH) {
    const int rows = n - c0 < CH ? n - c0 : CH;
    if (!single) {
      __syncthreads();
      ld.load(xs, x + static_cast<int64_t>(n0 + c0) * F, static_cast<uint32_t>(rows * F * 4));
      ld.wait();
    }
    for (int i = tid; i < rows; i += kPoolThreads) {
      float sacc = 0.f;
      for (int h = 0; h < heads; ++h)
        sacc += single ? as[h * CH + i] : attn[static_cast<int64_t>(h) * N + n0 + c0 + i];
      abar[i] = sacc;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = tid + j * kPoolThreads;
      if (c < F4) {
        for (int i = 0; i < rows; ++i) {
          const float a = abar[i];
          const float4 xv = reinterpret_cast<const float4*>(xs + static_cast<size_t>(i) * F)[c];
          acc[j].x += a * xv.x;
          acc[j].y += a * xv.y;
          acc[j].z += a * xv.z;
          acc[j].w += a * xv.w;
        }
      }
    }
  }
  const float inv_h = 1.f / static_cast<float>(heads);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = tid + j * kPoolThreads;
    if (c < F4) pooled4[c] = make_float4(acc[j].x * inv_h, acc[j].y * inv_h, acc[j].z * inv_h, acc[j].w * inv_h);
  }
}

// Forward for molecules of at most kDirectRows atoms (every drug-like batch): the same arithmetic in the same order
// as attn_pool_fwd_kernel, but x is NOT staged in shared memory -- pass 1 reads each row from global memory (one
// coalesced warp-wide read per row), pass 2 reads the molecule's rows again, which L2 still holds.  With ~10 KB of
// shared memory instead of ~68 KB per CTA an SM holds 12-16 molecules at a time instead of 3, which is what this
// latency-bound kernel (a few dependent phases over ~36 KB per molecule) needs.  The head weights (8 KB, the same for
// every CTA) are read through L1 instead of being copied to shared memory first.
constexpr int kDirectRows = 256;
// MAXH: compile-time bound of the head loops (the number of heads rounded up to 1 / 2 / 4 / 8): with the generic bound of 8
// half of the issued instructions of the default 4-head layer were predicated-off head iterations.
template <int MAXH>
__global__ void __launch_bounds__(kPoolThreads) attn_pool_fwd_direct_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ seg_ptr, int64_t N, int F, int heads,
    const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ temperature,
    float* __restrict__ pooled, float* __restrict__ attn, float* __restrict__ zbuf) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* as = reinterpret_cast<float*>(smem_raw);                 // [heads, kDirectRows]  scores then weights
  float* abar = as + static_cast<size_t>(heads) * kDirectRows;    // [kDirectRows]

  const int g = blockIdx.x;
  const int n0 = seg_ptr[g], n1 = seg_ptr[g + 1];
  const int n = n1 - n0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F4 = F >> 2;
  const bool fits = n <= kDirectRows;      // a molecule beyond the hint: scores / weights go through global memory
  float4* pooled4 = reinterpret_cast<float4*>(pooled + static_cast<int64_t>(g) * F);
  if (n == 0) {   // torch_scatter leaves untouched segments at 0
    for (int c = tid; c < F4; c += kPoolThreads) pooled4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float T = __ldg(temperature);
  const float4* w4 = reinterpret_cast<const float4*>(w);          // 8 KB shared by every CTA: served by L1

  // ---- pass 1: scores z[h,i] = (w_h . x_i + b_h) / T            (pooling.py:134-140)
  const float4* xg = reinterpret_cast<const float4*>(x + static_cast<int64_t>(n0) * F);
  for (int i = warp; i < n; i += kPoolThreads / 32) {
    float dot[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) dot[h] = 0.f;
    const float4* xr = xg + static_cast<size_t>(i) * F4;
    // (unrolling this loop so that the row's four loads issue together costs 90 registers and occupancy: 75 vs 60 us)
    for (int c = lane; c < F4; c += 32) {
      const float4 xv = __ldg(xr + c);
#pragma unroll
      for (int h = 0; h < MAXH; ++h)
        if (h < heads) {
          const float4 wv = __ldg(w4 + h * F4 + c);
          dot[h] += xv.x * wv.x + xv.y * wv.y + xv.z * wv.z + xv.w * wv.w;
        }
    }
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
      if (h < heads) {
        const float s = warp_sum(dot[h]);
        if (lane == 0) {
          const float z = (s + __ldg(b + h)) / T;
          zbuf[static_cast<int64_t>(h) * N + n0 + i] = z;
          if (fits) as[h * kDirectRows + i] = z;
        }
      }
  }
  __syncthreads();

  // ---- softmax per (head, molecule)                              (pooling.py:144-145, scatter_softmax)
  for (int h = warp; h < heads; h += kPoolThreads / 32) {
    float* zrow = fits ? as + h * kDirectRows : zbuf + static_cast<int64_t>(h) * N + n0;
    float m = -INFINITY;
    for (int i = lane; i < n; i += 32) m = fmaxf(m, zrow[i]);
    m = warp_max(m);
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += expf(zrow[i] - m);
    s = warp_sum(s);
    float* arow = attn + static_cast<int64_t>(h) * N + n0;
    for (int i = lane; i < n; i += 32) {
      const float a = expf(zrow[i] - m) / s;
      arow[i] = a;
      if (fits) zrow[i] = a;
    }
  }
  __syncthreads();

  // ---- pass 2: pooled = mean_h sum_i a[h,i] x_i                  (pooling.py:150-161)
  if (fits) {
    for (int i = tid; i < n; i += kPoolThreads) {
      float sacc = 0.f;
      for (int h = 0; h < heads; ++h) sacc += as[h * kDirectRows + i];
      abar[i] = sacc;
    }
    __syncthreads();
  }
  const float inv_h = 1.f / static_cast<float>(heads);
  for (int c = tid; c < F4; c += kPoolThreads) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
      float a;
      if (fits) {
        a = abar[i];
      } else {
        a = 0.f;
        for (int h = 0; h < heads; ++h) a += attn[static_cast<int64_t>(h) * N + n0 + i];
      }
      const float4 xv = __ldg(xg + static_cast<size_t>(i) * F4 + c);
      acc.x += a * xv.x;
      acc.y += a * xv.y;
      acc.z += a * xv.z;
      acc.w += a * xv.w;
    }
    pooled4[c] = make_float4(acc.x * inv_h, acc.y * inv_h, acc.z * inv_h, acc.w * inv_h);
  }
}

// ------------------------------------------------------------------------------------------ backward
// Formulas: SURVEY.md appendix C6.  d_i = (1/H) G_g . x_i ; da[h,i] = d_i (+ Ga[h,i]) ;
// dz[h,i] = a[h,i] (da[h,i] - sum_j a[h,j] da[h,j]) ; gx_i = (sum_h a[h,i]/H) G_g + (1/T) sum_h dz[h,i] w_h ;
// gw_h = (1/T) sum_i dz[h,i] x_i ; gb_h = (1/T) sum_i dz[h,i] ; gT = -(1/T) sum dz z.
template <int NH>
__global__ void __launch_bounds__(kPoolThreads) attn_pool_bwd_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ seg_ptr, int64_t B, int64_t N, int F,
    const float* __restrict__ w, const float* __restrict__ temperature, const float* __restrict__ attn,
    const float* __restrict__ zbuf, const float* __restrict__ g_pooled, const float* __restrict__ g_attn,
    float* __restrict__ gx, int64_t ldgx, float* __restrict__ partials, int CH) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ float red[kPoolThreads / 32][NH + 1];
  float* xs = reinterpret_cast<float*>(smem_raw);                 // [CH, F]
  float* gs = xs + static_cast<size_t>(CH) * F;                   // [F]
  float* dzs = gs + F;                                            // [NH, CH]
  float* coef = dzs + static_cast<size_t>(NH) * CH;               // [CH]  sum_h a / H
  float* ds = coef + CH;                                          // [CH]
  float* sdot = ds + CH;                                          // [NH]
  float* zmax = sdot + NH;                                        // [NH]  per-(head, molecule) max score
  float* resid = zmax + NH;                                       // [NH]  sum_i dz[h,i] (rounding residual)
  int* amax = reinterpret_cast<int*>(resid + NH);                 // [NH]  first arg-max atom (molecule-local)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F4 = F >> 2;
  const float T = __ldg(temperature);
  const float invT = 1.f / T;
  const float inv_h = 1.f / static_cast<float>(NH);

  // per-thread parameter-gradient accumulators for its float4 columns (tid, tid+128, ...)
  float4 gw_acc[NH][4];
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int j = 0; j < 4; ++j) gw_acc[h][j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float gb_acc[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) gb_acc[h] = 0.f;
  float gT_acc = 0.f;

  ChunkLoader ld;
  ld.init(&bar);

  for (int64_t g = blockIdx.x; g < B; g += gridDim.x) {
    const int n0 = seg_ptr[g], n1 = seg_ptr[g + 1];
    const int n = n1 - n0;
    if (n == 0) continue;
    const bool single = n <= CH;
    __syncthreads();
    ld.load(xs, x + static_cast<int64_t>(n0) * F, static_cast<uint32_t>((single ? n : CH) * F * 4));
    for (int c = tid; c < F4; c += kPoolThreads)
      reinterpret_cast<float4*>(gs)[c] = __ldg(reinterpret_cast<const float4*>(g_pooled + g * F) + c);
    __syncthreads();
    // ---- pass A: d_i for every row of the molecule
    for (int c0 = 0; c0 < n; c0 += CH) {
      const int rows = n - c0 < CH ? n - c0 : CH;
      if (c0 > 0) {
        __syncthreads();
        ld.load(xs, x + static_cast<int64_t>(n0 + c0) * F, static_cast<uint32_t>(rows * F * 4));
      }
      ld.wait();
      for (int i = warp; i < rows; i += kPoolThreads / 32) {
        const float4* xr = reinterpret_cast<const float4*>(xs + static_cast<size_t>(i) * F);
        float d = 0.f;
        for (int c = lane; c < F4; c += 32) {
          const float4 xv = xr[c];
          const float4 gv = reinterpret_cast<const float4*>(gs)[c];
          d += xv.x * gv.x + xv.y * gv.y + xv.z * gv.z + xv.w * gv.w;
        }
        d = warp_sum(d) * inv_h;
        if (lane == 0) {
          if (single) ds[i] = d;
          else gx[static_cast<int64_t>(n0 + c0 + i) * ldgx] = d;   // parked in column 0 until pass B
        }
      }
    }
    __syncthreads();
    // ---- per head: sdot_h = sum_i a[h,i] da[h,i], the score maximum / its first position, and the residual
    //      r_h = sum_i dz[h,i].  In exact arithmetic r_h == 0; the reference's scatter_softmax is the composite
    //      exp(z - max) / sum, whose autograd routes -sum_i dz[h,i] to the arg-max atom (gradient of the
    //      gathered maximum, torch_scatter softmax.py), i.e. the reference's dz sums to zero up to ONE rounding.
    //      The same correction is applied here: it is what keeps the shift-sensitive sums (gw, gb, gT) at the
    //      reference's accuracy when scores / features carry a large common offset.
    for (int h = warp; h < NH; h += kPoolThreads / 32) {
      float s = 0.f, zm = -INFINITY;
      int am = 0x7fffffff;
      for (int i = lane; i < n; i += 32) {
        float da = single ? ds[i] : gx[static_cast<int64_t>(n0 + i) * ldgx];
        if (g_attn != nullptr) da += g_attn[static_cast<int64_t>(h) * N + n0 + i];
        s += attn[static_cast<int64_t>(h) * N + n0 + i] * da;
        const float zv = zbuf[static_cast<int64_t>(h) * N + n0 + i];
        if (zv > zm) { zm = zv; am = i; }
      }
      s = warp_sum(s);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float oz = __shfl_xor_sync(0xffffffffu, zm, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (oz > zm || (oz == zm && oa < am)) { zm = oz; am = oa; }
      }
      float r = 0.f;
      for (int i = lane; i < n; i += 32) {
        float da = single ? ds[i] : gx[static_cast<int64_t>(n0 + i) * ldgx];
        if (g_attn != nullptr) da += g_attn[static_cast<int64_t>(h) * N + n0 + i];
        r += attn[static_cast<int64_t>(h) * N + n0 + i] * (da - s);
      }
      r = warp_sum(r);
      if (lane == 0) {
        sdot[h] = s;
        zmax[h] = zm;
        resid[h] = r;
        amax[h] = am;
      }
    }
    __syncthreads();
    // ---- pass B: dz, gx, parameter-gradient partials
    for (int c0 = 0; c0 < n; c0 += CH) {
      const int rows = n - c0 < CH ? n - c0 : CH;
      if (!single) {
        __syncthreads();
        ld.load(xs, x + static_cast<int64_t>(n0 + c0) * F, static_cast<uint32_t>(rows * F * 4));
        ld.wait();
      }
      float gb_loc[NH];
#pragma unroll
      for (int h = 0; h < NH; ++h) gb_loc[h] = 0.f;
      float gT_loc = 0.f;
      for (int i = tid; i < rows; i += kPoolThreads) {
        const float d = single ? ds[i] : gx[static_cast<int64_t>(n0 + c0 + i) * ldgx];
        float asum = 0.f;
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const int64_t o = static_cast<int64_t>(h) * N + n0 + c0 + i;
          const float a = attn[o];
          float da = d;
          if (g_attn != nullptr) da += g_attn[o];
          float dz = a * (da - sdot[h]);
          if (c0 + i == amax[h]) dz -= resid[h];
          dzs[h * CH + i] = dz;
          asum += a;
          gb_loc[h] += dz;
          // sum_i dz[h,i] == 0 per (head, molecule), so z may be shifted by its maximum: the terms shrink from
          // |dz| * |z| to |dz| * (max - z) and the rounding noise of the cancelling sum is no longer multiplied
          // by the common offset of the scores
          gT_loc += dz * (zbuf[o] - zmax[h]);
        }
        coef[i] = asum * inv_h;
      }
      // block-reduce the scalar partials in a fixed order
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float s = warp_sum(gb_loc[h]);
        if (lane == 0) red[warp][h] = s;
      }
      {
        const float s = warp_sum(gT_loc);
        if (lane == 0) red[warp][NH] = s;
      }
      __syncthreads();
      if (tid == 0) {
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          float s = 0.f;
          for (int wq = 0; wq < kPoolThreads / 32; ++wq) s += red[wq][h];
          gb_acc[h] += s;
        }
        float s = 0.f;
        for (int wq = 0; wq < kPoolThreads / 32; ++wq) s += red[wq][NH];
        gT_acc += s;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = tid + j * kPoolThreads;
        if (c < F4) {
          const float4 gv = reinterpret_cast<const float4*>(gs)[c];
          float4 wv[NH];
#pragma unroll
          for (int h = 0; h < NH; ++h) wv[h] = __ldg(reinterpret_cast<const float4*>(w + static_cast<size_t>(h) * F) + c);
          // two-level sums: dz sums to zero within a molecule, so sum_i dz[h,i] x_i cancels; the rows of this chunk go
          // into a local accumulator that is added to the CTA's running sum once (a few molecule-sized terms per CTA
          // instead of one add per atom into an ever larger partial sum)
          float4 loc[NH];
#pragma unroll
          for (int h = 0; h < NH; ++h) loc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int i = 0; i < rows; ++i) {
            const float4 xv = reinterpret_cast<const float4*>(xs + static_cast<size_t>(i) * F)[c];
            const float cf = coef[i];
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int h = 0; h < NH; ++h) {
              const float dz = dzs[h * CH + i];
              t.x += dz * wv[h].x; t.y += dz * wv[h].y; t.z += dz * wv[h].z; t.w += dz * wv[h].w;
              loc[h].x += dz * xv.x; loc[h].y += dz * xv.y;
              loc[h].z += dz * xv.z; loc[h].w += dz * xv.w;
            }
            float4 o;
            o.x = cf * gv.x + invT * t.x; o.y = cf * gv.y + invT * t.y;
            o.z = cf * gv.z + invT * t.z; o.w = cf * gv.w + invT * t.w;
            reinterpret_cast<float4*>(gx + static_cast<int64_t>(n0 + c0 + i) * ldgx)[c] = o;
          }
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            gw_acc[h][j].x += loc[h].x; gw_acc[h][j].y += loc[h].y;
            gw_acc[h][j].z += loc[h].z; gw_acc[h][j].w += loc[h].w;
          }
        }
      }
    }
  }
  // ---- per-CTA partials: [NH*F] gw, [NH] gb, [1] gT
  float* my = partials + static_cast<size_t>(blockIdx.x) * pool_partial_stride(NH, F);
#pragma unroll
  for (int h = 0; h < NH; ++h)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = tid + j * kPoolThreads;
      if (c < F4) reinterpret_cast<float4*>(my + static_cast<size_t>(h) * F)[c] = gw_acc[h][j];
    }
  if (tid == 0) {
#pragma unroll
    for (int h = 0; h < NH; ++h) my[static_cast<size_t>(NH) * F + h] = gb_acc[h];
    my[static_cast<size_t>(NH) * F + NH] = gT_acc;
  }
}

// fixed-order reduction of the per-CTA partial records: lane = entry, the 8 warps take partials p, p+8, ...
__global__ void __launch_bounds__(256) attn_pool_bwd_reduce_kernel(const float* __restrict__ partials, int n_part, int F,
                                                                   int heads, const float* __restrict__ temperature,
                                                                   float* __restrict__ gw, float* __restrict__ gb,
                                                                   float* __restrict__ gT) {
  pdl_enter();
  // (the per-CTA partials are summed in double: a few hundred terms per entry, and these sums cancel)
  __shared__ double red[8][32];
  const int stride = pool_partial_stride(heads, F);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int total = heads * F + heads + 1;
  double sd = 0.0;
  if (i < total) {
#pragma unroll 4
    for (int p = warp; p < n_part; p += 8) sd += static_cast<double>(__ldg(partials + static_cast<size_t>(p) * stride + i));
  }
  red[warp][lane] = sd;
  __syncthreads();
  if (warp != 0 || i >= total) return;
  sd = red[0][lane];
#pragma unroll
  for (int w = 1; w < 8; ++w) sd += red[w][lane];
  const float s = static_cast<float>(sd);
  const float invT = 1.f / __ldg(temperature);
  if (i < heads * F) gw[i] = s * invT;
  else if (i < heads * F + heads) gb[i - heads * F] = s * invT;
  else gT[0] = -s * invT;
}

static int pool_chunk_rows(int F, int heads, int max_rows_hint, bool bwd) {
  // shared-memory budget ~96 KB per CTA so that two CTAs fit on an SM
  const int budget = 96 * 1024;
  const int fixed = bwd ? (F * 4 + heads * 16 + 64) : (heads * F * 4 + 64);
  const int per_row = F * 4 + heads * 4 + 8;
  int ch = (budget - fixed) / per_row;
  if (ch < 1) ch = 1;
  if (max_rows_hint > 0 && max_rows_hint < ch) ch = max_rows_hint;
  return ch;
}

constexpr int kPoolBwdMaxGrid = kNumSMs * 3;

// ------------------------------------------------------------------------------------------ a8
__global__ void __launch_bounds__(128) seg_reduce_fwd_kernel(const float* __restrict__ x, int64_t ldx,
                                                             const int32_t* __restrict__ seg_ptr, int F4, int mode,
                                                             float* __restrict__ out, int32_t* __restrict__ arg) {
  pdl_enter();
  const int g = blockIdx.x;
  const int n0 = seg_ptr[g], n1 = seg_ptr[g + 1];
  for (int c = threadIdx.x; c < F4; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int4 am = make_int4(n0, n0, n0, n0);
    if (mode == AX2D_SEG_MAX && n1 > n0) acc = __ldg(reinterpret_cast<const float4*>(x + static_cast<int64_t>(n0) * ldx) + c);
    for (int i = (mode == AX2D_SEG_MAX ? n0 + 1 : n0); i < n1; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + static_cast<int64_t>(i) * ldx) + c);
      if (mode == AX2D_SEG_MAX) {   // strict '>' : first maximal row wins (torch_scatter reducer)
        if (v.x > acc.x) { acc.x = v.x; am.x = i; }
        if (v.y > acc.y) { acc.y = v.y; am.y = i; }
        if (v.z > acc.z) { acc.z = v.z; am.z = i; }
        if (v.w > acc.w) { acc.w = v.w; am.w = i; }
      } else {
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (mode == AX2D_SEG_MEAN) {
      const float cnt = static_cast<float>(n1 - n0 > 1 ? n1 - n0 : 1);
      acc.x /= cnt; acc.y /= cnt; acc.z /= cnt; acc.w /= cnt;
    }
    reinterpret_cast<float4*>(out + static_cast<int64_t>(g) * F4 * 4)[c] = acc;
    if (mode == AX2D_SEG_MAX && arg != nullptr) {
      if (n1 == n0) am = make_int4(-1, -1, -1, -1);
      reinterpret_cast<int4*>(arg + static_cast<int64_t>(g) * F4 * 4)[c] = am;
    }
  }
}

__global__ void __launch_bounds__(128) seg_reduce_bwd_kernel(const float* __restrict__ g_out,
                                                             const int32_t* __restrict__ seg_ptr, int F4, int mode,
                                                             const int32_t* __restrict__ arg, float* __restrict__ gx,
                                                             int64_t ldgx) {
  pdl_enter();
  const int g = blockIdx.x;
  const int n0 = seg_ptr[g], n1 = seg_ptr[g + 1];
  for (int c = threadIdx.x; c < F4; c += blockDim.x) {
    float4 gv = __ldg(reinterpret_cast<const float4*>(g_out + static_cast<int64_t>(g) * F4 * 4) + c);
    int4 am = make_int4(0, 0, 0, 0);
    if (mode == AX2D_SEG_MEAN) {
      const float cnt = static_cast<float>(n1 - n0 > 1 ? n1 - n0 : 1);
      gv.x /= cnt; gv.y /= cnt; gv.z /= cnt; gv.w /= cnt;
    } else if (mode == AX2D_SEG_MAX) {
      am = __ldg(reinterpret_cast<const int4*>(arg + static_cast<int64_t>(g) * F4 * 4) + c);
    }
    for (int i = n0; i < n1; ++i) {
      float4 o = gv;
      if (mode == AX2D_SEG_MAX) {
        o.x = am.x == i ? gv.x : 0.f; o.y = am.y == i ? gv.y : 0.f;
        o.z = am.z == i ? gv.z : 0.f; o.w = am.w == i ? gv.w : 0.f;
      }
      reinterpret_cast<float4*>(gx + static_cast<int64_t>(i) * ldgx)[c] = o;
    }
  }
}

}  // namespace ax2d

using namespace ax2d;

extern "C" int ax2d_attn_pool_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int64_t B, int64_t N, int F,
                                  int heads, const float* w, const float* b, const float* temperature, float* pooled,
                                  float* attn, float* z, int max_rows_hint, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(F > 0 && F % 4 == 0 && F <= 2048, "ax2d_attn_pool_fwd: F=%d must be a multiple of 4, <= 2048", F);
  AX2D_CHECK_ARG(ldx == F, "ax2d_attn_pool_fwd: x must be contiguous (ldx == F)");
  AX2D_CHECK_ARG(heads >= 1 && heads <= kMaxHeads, "ax2d_attn_pool_fwd: heads=%d not in [1,%d]", heads, kMaxHeads);
  AX2D_CHECK_ALIGN(x);
  AX2D_CHECK_ALIGN(w);
  AX2D_CHECK_ALIGN(pooled);
  if (B <= 0) return AX2D_OK;
  if (max_rows_hint > 0 && max_rows_hint <= kDirectRows) {       // every molecule fits the score buffer: no x staging
    const size_t smem_d = static_cast<size_t>(heads + 1) * kDirectRows * 4;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define AX2D_POOL_DIRECT(MH)                                                                                                   \
  do {                                                                                                                         \
    if (smem_d > 48 * 1024)                                                                                                    \
      cudaFuncSetAttribute(attn_pool_fwd_direct_kernel<MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_d)); \
    launch_k(attn_pool_fwd_direct_kernel<MH>, dim3(static_cast<unsigned>(B)), dim3(kPoolThreads), smem_d, st, x, seg_ptr, N, F, heads, w, b,      \
                                                                                            temperature, pooled, attn, z);     \
  } while (0)
    if (heads <= 1) AX2D_POOL_DIRECT(1);
    else if (heads <= 2) AX2D_POOL_DIRECT(2);
    else if (heads <= 4) AX2D_POOL_DIRECT(4);
    else AX2D_POOL_DIRECT(8);
#undef AX2D_POOL_DIRECT
    return launch_status("ax2d_attn_pool_fwd");
  }
  const int CH = pool_chunk_rows(F, heads, max_rows_hint, false);
  const size_t smem = (static_cast<size_t>(CH) * F + static_cast<size_t>(heads) * F + static_cast<size_t>(heads) * CH + CH) * 4;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(attn_pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  launch_k(attn_pool_fwd_kernel, dim3(static_cast<unsigned>(B)), dim3(kPoolThreads), smem, reinterpret_cast<cudaStream_t>(stream), 
      x, seg_ptr, N, F, heads, w, b, temperature, pooled, attn, z, CH);
  return launch_status("ax2d_attn_pool_fwd");
}

extern "C" int64_t ax2d_attn_pool_bwd_workspace(int64_t B, int F, int heads) {
  (void)B;
  return static_cast<int64_t>(kPoolBwdMaxGrid) * pool_partial_stride(heads, F) * 4;
}

template <int NH>
static int launch_pool_bwd(const float* x, const int32_t* seg_ptr, int64_t B, int64_t N, int F, const float* w,
                           const float* temperature, const float* attn, const float* z, const float* g_pooled,
                           const float* g_attn, float* gx, int64_t ldgx, float* gw, float* gb, float* gT, void* ws,
                           int max_rows_hint, cudaStream_t st) {
  const int CH = pool_chunk_rows(F, NH, max_rows_hint, true);
  const size_t smem = (static_cast<size_t>(CH) * F + F + static_cast<size_t>(NH) * CH + 2 * CH + 4 * NH) * 4;
  auto kern = attn_pool_bwd_kernel<NH>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const int grid = B < kPoolBwdMaxGrid ? static_cast<int>(B) : kPoolBwdMaxGrid;
  launch_k(kern, dim3(grid), dim3(kPoolThreads), smem, st, x, seg_ptr, B, N, F, w, temperature, attn, z, g_pooled, g_attn, gx, ldgx,
                                         static_cast<float*>(ws), CH);
  int rc = launch_status("ax2d_attn_pool_bwd");
  if (rc != AX2D_OK) return rc;
  const int total = NH * F + NH + 1;
  launch_k(attn_pool_bwd_reduce_kernel, dim3((total + 31) / 32), dim3(256), 0, st, static_cast<const float*>(ws), grid, F, NH,
                                                                    temperature, gw, gb, gT);
  return launch_status("ax2d_attn_pool_bwd(reduce)");
}

extern "C" int ax2d_attn_pool_bwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int64_t B, int64_t N, int F,
                                  int heads, const float* w, const float* temperature, const float* attn,
                                  const float* z, const float* g_pooled, const float* g_attn, float* gx, int64_t ldgx,
                                  float* gw, float* gb, float* gT, void* workspace, int max_rows_hint,
                                  ax2d_stream_t stream) {
  AX2D_CHECK_ARG(F > 0 && F % 4 == 0 && F <= 2048, "ax2d_attn_pool_bwd: F=%d must be a multiple of 4, <= 2048", F);
  AX2D_CHECK_ARG(ldx == F && ldgx % 4 == 0 && ldgx >= F, "ax2d_attn_pool_bwd: bad leading dimensions");
  AX2D_CHECK_ARG(heads >= 1 && heads <= kMaxHeads, "ax2d_attn_pool_bwd: heads=%d not in [1,%d]", heads, kMaxHeads);
  AX2D_CHECK_ARG(workspace != nullptr, "ax2d_attn_pool_bwd: workspace required");
  AX2D_CHECK_ALIGN(x);
  AX2D_CHECK_ALIGN(w);
  AX2D_CHECK_ALIGN(gx);
  AX2D_CHECK_ALIGN(g_pooled);
  AX2D_CHECK_ALIGN(workspace);
  if (B <= 0) return AX2D_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define AX2D_PB(NHV)                                                                                              \
  case NHV:                                                                                                       \
    return launch_pool_bwd<NHV>(x, seg_ptr, B, N, F, w, temperature, attn, z, g_pooled, g_attn, gx, ldgx, gw, gb, gT, \
                                workspace, max_rows_hint, st);
  switch (heads) {
    AX2D_PB(1) AX2D_PB(2) AX2D_PB(3) AX2D_PB(4) AX2D_PB(5) AX2D_PB(6) AX2D_PB(7) AX2D_PB(8)
  }
#undef AX2D_PB
  return AX2D_ERR_UNSUPPORTED;
}

extern "C" int ax2d_seg_reduce_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int64_t B, int F, int mode,
                                   float* out, int32_t* arg, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(F > 0 && F % 4 == 0 && ldx % 4 == 0, "ax2d_seg_reduce_fwd: F and ldx must be multiples of 4");
  AX2D_CHECK_ARG(mode >= AX2D_SEG_SUM && mode <= AX2D_SEG_MAX, "ax2d_seg_reduce_fwd: bad mode %d", mode);
  AX2D_CHECK_ALIGN(x);
  AX2D_CHECK_ALIGN(out);
  AX2D_CHECK_ALIGN(arg);
  if (B <= 0) return AX2D_OK;
  launch_k(seg_reduce_fwd_kernel, dim3(static_cast<unsigned>(B)), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, ldx, seg_ptr, F / 4, mode, out, arg);
  return launch_status("ax2d_seg_reduce_fwd");
}

extern "C" int ax2d_seg_reduce_bwd(const float* g_out, const int32_t* seg_ptr, int64_t B, int F, int mode,
                                   const int32_t* arg, float* gx, int64_t ldgx, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(F > 0 && F % 4 == 0 && ldgx % 4 == 0, "ax2d_seg_reduce_bwd: F and ldgx must be multiples of 4");
  AX2D_CHECK_ARG(mode != AX2D_SEG_MAX || arg != nullptr, "ax2d_seg_reduce_bwd: max needs arg");
  AX2D_CHECK_ALIGN(g_out);
  AX2D_CHECK_ALIGN(gx);
  if (B <= 0) return AX2D_OK;
  launch_k(seg_reduce_bwd_kernel, dim3(static_cast<unsigned>(B)), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 
      g_out, seg_ptr, F / 4, mode, arg, gx, ldgx);
  return launch_status("ax2d_seg_reduce_bwd");
}
