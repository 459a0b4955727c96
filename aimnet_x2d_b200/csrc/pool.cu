// a7 / a8 -- graph pooling (models/pooling.py).
//
// Attention pooling is ONE kernel per direction; the reference's [heads, N, F] temporary (pooling.py:150-154) never
// exists.  Forward (default, attn_pool_fwd_stream_kernel): one pass over x with an online softmax per (head, molecule),
// 1 / 2 / 4 warps per molecule, rows streamed through per-warp cp.async rings.  Two-phase forward kernels (rows wider
// than 512 features or heads x F > 2048) and the backward: a CTA owns a molecule, stages its rows of x in shared memory
// with a bulk async copy and does score -> per-(head, molecule) softmax -> weighted sum from there; molecules with more
// rows than the shared-memory chunk are processed in chunks (second pass re-reads x, served by L2).
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

// ------------------------------------------------------------------------------------------ forward, streaming
// One pass over x with an online softmax per (head, molecule): G warps share a molecule (rows sub, sub + G, ...), each
// warp streams its rows through a private ring of kStreamRing rows in shared memory filled with cp.async (every lane
// copies exactly the 16-byte columns it later reads, so the ring needs no barrier at all) and keeps, per head, the
// running maximum m_h, the running sum s_h of exp(z - m_h) and the running weighted sum acc_h = sum_i exp(z_i - m_h) x_i
// in registers (rescaled when the maximum moves).  The G partial states of a molecule are merged once at the end.
// x is read from HBM exactly once and never again (the two-phase kernels above re-read it from L2 for the weighted sum and
// wait for one dependent 16-byte load per lane at a time: 0.07 - 0.23 of the HBM rate); 4 CTAs x 4 warps x 3 rows of
// 2 KB are in flight per SM.  Same arithmetic as pooling.py:134-161 up to the order of the fp32 sums.
constexpr int kStreamRingS = 4, kStreamRingW = 6;
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int K>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(K) : "memory");
}

// packed fp32 pairs (Blackwell FFMA2 / FMUL2): a float4 read from shared memory is two 64-bit units
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float ex2_approx(float v) {         // 2^v, relative error <= 2^-22
  float r;
  asm("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
constexpr float kLog2e = 1.4426950408889634f;

// Sum H per-lane values over the warp with log2(H) halving stages (a lane keeps half of its values and sends the other
// half: H/2 + H/4 + ... shuffles), the remaining butterfly stages on the single value left, and one indexed shuffle per
// head to hand every total to every lane (bitwise identical in all lanes: the branches on them stay warp-uniform).
// 4 heads: 10 shuffles + 6 adds instead of 20 + 20.
template <int H>
__device__ __forceinline__ void warp_sum_heads(float (&d)[H], int lane) {
  static_assert(H == 1 || H == 2 || H == 4 || H == 8, "power of two");
  int off = 16;
#pragma unroll
  for (int nv = H; nv > 1; nv >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int q = 0; q < nv / 2; ++q) {
      const float keep = up ? d[q + nv / 2] : d[q];
      const float send = up ? d[q] : d[q + nv / 2];
      d[q] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off > 0; off >>= 1) d[0] += __shfl_xor_sync(0xffffffffu, d[0], off);
  constexpr int L = H == 1 ? 0 : H == 2 ? 1 : H == 4 ? 2 : 3;
  const float mine = d[0];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    int src = 0;
#pragma unroll
    for (int k = 0; k < L; ++k) src |= ((h >> (L - 1 - k)) & 1) << (4 - k);
    d[h] = __shfl_sync(0xffffffffu, mine, src);
  }
}

// WREG (kept, off: ax2d_attn_pool_fwd_config mode 2): the head weights of a lane's columns live in registers (2 CTAs per
// SM, a ring of 6 rows per warp) instead of being re-read from shared memory for every row (4 CTAs per SM, ring of 4;
// 16 of the 20 LDS.128 per row are weights).  Measured slower, 201 vs 158 us on 4160 drug-like molecules: the kernel is
// bound by instruction issue and fixed-latency dependencies (ncu: issue-active 47 %, stalls wait / short scoreboard),
// not by the shared-memory pipe, and 8 warps per SM hide less of it than 16.
template <int MAXH, int J, bool WREG>
__global__ void __launch_bounds__(kPoolThreads, WREG ? 2 : 4) attn_pool_fwd_stream_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ seg_ptr, int64_t B, int64_t N, int F, int heads, int G,
    const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ temperature,
    float* __restrict__ pooled, float* __restrict__ attn, float* zbuf) {
  pdl_enter();
  constexpr int W4 = 32 * J;                                           // float4 slots of a staged row
  constexpr int kStreamRing = WREG ? kStreamRingW : kStreamRingS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ulonglong2* ws4 = reinterpret_cast<ulonglong2*>(smem_raw);           // [MAXH][W4] head weights, zero beyond heads / F
  ulonglong2* ring = ws4 + (WREG ? 0 : MAXH * W4);                     // [4 warps][kStreamRing][W4]
  float* stat = reinterpret_cast<float*>(ring + 4 * kStreamRing * W4); // [4 warps][2 * MAXH]: m_h, s_h

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int F4 = F >> 2;
  ulonglong2 wr[WREG ? MAXH : 1][WREG ? J : 1];
  if (WREG) {
#pragma unroll
    for (int h = 0; h < (WREG ? MAXH : 1); ++h)
#pragma unroll
      for (int j = 0; j < (WREG ? J : 1); ++j) {
        const int c = lane + 32 * j;
        wr[h][j] = (h < heads && c < F4) ? __ldg(reinterpret_cast<const ulonglong2*>(w) + h * F4 + c) : make_ulonglong2(0ull, 0ull);
      }
  } else {
    for (int i = tid; i < MAXH * W4; i += kPoolThreads) {
      const int h = i / W4, c = i - h * W4;
      reinterpret_cast<float4*>(ws4)[i] =
          (h < heads && c < F4) ? __ldg(reinterpret_cast<const float4*>(w) + h * F4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  ulonglong2* myring = ring + warp * (kStreamRing * W4);
  if (F4 < W4) {                               // columns beyond the row are never copied: zero them once (no masks below)
    for (int i = lane; i < kStreamRing * W4; i += 32)
      if (i % W4 >= F4) myring[i] = make_ulonglong2(0ull, 0ull);
  }
  __syncthreads();
  const float invT = 1.f / __ldg(temperature);
  const int64_t mol = (static_cast<int64_t>(blockIdx.x) * 4 + warp) / G;     // G in {1, 2, 4}: a group never straddles CTAs
  const int sub = warp % G;
  const bool active = mol < B;
  int n0 = 0, n = 0;
  if (active) {
    n0 = seg_ptr[mol];
    n = seg_ptr[mol + 1] - n0;
  }
  const float4* xg = reinterpret_cast<const float4*>(x) + static_cast<int64_t>(n0) * F4;
  const int my_rows = n > sub ? (n - sub + G - 1) / G : 0;

  auto issue = [&](int k) {                    // k-th row of this warp -> ring slot k % kStreamRing (one group per call)
    if (k < my_rows) {
      const float4* src = xg + static_cast<int64_t>(sub + k * G) * F4;
      ulonglong2* dst = myring + (k % kStreamRing) * W4;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c = lane + 32 * j;
        if (c < F4) cp_async16(dst + c, src + c);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int k = 0; k < kStreamRing - 1; ++k) issue(k);

  // scores are kept in log2 units (z * log2 e) so that every exponential is one ex2.approx
  ulonglong2 acc[MAXH][J];
  float m[MAXH], s[MAXH], bias[MAXH];
#pragma unroll
  for (int h = 0; h < MAXH; ++h) {
    m[h] = -INFINITY;
    s[h] = 0.f;
    bias[h] = h < heads ? __ldg(b + h) : 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) acc[h][j] = make_ulonglong2(0ull, 0ull);
  }

  for (int k = 0; k < my_rows; ++k) {
    issue(k + kStreamRing - 1);                // into the slot this lane finished reading in the previous iteration
    cp_async_wait<kStreamRing - 1>();          // this lane's copies of row k have landed (it reads only those)
    const ulonglong2* row = myring + (k % kStreamRing) * W4;
    ulonglong2 xv[J];
#pragma unroll
    for (int j = 0; j < J; ++j) xv[j] = row[lane + 32 * j];
    float dot[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      uint64_t d2 = 0ull;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const ulonglong2 wv = WREG ? wr[WREG ? h : 0][WREG ? j : 0] : ws4[h * W4 + lane + 32 * j];
        d2 = fma2(xv[j].x, wv.x, d2);
        d2 = fma2(xv[j].y, wv.y, d2);
      }
      float lo, hi;
      upk2(d2, lo, hi);
      dot[h] = lo + hi;
    }
    warp_sum_heads<MAXH>(dot, lane);                                   // every lane: all head totals, bitwise identical
    const int i = sub + k * G;
    float zsel = 0.f;
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      if (h < heads) {
        const float z = (dot[h] + bias[h]) * invT;                     // pooling.py:134-140
        const float z2 = z * kLog2e;
        if (lane == h) zsel = z;
        if (z2 > m[h]) {                                               // warp-uniform
          const float r = ex2_approx(m[h] - z2);                       // first row: 2^-inf = 0 on zeros
          const uint64_t r2 = pk2(r, r);
          s[h] *= r;
#pragma unroll
          for (int j = 0; j < J; ++j) {
            acc[h][j].x = mul2(acc[h][j].x, r2);
            acc[h][j].y = mul2(acc[h][j].y, r2);
          }
          m[h] = z2;
        }
        const float p = ex2_approx(z2 - m[h]);
        const uint64_t p2 = pk2(p, p);
        s[h] += p;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          acc[h][j].x = fma2(p2, xv[j].x, acc[h][j].x);
          acc[h][j].y = fma2(p2, xv[j].y, acc[h][j].y);
        }
      }
    }
    if (lane < heads) zbuf[static_cast<int64_t>(lane) * N + n0 + i] = zsel;
  }
  cp_async_wait<0>();
  float4* pooled4 = reinterpret_cast<float4*>(pooled) + mol * F4;
  const float inv_h = 1.f / static_cast<float>(heads);
  auto axpy = [](float4& o, float f, const ulonglong2& v) {
    float a0, a1, a2, a3;
    upk2(v.x, a0, a1);
    upk2(v.y, a2, a3);
    o.x += f * a0; o.y += f * a1; o.z += f * a2; o.w += f * a3;
  };

  // The returned weights (saved for the backward, whose dz = a (da - sum a da) cancels) are NOT taken from the ex2.approx
  // values of the loop: a = expf(z - max) / sum expf(z - max) with the accurate exponential over the molecule's scores
  // (n x heads values from L2; every warp of the group does the two reductions redundantly, then writes its slice).
  auto write_attn = [&](int part, int parts) {
    for (int h = 0; h < heads; ++h) {
      const float* zr = zbuf + static_cast<int64_t>(h) * N + n0;
      float mx = -INFINITY;
      for (int r = lane; r < n; r += 32) mx = fmaxf(mx, __ldcg(zr + r));
      mx = warp_max(mx);
      float sm = 0.f;
      for (int r = lane; r < n; r += 32) sm += expf(__ldcg(zr + r) - mx);
      sm = warp_sum(sm);
      float* ar = attn + static_cast<int64_t>(h) * N + n0;
      for (int r = part * 32 + lane; r < n; r += 32 * parts) ar[r] = expf(__ldcg(zr + r) - mx) / sm;   // pooling.py:144-145
    }
  };

  if (G == 1) {                                // kernel argument: uniform over the grid
    if (!active) return;
    float rs[MAXH];
#pragma unroll
    for (int h = 0; h < MAXH; ++h) rs[h] = (h < heads && n > 0) ? 1.f / s[h] : 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = lane + 32 * j;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);                      // n == 0: torch_scatter leaves the segment at 0
#pragma unroll
      for (int h = 0; h < MAXH; ++h)
        if (h < heads) axpy(o, inv_h * rs[h], acc[h][j]);
      if (c < F4) pooled4[c] = o;
    }
    __syncwarp();                              // the scores written by lanes 0 .. heads-1 are read back by every lane
    write_attn(0, 1);
    return;
  }

  // ---- merge of the G partial states of a molecule (every warp of the CTA takes part in the barriers)
  if (lane == 0) {
#pragma unroll
    for (int h = 0; h < MAXH; ++h) {
      stat[warp * 2 * MAXH + h] = m[h];
      stat[warp * 2 * MAXH + MAXH + h] = s[h];
    }
  }
  __syncthreads();
  const int w0 = warp - sub;
  float mg[MAXH], rsg[MAXH];
#pragma unroll
  for (int h = 0; h < MAXH; ++h) {
    float mm = -INFINITY;
    for (int g = 0; g < G; ++g) mm = fmaxf(mm, stat[(w0 + g) * 2 * MAXH + h]);
    float ss = 0.f;
    for (int g = 0; g < G; ++g) {
      const float sw = stat[(w0 + g) * 2 * MAXH + MAXH + h];
      if (sw > 0.f) ss += sw * ex2_approx(stat[(w0 + g) * 2 * MAXH + h] - mm);   // a warp without rows has s = 0, m = -inf
    }
    mg[h] = mm;
    rsg[h] = ss > 0.f ? 1.f / ss : 0.f;
  }
  float4* mypart = reinterpret_cast<float4*>(myring);   // the ring is drained: its first slot takes this warp's weighted sum
#pragma unroll
  for (int j = 0; j < J; ++j) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int h = 0; h < MAXH; ++h)
      if (h < heads && s[h] > 0.f) axpy(o, ex2_approx(m[h] - mg[h]) * inv_h * rsg[h], acc[h][j]);
    mypart[lane + 32 * j] = o;
  }
  __syncthreads();
  if (!active) return;
  if (sub == 0) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c = lane + 32 * j;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int g = 0; g < G; ++g) {
        const float4 v = reinterpret_cast<const float4*>(ring + (w0 + g) * (kStreamRing * W4))[c];
        o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
      }
      if (c < F4) pooled4[c] = o;
    }
  }
  write_attn(sub, G);
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

// development / test switch: mode 0 = streaming kernel where it applies, 1 = the two-phase kernels only;
// group 0 = warps per molecule chosen from the batch size, 1 / 2 / 4 = forced
static int g_pool_fwd_mode = 0, g_pool_fwd_group = 0, g_pool_fwd_wreg = 0;
extern "C" int ax2d_attn_pool_fwd_config(int mode, int group) {
  AX2D_CHECK_ARG((mode == 0 || mode == 1 || mode == 2) && (group == 0 || group == 1 || group == 2 || group == 4),
                 "ax2d_attn_pool_fwd_config: mode in {0,1,2}, group in {0,1,2,4}");
  g_pool_fwd_wreg = mode == 2 ? 1 : 0;         // mode 2: streaming kernel with the head weights in registers (measured slower)
  if (mode == 2) mode = 0;
  g_pool_fwd_mode = mode;
  g_pool_fwd_group = group;
  return AX2D_OK;
}

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
  {
    // streaming single-pass kernel: rows of up to 512 features, per-head accumulators in registers (heads x F <= 2048)
    const int F4 = F / 4;
    const int J = F4 <= 32 ? 1 : F4 <= 64 ? 2 : F4 <= 128 ? 4 : 0;
    const int MH = heads <= 1 ? 1 : heads <= 2 ? 2 : heads <= 4 ? 4 : 8;
    if (g_pool_fwd_mode == 0 && J != 0 && MH * J <= 16) {
      // warps per molecule: ~12+ rows per warp (the prologue and the merge are per warp), more warps when the batch
      // is too small to fill half of the 148 x 16 warp slots (measured: tools/bench_pool.py)
      int G = g_pool_fwd_group;
      if (G != 1 && G != 2 && G != 4) {
        const int64_t avg = N / B;
        G = avg >= 48 ? 4 : avg >= 24 ? 2 : 1;
        while (G < 4 && B * G < 1184 && avg / G >= 8) G *= 2;
      }
      const unsigned grid = static_cast<unsigned>((B * G + 3) / 4);
      cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define AX2D_POOL_STREAM(MH_, J_)                                                                                          \
  do {                                                                                                                     \
    if (wreg) {                                                                                                            \
      const size_t sm = static_cast<size_t>(4 * kStreamRingW * 32 * J_) * 16 + 4 * 2 * MH_ * 4;                            \
      if (sm > 48 * 1024)                                                                                                  \
        cudaFuncSetAttribute(attn_pool_fwd_stream_kernel<MH_, J_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                             static_cast<int>(sm));                                                                        \
      launch_k(attn_pool_fwd_stream_kernel<MH_, J_, true>, dim3(grid), dim3(kPoolThreads), sm, st, x, seg_ptr, B, N, F,    \
               heads, G, w, b, temperature, pooled, attn, z);                                                              \
    } else {                                                                                                               \
      launch_k(attn_pool_fwd_stream_kernel<MH_, J_, false>, dim3(grid), dim3(kPoolThreads),                                \
               static_cast<size_t>((MH_ * 32 * J_ + 4 * kStreamRingS * 32 * J_) * 16 + 4 * 2 * MH_ * 4), st, x, seg_ptr,   \
               B, N, F, heads, G, w, b, temperature, pooled, attn, z);                                                     \
    }                                                                                                                      \
  } while (0)
      const bool wreg = g_pool_fwd_wreg != 0;
      switch (MH * 8 + J) {
        case 1 * 8 + 1: AX2D_POOL_STREAM(1, 1); break;
        case 1 * 8 + 2: AX2D_POOL_STREAM(1, 2); break;
        case 1 * 8 + 4: AX2D_POOL_STREAM(1, 4); break;
        case 2 * 8 + 1: AX2D_POOL_STREAM(2, 1); break;
        case 2 * 8 + 2: AX2D_POOL_STREAM(2, 2); break;
        case 2 * 8 + 4: AX2D_POOL_STREAM(2, 4); break;
        case 4 * 8 + 1: AX2D_POOL_STREAM(4, 1); break;
        case 4 * 8 + 2: AX2D_POOL_STREAM(4, 2); break;
        case 4 * 8 + 4: AX2D_POOL_STREAM(4, 4); break;
        case 8 * 8 + 1: AX2D_POOL_STREAM(8, 1); break;
        default: AX2D_POOL_STREAM(8, 2); break;
      }
#undef AX2D_POOL_STREAM
      return launch_status("ax2d_attn_pool_fwd");
    }
  }
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
