// a4 charge equilibration (models/gnn.py:622-658) and a5 stereo features (models/gnn.py:387-509).
// Per-molecule segment work; every reduction has a fixed order (no atomics).
#include "common.cuh"

namespace ax2d {

__device__ __forceinline__ float block_sum_128(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w];
  return s;
}

// ----------------------------------------------------------------------------------------- charge eq.
__global__ void __launch_bounds__(128) charge_eq_fwd_kernel(const float* __restrict__ x, int64_t ldx,
                                                            const int32_t* __restrict__ seg_ptr, int W4,
                                                            const float* __restrict__ total_charges,
                                                            float* __restrict__ out, int64_t ldo,
                                                            float* __restrict__ stats) {
  pdl_enter();
  __shared__ float red[4];
  const int g = blockIdx.x;
  const int n0 = seg_ptr[g], n1 = seg_ptr[g + 1];
  float q = 0.f, f = 0.f;
  for (int i = n0 + threadIdx.x; i < n1; i += blockDim.x) {
    q += x[static_cast<int64_t>(i) * ldx];
    f += fmaxf(x[static_cast<int64_t>(i) * ldx + 1], 1e-6f);           // gnn.py:634
  }
  const float Q = block_sum_128(q, red);
  const float Fs = block_sum_128(f, red);
  const float Fu = fmaxf(Fs + 1e-6f, 1e-6f);                            // gnn.py:642-646
  const float dQ = total_charges[g] - Q;                                // gnn.py:649
  if (threadIdx.x == 0) {
    stats[2 * g] = dQ;
    stats[2 * g + 1] = Fu;
  }
  const int total = (n1 - n0) * W4;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int i = n0 + t / W4, c = t % W4;
    float4 v = __ldg(reinterpret_cast<const float4*>(x + static_cast<int64_t>(i) * ldx) + c);
    if (c == 0) {
      const float fn = fmaxf(v.y, 1e-6f) / Fu;                          // gnn.py:655
      v.x = v.x + fn * dQ;                                              // gnn.py:656
      v.y = fn;
    }
    reinterpret_cast<float4*>(out + static_cast<int64_t>(i) * ldo)[c] = v;
  }
}

__global__ void __launch_bounds__(128) charge_eq_bwd_kernel(const float* __restrict__ x, int64_t ldx,
                                                            const float* __restrict__ out, int64_t ldo,
                                                            const int32_t* __restrict__ seg_ptr, int W4,
                                                            const float* __restrict__ stats,
                                                            const float* __restrict__ g_out, int64_t ldg,
                                                            float* __restrict__ gx, int64_t ldgx) {
  pdl_enter();
  __shared__ float red[4];
  const int g = blockIdx.x;
  const int n0 = seg_ptr[g], n1 = seg_ptr[g + 1];
  const float dQ = stats[2 * g], Fu = stats[2 * g + 1];
  float s1 = 0.f, s2 = 0.f;
  for (int i = n0 + threadIdx.x; i < n1; i += blockDim.x) {
    const float gq = g_out[static_cast<int64_t>(i) * ldg], gf = g_out[static_cast<int64_t>(i) * ldg + 1];
    const float fn = out[static_cast<int64_t>(i) * ldo + 1];
    s1 += (gq * dQ + gf) * fn;
    s2 += gq * fn;
  }
  const float S1 = block_sum_128(s1, red);
  const float S2 = block_sum_128(s2, red);
  const int total = (n1 - n0) * W4;
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int i = n0 + t / W4, c = t % W4;
    float4 v = __ldg(reinterpret_cast<const float4*>(g_out + static_cast<int64_t>(i) * ldg) + c);
    if (c == 0) {
      const float a = v.x * dQ + v.y;
      const float fraw = x[static_cast<int64_t>(i) * ldx + 1];
      v.y = fraw >= 1e-6f ? (a - S1) / Fu : 0.f;                        // clamp(min) passes grad where f >= min
      v.x = v.x - S2;
    }
    reinterpret_cast<float4*>(gx + static_cast<int64_t>(i) * ldgx)[c] = v;
  }
}

// ----------------------------------------------------------------------------------------- tetrahedral
constexpr int kTetraMaxV = 16;   // columns per lane: width <= 512

struct TetraRows {
  float e[4][kTetraMaxV];
  float nrm[4];
  float sigma;
};

// warp-cooperative: normalised rows e_i, norms, sigma = tanh(mean norm / 3)      (gnn.py:409-443)
__device__ __forceinline__ void tetra_load(const float* __restrict__ x, int64_t ldx, const int32_t* __restrict__ idx4,
                                           int true_width, int lane, TetraRows& t) {
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float* row = x + static_cast<int64_t>(idx4[i]) * ldx;
    float ss = 0.f;
#pragma unroll
    for (int v = 0; v < kTetraMaxV; ++v) {
      const int c = lane + 32 * v;
      const float val = c < true_width ? __ldg(row + c) : 0.f;
      t.e[i][v] = val;
      ss += val * val;
    }
    const float n = sqrtf(warp_sum(ss));
    t.nrm[i] = n;
    mean += n;
    const float d = fmaxf(n, 1e-8f);                                     // F.normalize eps
#pragma unroll
    for (int v = 0; v < kTetraMaxV; ++v) t.e[i][v] = t.e[i][v] / d;
  }
  t.sigma = tanhf(mean * 0.25f / 3.0f);
}

__global__ void __launch_bounds__(128) tetra_contrib_kernel(const float* __restrict__ x, int64_t ldx, int true_width,
                                                            int width, const int32_t* __restrict__ idx, int64_t M,
                                                            float* __restrict__ contrib) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t m = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  TetraRows t;
  tetra_load(x, ldx, idx + 4 * m, true_width, lane, t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = (i + 1) & 3, b = (i + 2) & 3, c = (i + 3) & 3;
    float* o = contrib + (4 * m + i) * width;
#pragma unroll
    for (int v = 0; v < kTetraMaxV; ++v) {
      const int col = lane + 32 * v;
      if (col < width) {
        const float ea = t.e[a][v], eb = t.e[b][v], ec = t.e[c][v];
        const float ci = ea * ea * (eb - ec) + eb * eb * (ec - ea) + ec * ec * (ea - eb);   // gnn.py:429-433
        o[col] = col < true_width ? ci * t.sigma : 0.f;
      }
    }
  }
}

// out[a] = listed(a) ? x[a] + sum_{slots of a} rows[slot] : 0        (gnn.py:449-460, index_add_ order)
__global__ void __launch_bounds__(256) tetra_apply_kernel(const float* __restrict__ x, int64_t ldx, int64_t N, int W4,
                                                          const int32_t* __restrict__ slot_ptr,
                                                          const int32_t* __restrict__ slot_idx,
                                                          const float* __restrict__ rows, float* __restrict__ out,
                                                          int64_t ldo) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t a = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (a >= N) return;
  const int beg = __ldg(slot_ptr + a), end = __ldg(slot_ptr + a + 1);
  for (int c = lane; c < W4; c += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (end > beg) {
      acc = __ldg(reinterpret_cast<const float4*>(x + a * ldx) + c);
      for (int k = beg; k < end; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(rows + static_cast<int64_t>(__ldg(slot_idx + k)) * W4 * 4) + c);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
    }
    reinterpret_cast<float4*>(out + a * ldo)[c] = acc;
  }
}

// backward of the per-centre map: g_rows[4m+j] = d L / d x[idx[m,j]] through centre m   
__global__ void __launch_bounds__(128) tetra_bwd_rows_kernel(const float* __restrict__ x, int64_t ldx, int true_width,
                                                             int width, const int32_t* __restrict__ idx, int64_t M,
                                                             const float* __restrict__ g_out, int64_t ldg,
                                                             float* __restrict__ g_rows) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t m = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  TetraRows t;
  const int32_t* id = idx + 4 * m;
  tetra_load(x, ldx, id, true_width, lane, t);
  float ge[4][kTetraMaxV];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int v = 0; v < kTetraMaxV; ++v) ge[i][v] = 0.f;
  float gsig = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = (i + 1) & 3, b = (i + 2) & 3, c = (i + 3) & 3;
    const float* hrow = g_out + static_cast<int64_t>(id[i]) * ldg;
#pragma unroll
    for (int v = 0; v < kTetraMaxV; ++v) {
      const int col = lane + 32 * v;
      const float h = col < true_width ? __ldg(hrow + col) : 0.f;
      const float ea = t.e[a][v], eb = t.e[b][v], ec = t.e[c][v];
      const float sa = ea * ea, sb = eb * eb, sc = ec * ec;
      const float ci = sa * (eb - ec) + sb * (ec - ea) + sc * (ea - eb);
      gsig += h * ci;
      const float gc = t.sigma * h;
      ge[a][v] += gc * (2.f * ea * (eb - ec) + sc - sb);
      ge[b][v] += gc * (2.f * eb * (ec - ea) + sa - sc);
      ge[c][v] += gc * (2.f * ec * (ea - eb) + sb - sa);
    }
  }
  gsig = warp_sum(gsig);
  const float ks = gsig * (1.f - t.sigma * t.sigma) * (1.f / 12.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float dot = 0.f;
#pragma unroll
    for (int v = 0; v < kTetraMaxV; ++v) dot += t.e[j][v] * ge[j][v];
    dot = warp_sum(dot);
    const float n = t.nrm[j];
    float* o = g_rows + (4 * m + j) * width;
#pragma unroll
    for (int v = 0; v < kTetraMaxV; ++v) {
      const int col = lane + 32 * v;
      if (col < width) {
        float r;
        if (n > 1e-8f) r = (ge[j][v] - t.e[j][v] * dot) / n + ks * t.e[j][v];
        else r = ge[j][v] * 1e8f;          // clamp active: e = r / eps, norm gradient is 0 at r = 0
        o[col] = col < true_width ? r : 0.f;
      }
    }
  }
}

// gx[a] = listed(a) ? g_out[a] + sum_{slots of a} g_rows[slot] : 0  -- same shape as tetra_apply_kernel.

// ----------------------------------------------------------------------------------------- cis / trans
struct CisTransUpd {
  int n;
};
__global__ void __launch_bounds__(256) cistrans_kernel(const float* __restrict__ x, int64_t ldx, int64_t N, int W4,
                                                       const int32_t* __restrict__ u_src,
                                                       const int32_t* __restrict__ u_tgt,
                                                       const float* __restrict__ u_sign, int n_upd, int transpose,
                                                       float* __restrict__ out, int64_t ldo) {
  pdl_enter();
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= N * W4) return;
  const int64_t r = t / W4;
  const int c = static_cast<int>(t % W4);
  float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ldx) + c);
  for (int j = 0; j < n_upd; ++j) {                                    // gnn.py:496-505, sequential
    const int dst = transpose ? u_src[j] : u_tgt[j];
    if (dst == r) {
      const int from = transpose ? u_tgt[j] : u_src[j];
      const float s = u_sign[j];
      const float4 w = __ldg(reinterpret_cast<const float4*>(x + static_cast<int64_t>(from) * ldx) + c);
      v.x += s * w.x; v.y += s * w.y; v.z += s * w.z; v.w += s * w.w;
    }
  }
  reinterpret_cast<float4*>(out + r * ldo)[c] = v;
}

}  // namespace ax2d

using namespace ax2d;

extern "C" int ax2d_charge_eq_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int64_t B, int width,
                                  const float* total_charges, float* out, int64_t ldo, float* stats,
                                  ax2d_stream_t stream) {
  AX2D_CHECK_ARG(width >= 4 && width % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "ax2d_charge_eq_fwd: widths % 4");
  AX2D_CHECK_ALIGN(x);
  AX2D_CHECK_ALIGN(out);
  if (B <= 0) return AX2D_OK;
  launch_k(charge_eq_fwd_kernel, dim3(static_cast<unsigned>(B)), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, ldx, seg_ptr, width / 4, total_charges, out, ldo, stats);
  return launch_status("ax2d_charge_eq_fwd");
}

extern "C" int ax2d_charge_eq_bwd(const float* x, int64_t ldx, const float* out, int64_t ldo, const int32_t* seg_ptr,
                                  int64_t B, int width, const float* stats, const float* g_out, int64_t ldg, float* gx,
                                  int64_t ldgx, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(width >= 4 && width % 4 == 0 && ldg % 4 == 0 && ldgx % 4 == 0, "ax2d_charge_eq_bwd: widths % 4");
  AX2D_CHECK_ALIGN(g_out);
  AX2D_CHECK_ALIGN(gx);
  if (B <= 0) return AX2D_OK;
  launch_k(charge_eq_bwd_kernel, dim3(static_cast<unsigned>(B)), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, ldx, out, ldo, seg_ptr, width / 4, stats, g_out, ldg, gx, ldgx);
  return launch_status("ax2d_charge_eq_bwd");
}

extern "C" int ax2d_tetra_fwd(const float* x, int64_t ldx, int64_t N, int width, int true_width, const int32_t* idx,
                              int64_t M, const int32_t* slot_ptr, const int32_t* slot_idx, float* contrib, float* out,
                              int64_t ldo, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(width % 4 == 0 && true_width <= width && width <= 32 * kTetraMaxV,
                 "ax2d_tetra_fwd: width %d (true %d) unsupported", width, true_width);
  AX2D_CHECK_ARG(ldx % 4 == 0 && ldo % 4 == 0, "ax2d_tetra_fwd: leading dimensions % 4");
  AX2D_CHECK_ALIGN(x);
  AX2D_CHECK_ALIGN(out);
  AX2D_CHECK_ALIGN(contrib);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M > 0) {
    launch_k(tetra_contrib_kernel, dim3(static_cast<unsigned>((M + 3) / 4)), dim3(128), 0, st, x, ldx, true_width, width, idx, M, contrib);
    int rc = launch_status("ax2d_tetra_fwd(contrib)");
    if (rc != AX2D_OK) return rc;
  }
  if (N > 0)
    launch_k(tetra_apply_kernel, dim3(static_cast<unsigned>((N + 7) / 8)), dim3(256), 0, st, x, ldx, N, width / 4, slot_ptr, slot_idx,
                                                                           contrib, out, ldo);
  return launch_status("ax2d_tetra_fwd(apply)");
}

extern "C" int ax2d_tetra_bwd(const float* x, int64_t ldx, int64_t N, int width, int true_width, const int32_t* idx,
                              int64_t M, const int32_t* slot_ptr, const int32_t* slot_idx, const float* g_out,
                              int64_t ldg, float* g_rows, float* gx, int64_t ldgx, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(width % 4 == 0 && true_width <= width && width <= 32 * kTetraMaxV,
                 "ax2d_tetra_bwd: width %d (true %d) unsupported", width, true_width);
  AX2D_CHECK_ARG(ldx % 4 == 0 && ldg % 4 == 0 && ldgx % 4 == 0, "ax2d_tetra_bwd: leading dimensions % 4");
  AX2D_CHECK_ALIGN(g_out);
  AX2D_CHECK_ALIGN(gx);
  AX2D_CHECK_ALIGN(g_rows);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (M > 0) {
    launch_k(tetra_bwd_rows_kernel, dim3(static_cast<unsigned>((M + 3) / 4)), dim3(128), 0, st, x, ldx, true_width, width, idx, M, g_out,
                                                                              ldg, g_rows);
    int rc = launch_status("ax2d_tetra_bwd(rows)");
    if (rc != AX2D_OK) return rc;
  }
  if (N > 0)
    launch_k(tetra_apply_kernel, dim3(static_cast<unsigned>((N + 7) / 8)), dim3(256), 0, st, g_out, ldg, N, width / 4, slot_ptr, slot_idx,
                                                                           g_rows, gx, ldgx);
  return launch_status("ax2d_tetra_bwd(apply)");
}

extern "C" int ax2d_cistrans(const float* x, int64_t ldx, int64_t N, int width, const int32_t* upd_src,
                             const int32_t* upd_tgt, const float* upd_sign, int n_upd, int transpose, float* out,
                             int64_t ldo, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(width % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "ax2d_cistrans: widths % 4");
  AX2D_CHECK_ARG(n_upd >= 0 && n_upd <= 64, "ax2d_cistrans: n_upd=%d out of range", n_upd);
  AX2D_CHECK_ALIGN(x);
  AX2D_CHECK_ALIGN(out);
  if (N <= 0) return AX2D_OK;
  const int64_t total = N * (width / 4);
  launch_k(cistrans_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, ldx, N, width / 4, upd_src, upd_tgt, upd_sign, n_upd, transpose, out, ldo);
  return launch_status("ax2d_cistrans");
}
