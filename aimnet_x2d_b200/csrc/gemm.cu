// Dense projections of the hot path (every nn.Linear: layers.py:47-61,192-199; gnn.py:96-146,190-195).
//
// This file is the exact-fp32 SIMT path (FFMA, fp32 accumulate): the fp32 configuration must match the
// reference within 1e-5, which rules out plain TF32.  Operands are "segmented" matrices so that the
// reference's torch.cat calls (layers.py:79, gnn.py:245,257,323) are never materialised, and the
// epilogue fuses bias, activation, dropout, residual adds and the activation backward.
//
// Tiling: CTA 128x64, BK = 16, 256 threads, 8x4 register tile per thread, shared-memory double buffer
// with register prefetch of the next k-tile.
#include "gemm_common.cuh"

namespace ax2d {

constexpr int BM = 128, BN = 64, BK = 16, GEMM_THREADS = 256;
constexpr int AS_LD = BM + 4, BS_LD = BN + 4;

struct GemmArgs {
  SegView a, b;
  EpiArgs e;
  int64_t M, N, K;
  int64_t k_begin_stride;     // split-k: k range per blockIdx.z (multiple of BK); 0 if no split
  float* ws;                  // split-k partial output [split][M][N]
};

// Loads one float4 of a logical operand element block.
//  REDUCE_COLS: operand is [rows = m or n][cols = k]  (k contiguous);  float4 spans k..k+3 of row `r`.
//  else:        operand is [rows = k][cols = m or n]  (m/n contiguous); float4 spans c..c+3 of row `k`.
template <bool REDUCE_COLS>
__device__ __forceinline__ float4 load_op(const SegView& v, int64_t r, int64_t rows, int c, int cols) {
  float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r >= rows || c >= cols) return z;
  const int s = find_seg(v.start, v.n_seg, c);
  return __ldg(reinterpret_cast<const float4*>(v.ptr[s] + r * v.ld[s] + (c - v.start[s])));
}

template <bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_kernel(const __grid_constant__ GemmArgs g) {
  pdl_enter();
  __shared__ __align__(16) float As[2][BK][AS_LD];
  __shared__ __align__(16) float Bs[2][BK][BS_LD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM;
  const int n0 = blockIdx.x * BN;
  int64_t k_lo = 0, k_hi = g.K;
  if (g.k_begin_stride > 0) {
    k_lo = static_cast<int64_t>(blockIdx.z) * g.k_begin_stride;
    k_hi = k_lo + g.k_begin_stride < g.K ? k_lo + g.k_begin_stride : g.K;
  }

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb;
  auto fetch = [&](int64_t k0) {
    if (!TA) {      // A[m][k], k contiguous: thread -> (row = tid/4 + 64 i, kq = tid%4)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t r = m0 + (tid >> 2) + 64 * i;
        const int64_t k = k0 + (tid & 3) * 4;
        ra[i] = (k < k_hi) ? load_op<true>(g.a, r, g.M, static_cast<int>(k), static_cast<int>(g.K)) : make_float4(0, 0, 0, 0);
      }
    } else {        // A stored [k][m], m contiguous: thread -> (k = tid/32 + 8 i, mq = tid%32)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t k = k0 + (tid >> 5) + 8 * i;
        const int64_t m = m0 + (tid & 31) * 4;
        ra[i] = (k < k_hi) ? load_op<false>(g.a, k, g.K, static_cast<int>(m), static_cast<int>(g.M)) : make_float4(0, 0, 0, 0);
      }
    }
    if (TB) {       // B[k][n] = W[n][k], k contiguous: thread -> (n = tid/4, kq = tid%4)
      const int64_t r = n0 + (tid >> 2);
      const int64_t k = k0 + (tid & 3) * 4;
      rb = (k < k_hi) ? load_op<true>(g.b, r, g.N, static_cast<int>(k), static_cast<int>(g.K)) : make_float4(0, 0, 0, 0);
    } else {        // B stored [k][n], n contiguous: thread -> (k = tid/16, nq = tid%16)
      const int64_t k = k0 + (tid >> 4);
      const int n = n0 + (tid & 15) * 4;
      rb = (k < k_hi) ? load_op<false>(g.b, k, g.K, n, static_cast<int>(g.N)) : make_float4(0, 0, 0, 0);
    }
  };
  auto stash = [&](int buf) {
    if (!TA) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = (tid >> 2) + 64 * i, kq = (tid & 3) * 4;
        As[buf][kq + 0][r] = ra[i].x; As[buf][kq + 1][r] = ra[i].y;
        As[buf][kq + 2][r] = ra[i].z; As[buf][kq + 3][r] = ra[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 2; ++i)
        *reinterpret_cast<float4*>(&As[buf][(tid >> 5) + 8 * i][(tid & 31) * 4]) = ra[i];
    }
    if (TB) {
      const int r = tid >> 2, kq = (tid & 3) * 4;
      Bs[buf][kq + 0][r] = rb.x; Bs[buf][kq + 1][r] = rb.y; Bs[buf][kq + 2][r] = rb.z; Bs[buf][kq + 3][r] = rb.w;
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][tid >> 4][(tid & 15) * 4]) = rb;
    }
  };

  const int64_t n_tiles = (k_hi - k_lo + BK - 1) / BK;
  if (n_tiles > 0) {
    fetch(k_lo);
    stash(0);
  }
  __syncthreads();
  for (int64_t t = 0; t < n_tiles; ++t) {
    const int buf = static_cast<int>(t & 1);
    if (t + 1 < n_tiles) fetch(k_lo + (t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (t + 1 < n_tiles) stash(buf ^ 1);
    __syncthreads();
  }

  // ------------------------------------------------------------------ epilogue
  const int n = n0 + tx * 4;
  if (n >= g.N) return;
  if (g.ws != nullptr) {       // split-k: raw partials, reduced later in a fixed order
    float* w = g.ws + static_cast<int64_t>(blockIdx.z) * g.M * g.N;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = m0 + ty * 4 + (i & 3) + 64 * (i >> 2);
      if (m < g.M) *reinterpret_cast<float4*>(w + m * g.N + n) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    return;
  }
  const EpiCtx cx = epi_ctx(g.e);
  const EpiCol col = epi_col(g.e, n);
  if constexpr (TA) {      // weight gradients: plain store / accumulate only (checked on the host)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = m0 + ty * 4 + (i & 3) + 64 * (i >> 2);
      if (m >= g.M) continue;
      EpiOperands o;
      epi_prefetch(g.e, col, m, o);
      epi_finish<AX2D_ACT_NONE, AX2D_ACT_NONE, false>(g.e, cx, col, m, o, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
  } else {
    const bool dropping = cx.dropping;
    AX2D_EPI_DISPATCH(g.e.act, g.e.dact, dropping, {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ty * 4 + (i & 3) + 64 * (i >> 2);
        if (m < g.M) {
          EpiOperands o;
          epi_prefetch(g.e, col, m, o);
          epi_finish<ACT, DACT, DROP>(g.e, cx, col, m, o, acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
      }
    });
  }
}

// Fixed-order reduction of `split` partial [M, N] tiles.  Eight lanes share one float4 of the output: lane z sums the
// partials z, z + 8, ... and a three-step shuffle tree combines them -- the same order on every run (deterministic),
// and 8x the threads of a one-thread-per-float4 loop, which for the 160 x 160 weight gradients (37 partials, 6400
// float4) kept only 25 CTAs of the 148 SMs busy with serial dependent loads.
// vec_ws / vec_out (optional): `split` partial vectors of length M reduced the same way (fused bias gradient)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int split, int64_t M,
                                                            int64_t N, SegOut c, int accumulate,
                                                            const float* __restrict__ vec_ws, float* __restrict__ vec_out) {
  pdl_enter();
  const int64_t t = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;   // float4 index
  const int z0 = threadIdx.x & 7;
  const int64_t N4 = N >> 2;
  const int64_t total4 = M * N4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool is_vec = t >= total4;
  const int64_t vi = (t - total4) * 4;                 // first of four bias entries handled by this lane group
  if (!is_vec) {
    const int64_t m = t / N4;
    const int n = static_cast<int>(t % N4) * 4;
    for (int z = z0; z < split; z += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ws + (static_cast<int64_t>(z) * M + m) * N + n));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  } else if (vec_ws != nullptr && vi < M) {
    for (int z = z0; z < split; z += 8) {
      const float* p = vec_ws + static_cast<int64_t>(z) * M + vi;
      s.x += __ldg(p);
      if (vi + 1 < M) s.y += __ldg(p + 1);
      if (vi + 2 < M) s.z += __ldg(p + 2);
      if (vi + 3 < M) s.w += __ldg(p + 3);
    }
  }
#pragma unroll
  for (int d = 4; d >= 1; d >>= 1) {
    s.x += __shfl_xor_sync(0xffffffffu, s.x, d);
    s.y += __shfl_xor_sync(0xffffffffu, s.y, d);
    s.z += __shfl_xor_sync(0xffffffffu, s.z, d);
    s.w += __shfl_xor_sync(0xffffffffu, s.w, d);
  }
  if (z0 != 0) return;
  if (is_vec) {
    if (vec_ws == nullptr || vi >= M) return;
    vec_out[vi] = s.x;
    if (vi + 1 < M) vec_out[vi + 1] = s.y;
    if (vi + 2 < M) vec_out[vi + 2] = s.z;
    if (vi + 3 < M) vec_out[vi + 3] = s.w;
    return;
  }
  const int64_t m = t / N4;
  const int n = static_cast<int>(t % N4) * 4;
  const int cs = find_seg(c.start, c.n_seg, n);
  float4* dst = reinterpret_cast<float4*>(c.ptr[cs] + m * c.ld[cs] + (n - c.start[cs]));
  if (accumulate) {
    const float4 o = *dst;
    s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w;
  }
  *dst = s;
}

// column sums, two passes with a fixed order.  Pass 1: a CTA covers 128 columns x kColsumRows rows; a warp reads
// 512 contiguous bytes of one row (lane = float4 column), the 8 warps take rows r, r+8, ...; their partial sums
// are combined through shared memory in warp order.
constexpr int kColsumRows = 256;
__global__ void __launch_bounds__(256) colsum_partial_kernel(SegView a, int64_t M, int N4, float* __restrict__ partial) {
  pdl_enter();
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kColsumRows;
  const int64_t r1 = r0 + kColsumRows < M ? r0 + kColsumRows : M;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < N4) {
    const int s = find_seg(a.start, a.n_seg, c * 4);
    const float* base = a.ptr[s] + (c * 4 - a.start[s]);
    const int64_t ld = a.ld[s];
#pragma unroll 4
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + r * ld));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < N4) {
    float4 t = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 v = red[w][lane];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    reinterpret_cast<float4*>(partial + static_cast<int64_t>(blockIdx.y) * N4 * 4)[c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int n_part, int N, float* __restrict__ out,
                                    int accumulate) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int p = 0; p < n_part; ++p) s += partial[static_cast<int64_t>(p) * N + c];
  out[c] = accumulate ? out[c] + s : s;
}

__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ gr, int64_t ldg,
                                                      const float* __restrict__ pre, int64_t ldp,
                                                      float* __restrict__ out, int64_t ldo, int64_t M, int W4, int act) {
  pdl_enter();
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= M * W4) return;
  const int64_t m = t / W4;
  const int c = static_cast<int>(t % W4);
  const float4 gv = __ldg(reinterpret_cast<const float4*>(gr + m * ldg) + c);
  const float4 pv = __ldg(reinterpret_cast<const float4*>(pre + m * ldp) + c);
  reinterpret_cast<float4*>(out + m * ldo)[c] = make_float4(gv.x * act_bwd(act, pv.x), gv.y * act_bwd(act, pv.y),
                                                            gv.z * act_bwd(act, pv.z), gv.w * act_bwd(act, pv.w));
}
__global__ void tick_kernel(unsigned long long* c) {
  pdl_enter(); c[0] += 1ull; }

int to_view(const ax2d_cmat* m, SegView* v, int64_t total, const char* what) {
  int acc = 0;
  if (m->n_seg < 1 || m->n_seg > AX2D_MAX_SEG) {
    set_error("ax2d_gemm: %s has %d segments", what, m->n_seg);
    return AX2D_ERR_ARG;
  }
  v->n_seg = m->n_seg;
  for (int s = 0; s < m->n_seg; ++s) {
    if (m->width[s] <= 0 || m->width[s] % 4 != 0 || m->ld[s] % 4 != 0) {
      set_error("ax2d_gemm: %s segment %d width/ld must be multiples of 4", what, s);
      return AX2D_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(m->ptr[s]) & 15u) != 0) {
      set_error("ax2d_gemm: %s segment %d is not 16-byte aligned", what, s);
      return AX2D_ERR_ALIGN;
    }
    v->ptr[s] = m->ptr[s];
    v->ld[s] = m->ld[s];
    v->start[s] = acc;
    acc += m->width[s];
  }
  for (int s = m->n_seg; s <= AX2D_MAX_SEG; ++s) v->start[s] = acc;
  if (acc != total) {
    set_error("ax2d_gemm: %s segments cover %d columns, expected %lld", what, acc, (long long)total);
    return AX2D_ERR_ARG;
  }
  return AX2D_OK;
}
int to_out(const ax2d_mat* m, SegOut* v, int64_t total, const char* what, bool allow_null) {
  int acc = 0;
  if (m->n_seg < 1 || m->n_seg > AX2D_MAX_SEG) {
    set_error("ax2d_gemm: %s has %d segments", what, m->n_seg);
    return AX2D_ERR_ARG;
  }
  v->n_seg = m->n_seg;
  for (int s = 0; s < m->n_seg; ++s) {
    if (m->width[s] <= 0 || m->width[s] % 4 != 0 || m->ld[s] % 4 != 0) {
      set_error("ax2d_gemm: %s segment %d width/ld must be multiples of 4", what, s);
      return AX2D_ERR_ARG;
    }
    if (m->ptr[s] == nullptr && !allow_null) {
      set_error("ax2d_gemm: %s segment %d is NULL", what, s);
      return AX2D_ERR_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(m->ptr[s]) & 15u) != 0) {
      set_error("ax2d_gemm: %s segment %d is not 16-byte aligned", what, s);
      return AX2D_ERR_ALIGN;
    }
    v->ptr[s] = m->ptr[s];
    v->ld[s] = m->ld[s];
    v->start[s] = acc;
    acc += m->width[s];
  }
  for (int s = m->n_seg; s <= AX2D_MAX_SEG; ++s) v->start[s] = acc;
  if (acc != total) {
    set_error("ax2d_gemm: %s segments cover %d columns, expected %lld", what, acc, (long long)total);
    return AX2D_ERR_ARG;
  }
  return AX2D_OK;
}

int splitk_reduce(const float* ws, int split, int64_t M, int64_t N, const SegOut& c, int accumulate, const float* vec_ws,
                  float* vec_out, cudaStream_t st) {
  const int64_t total = 8 * (M * (N / 4) + (vec_ws != nullptr ? (M + 3) / 4 : 0));      // 8 lanes per float4
  launch_k(splitk_reduce_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, ws, split, M, N, c, accumulate, vec_ws,
                                                                                  vec_out);
  return launch_status("split-k reduce");
}

int fill_epilogue(const ax2d_epilogue* ep, int64_t M, int64_t N, EpiArgs* e) {
  // e->c has been set by the caller; everything else starts from zero
  SegOut keep = e->c;
  memset(e, 0, sizeof(*e));
  e->c = keep;
  e->M = M;
  e->N = N;
  e->dact_cols = static_cast<int>(N);
  e->act_cols = static_cast<int>(N);
  if (ep == nullptr) return AX2D_OK;
  int rc;
  e->bias = ep->bias;
  if (ep->pre.n_seg > 0 && (rc = to_out(&ep->pre, &e->pre, N, "pre", true)) != AX2D_OK) return rc;
  e->act = ep->act; e->act_cols = ep->act_cols;
  e->mask = ep->mask; e->ld_mask = ep->ld_mask;
  e->drop_p = ep->drop_p; e->drop_seed = ep->drop_seed; e->drop_tick = ep->drop_tick;
  AX2D_CHECK_ARG(ep->resid.n_seg >= 0 && ep->resid.n_seg <= AX2D_MAX_SEG, "ax2d_gemm: too many residuals");
  e->n_resid = ep->resid.n_seg;
  for (int r = 0; r < e->n_resid; ++r) {
    AX2D_CHECK_ALIGN(ep->resid.ptr[r]);
    AX2D_CHECK_ARG(ep->resid.ld[r] % 4 == 0, "ax2d_gemm: residual ld %% 4");
    e->resid[r] = ep->resid.ptr[r];
    e->ld_resid[r] = ep->resid.ld[r];
    e->resid_cols[r] = ep->resid.width[r] > 0 ? ep->resid.width[r] : static_cast<int>(N);
    AX2D_CHECK_ARG(e->resid_cols[r] % 4 == 0, "ax2d_gemm: residual width %% 4");
  }
  e->dact_pre = ep->dact_pre; e->ld_dact = ep->ld_dact; e->dact = ep->dact_pre != nullptr ? ep->dact : AX2D_ACT_NONE;
  e->dact_cols = ep->dact_cols > 0 ? ep->dact_cols : static_cast<int>(N);
  e->accumulate = ep->accumulate;
  AX2D_CHECK_ALIGN(ep->bias);
  AX2D_CHECK_ALIGN(ep->mask);
  AX2D_CHECK_ALIGN(ep->dact_pre);
  AX2D_CHECK_ARG(ep->drop_p >= 0.f && ep->drop_p < 1.f, "ax2d_gemm: dropout p=%f", ep->drop_p);
  AX2D_CHECK_ARG(ep->mask == nullptr || ep->ld_mask % 4 == 0, "ax2d_gemm: mask ld %% 4");
  AX2D_CHECK_ARG(e->act == AX2D_ACT_NONE || e->dact == AX2D_ACT_NONE, "ax2d_gemm: act and dact cannot be combined");
  AX2D_CHECK_ARG(e->act >= AX2D_ACT_NONE && e->act <= AX2D_ACT_SILU && e->dact >= AX2D_ACT_NONE && e->dact <= AX2D_ACT_SILU,
                 "ax2d_gemm: unknown activation code");
  AX2D_CHECK_ARG(e->act_cols % 4 == 0 && e->dact_cols % 4 == 0, "ax2d_gemm: act_cols / dact_cols must be multiples of 4");
  return AX2D_OK;
}

}  // namespace ax2d

using namespace ax2d;

extern "C" int64_t ax2d_gemm_workspace(int64_t M, int64_t N, int64_t K, int trans_a, int split_k) {
  (void)K;
  (void)trans_a;
  return split_k > 1 ? static_cast<int64_t>(split_k) * M * N * 4 : 0;
}

extern "C" int ax2d_gemm(const ax2d_cmat* a, int trans_a, const ax2d_cmat* b, int trans_b, const ax2d_mat* c, int64_t M,
                         int64_t N, int64_t K, const ax2d_epilogue* ep, int split_k, void* workspace,
                         ax2d_stream_t stream) {
  AX2D_CHECK_ARG(a != nullptr && b != nullptr && c != nullptr, "ax2d_gemm: null operand");
  AX2D_CHECK_ARG(M >= 0 && N > 0 && K >= 0 && N % 4 == 0, "ax2d_gemm: bad sizes M=%lld N=%lld K=%lld", (long long)M,
                 (long long)N, (long long)K);
  if (M == 0) return AX2D_OK;
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  int rc;
  // A: segments run along K (trans_a == 0) or along M (trans_a == 1)
  if ((rc = to_view(a, &g.a, trans_a ? M : K, "A")) != AX2D_OK) return rc;
  if ((rc = to_view(b, &g.b, trans_b ? K : N, "B")) != AX2D_OK) return rc;
  if ((rc = to_out(c, &g.e.c, N, "C", false)) != AX2D_OK) return rc;
  if (!trans_a && g.a.n_seg > 1)
    for (int s = 0; s < g.a.n_seg; ++s)
      AX2D_CHECK_ARG(a->width[s] % BK == 0, "ax2d_gemm: K-segments of A must be multiples of %d", BK);
  if (trans_b && g.b.n_seg > 1)
    for (int s = 0; s < g.b.n_seg; ++s)
      AX2D_CHECK_ARG(b->width[s] % BK == 0, "ax2d_gemm: K-segments of B must be multiples of %d", BK);
  AX2D_CHECK_ARG(!trans_a || M % 4 == 0, "ax2d_gemm: trans_a needs M %% 4 == 0");
  AX2D_CHECK_ARG(K % 4 == 0 || (trans_a && !trans_b), "ax2d_gemm: K must be a multiple of 4 for k-contiguous operands");
  g.M = M; g.N = N; g.K = K;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((rc = fill_epilogue(ep, M, N, &g.e)) != AX2D_OK) return rc;
  dim3 grid(static_cast<unsigned>((N + BN - 1) / BN), static_cast<unsigned>((M + BM - 1) / BM), 1);
  if (split_k > 1) {
    AX2D_CHECK_ARG(workspace != nullptr, "ax2d_gemm: split_k needs a workspace");
    AX2D_CHECK_ARG(ep == nullptr || (ep->bias == nullptr && ep->act == AX2D_ACT_NONE && ep->mask == nullptr &&
                                     ep->drop_p == 0.f && ep->resid.n_seg == 0 && ep->dact_pre == nullptr &&
                                     ep->pre.n_seg == 0),
                   "ax2d_gemm: split_k supports only the accumulate epilogue");
    AX2D_CHECK_ALIGN(workspace);
    int64_t per = (K + split_k - 1) / split_k;
    per = (per + BK - 1) / BK * BK;
    g.k_begin_stride = per;
    g.ws = static_cast<float*>(workspace);
    grid.z = static_cast<unsigned>((K + per - 1) / per);
  }
  AX2D_CHECK_ARG(!trans_a || (g.e.act == AX2D_ACT_NONE && g.e.dact == AX2D_ACT_NONE && g.e.mask == nullptr &&
                              g.e.drop_p == 0.f),
                 "ax2d_gemm: trans_a (weight-gradient) products support only bias / residual / accumulate epilogues");
  if (trans_a && trans_b) launch_k(gemm_kernel<true, true>, dim3(grid), dim3(GEMM_THREADS), 0, st, g);
  else if (trans_a) launch_k(gemm_kernel<true, false>, dim3(grid), dim3(GEMM_THREADS), 0, st, g);
  else if (trans_b) launch_k(gemm_kernel<false, true>, dim3(grid), dim3(GEMM_THREADS), 0, st, g);
  else launch_k(gemm_kernel<false, false>, dim3(grid), dim3(GEMM_THREADS), 0, st, g);
  rc = launch_status("ax2d_gemm");
  if (rc != AX2D_OK) return rc;
  if (split_k > 1) {
    const int64_t total = 8 * M * (N / 4);
    launch_k(splitk_reduce_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 
        g.ws, static_cast<int>(grid.z), M, N, g.e.c, g.e.accumulate, nullptr, nullptr);
    rc = launch_status("ax2d_gemm(split-k reduce)");
  }
  return rc;
}

extern "C" int64_t ax2d_colsum_workspace(int64_t M, int64_t N) {
  return ((M + kColsumRows - 1) / kColsumRows) * N * 4;
}

extern "C" int ax2d_colsum(const ax2d_cmat* a, int64_t M, int64_t N, float* out, int accumulate, void* workspace,
                           ax2d_stream_t stream) {
  AX2D_CHECK_ARG(a != nullptr && out != nullptr && workspace != nullptr && N > 0 && N % 4 == 0, "ax2d_colsum: bad arguments");
  SegView v;
  int rc = to_view(a, &v, N, "colsum input");
  if (rc != AX2D_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int n_part = static_cast<int>((M + kColsumRows - 1) / kColsumRows);
  if (n_part > 0) {
    dim3 grid(static_cast<unsigned>((N / 4 + 31) / 32), static_cast<unsigned>(n_part));
    launch_k(colsum_partial_kernel, dim3(grid), dim3(256), 0, st, v, M, static_cast<int>(N / 4), static_cast<float*>(workspace));
  }
  launch_k(colsum_final_kernel, dim3(static_cast<unsigned>((N + 127) / 128)), dim3(128), 0, st, static_cast<const float*>(workspace), n_part,
                                                                               static_cast<int>(N), out, accumulate);
  return launch_status("ax2d_colsum", n_part > 0 ? 2 : 1);
}

extern "C" int ax2d_act_bwd(const float* gr, int64_t ldg, const float* pre, int64_t ldp, float* out, int64_t ldo,
                            int64_t M, int width, int act, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(width > 0 && width % 4 == 0 && ldg % 4 == 0 && ldp % 4 == 0 && ldo % 4 == 0, "ax2d_act_bwd: widths % 4");
  AX2D_CHECK_ALIGN(gr);
  AX2D_CHECK_ALIGN(pre);
  AX2D_CHECK_ALIGN(out);
  if (M <= 0) return AX2D_OK;
  const int64_t total = M * (width / 4);
  launch_k(act_bwd_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      gr, ldg, pre, ldp, out, ldo, M, width / 4, act);
  return launch_status("ax2d_act_bwd");
}

extern "C" int ax2d_tick(uint64_t* counter, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(counter != nullptr, "ax2d_tick: null counter");
  launch_k(tick_kernel, dim3(1), dim3(1), 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<unsigned long long*>(counter));
  return launch_status("ax2d_tick");
}
