// Shared pieces of the dense kernels: segmented operands and the fused epilogue.
#pragma once
#include "common.cuh"

namespace ax2d {

struct SegView {          // column-segmented read-only matrix
  const float* ptr[AX2D_MAX_SEG];
  int64_t ld[AX2D_MAX_SEG];
  int start[AX2D_MAX_SEG + 1];
  int n_seg;
};
struct SegOut {
  float* ptr[AX2D_MAX_SEG];
  int64_t ld[AX2D_MAX_SEG];
  int start[AX2D_MAX_SEG + 1];
  int n_seg;
};

// Everything the fused epilogue needs (shared by the SIMT and the tcgen05 kernels).
struct EpiArgs {
  SegOut c, pre;
  int64_t M, N;
  const float* bias;
  int act, act_cols;
  const float* mask; int64_t ld_mask;
  float drop_p; uint64_t drop_seed; const uint64_t* drop_tick;
  const float* resid[AX2D_MAX_SEG]; int64_t ld_resid[AX2D_MAX_SEG]; int resid_cols[AX2D_MAX_SEG]; int n_resid;
  const float* dact_pre; int64_t ld_dact; int dact; int dact_cols;
  int accumulate;
};

__device__ __forceinline__ int find_seg(const int* start, int n_seg, int col) {
  int s = 0;
#pragma unroll
  for (int i = 1; i < AX2D_MAX_SEG; ++i)
    if (i < n_seg && col >= start[i]) s = i;
  return s;
}

// ------------------------------------------------------------------------------------------------ fused epilogue
// Per float4 of accumulators at (row m, columns n..n+3), m < M, n < N (N % 4 == 0), in this order:
//   v = acc + bias;  pre = v;  v = act(v) (n < act_cols);  v *= dropout;  v += residuals;
//   v *= act'(dact_pre) * dropout  (backward of the two steps above);  c (+)= v.
// act_cols, dact_cols and residual widths are multiples of 4 (checked on the host), so a float4 never straddles.
// The code is templated on the activation kinds and on dropout so that only the live path is compiled into each
// instance (a runtime switch per element costs ~1400 issue slots per float4 once if-converted), and split into a
// per-column setup (EpiCol, once per thread and column group), a LOAD half and a COMPUTE/STORE half so that
// callers with few resident warps can issue the loads of several rows before consuming any.

// sigmoid through the two special-function approximations (ex2.approx: 2^-22 relative, rcp.approx: 1 ulp): ~1e-7
// relative, two orders below the 1e-5 parity tolerance, and 5 instructions instead of the ~25 of expf + IEEE
// reciprocal -- the fused epilogues are bound by their instruction count
__device__ __forceinline__ float fast_sigmoid(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * v));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
template <int ACT>
__device__ __forceinline__ float act_fwd_t(float v) {
  if constexpr (ACT == AX2D_ACT_RELU) return v > 0.f ? v : 0.f;
  else if constexpr (ACT == AX2D_ACT_LEAKYRELU) return v > 0.f ? v : 0.01f * v;
  else if constexpr (ACT == AX2D_ACT_ELU) return v > 0.f ? v : expm1f(v);
  else if constexpr (ACT == AX2D_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
  else if constexpr (ACT == AX2D_ACT_SILU) return v * fast_sigmoid(v);
  else return v;
}
template <int ACT>
__device__ __forceinline__ float act_bwd_t(float v) {
  if constexpr (ACT == AX2D_ACT_RELU) return v > 0.f ? 1.f : 0.f;
  else if constexpr (ACT == AX2D_ACT_LEAKYRELU) return v > 0.f ? 1.f : 0.01f;
  else if constexpr (ACT == AX2D_ACT_ELU) return v > 0.f ? 1.f : expf(v);
  else if constexpr (ACT == AX2D_ACT_GELU) {
    const float cdf = 0.5f * (1.f + erff(v * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * v * v);
    return cdf + v * pdf;
  } else if constexpr (ACT == AX2D_ACT_SILU) {
    const float sg = fast_sigmoid(v);
    return sg * (1.f + v * (1.f - sg));
  } else return 1.f;
}

// Counter-based dropout: the keep decision of element (m, n) is a pure function of (seed, m * N + n), so the
// backward pass regenerates the forward mask instead of storing it.  One 32-bit avalanche hash (murmur3 finaliser)
// per GROUP OF FOUR elements and one more odd multiply of it give 4 x 16 bits: keep iff bits >= round(p * 65536).
struct EpiCtx {      // per-thread constants of the epilogue
  bool dropping;
  float inv_keep;
  uint32_t seed_lo, seed_hi, thresh;
  uint32_t base0;    // seed mix for linear indices below 2^34 (every realistic matrix)
  float acc_scale;   // 1, or the truncation compensation of a tensor-core accumulator (gemm_tc.cu: tc_acc_scale)
};
__device__ __forceinline__ EpiCtx epi_ctx_seed(float drop_p, uint64_t drop_seed, const uint64_t* drop_tick, bool has_mask = false) {
  EpiCtx c;
  c.acc_scale = 1.f;
  c.dropping = has_mask || drop_p > 0.f;
  c.thresh = static_cast<uint32_t>(drop_p * 65536.f + 0.5f);
  // the keep probability that is actually realised is (65536 - thresh) / 65536 (p quantised to 16 bits): scale by ITS
  // reciprocal so that E[mask] == 1 exactly
  c.inv_keep = drop_p > 0.f ? 65536.f / static_cast<float>(65536u - (c.thresh < 65535u ? c.thresh : 65535u)) : 1.f;
  const uint64_t seed = drop_seed + (drop_tick != nullptr
                                         ? __ldg(reinterpret_cast<const unsigned long long*>(drop_tick)) * 0xD1B54A32D192ED03ull
                                         : 0ull);
  c.seed_lo = static_cast<uint32_t>(seed);
  c.seed_hi = static_cast<uint32_t>(seed >> 32);
  uint32_t h = c.seed_hi;
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  c.base0 = h ^ c.seed_lo;
  return c;
}
__device__ __forceinline__ EpiCtx epi_ctx(const EpiArgs& g) { return epi_ctx_seed(g.drop_p, g.drop_seed, g.drop_tick, g.mask != nullptr); }
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// keep-scales of the four elements starting at linear index idx (idx % 4 == 0)
__device__ __forceinline__ void drop_scale4(const EpiCtx& cx, uint64_t idx, float d[4]) {
  const uint32_t lo = static_cast<uint32_t>(idx >> 2), hi = static_cast<uint32_t>(idx >> 34);
  const uint32_t base = hi == 0u ? cx.base0 : (mix32(hi ^ cx.seed_hi) ^ cx.seed_lo);
  const uint32_t h0 = mix32((lo * 0x9E3779B1u) ^ base);
  // second 32 bits: a multiply-xorshift round of h0 (not an affine map of it: under h0 * odd + c the low half of h1 would
  // be a function of the low half of h0 alone and the keep decisions of elements 0 / 2 and 1 / 3 would be coupled)
  uint32_t h1 = (h0 ^ (h0 >> 15)) * 0x2C1B3C6Du;
  h1 ^= h1 >> 12;
  d[0] = (h0 & 0xFFFFu) >= cx.thresh ? cx.inv_keep : 0.f;
  d[1] = (h0 >> 16) >= cx.thresh ? cx.inv_keep : 0.f;
  d[2] = (h1 & 0xFFFFu) >= cx.thresh ? cx.inv_keep : 0.f;
  d[3] = (h1 >> 16) >= cx.thresh ? cx.inv_keep : 0.f;
}

constexpr int kEpiPrefetchResid = 2;
struct EpiCol {      // per-thread constants of one float4 column group n..n+3
  float* c; int64_t ldc;
  float* pre; int64_t ldpre;                                          // nullptr: no pre-activation copy here
  const float* resid[kEpiPrefetchResid]; int64_t ldr[kEpiPrefetchResid];   // nullptr: residual does not cover n
  const float* dpre; int64_t lddp;                                    // nullptr: act' not applied at n
  const float* mask; int64_t ldm;
  float4 bias;
  bool act_on;
  bool accumulate;
  bool more_resid;     // residuals beyond the prefetched ones exist (rare; read through g)
  int n;
  uint32_t ncols;      // N: row pitch of the dropout counter
};
__device__ __forceinline__ EpiCol epi_col(const EpiArgs& g, int n) {
  EpiCol k;
  const int cs = find_seg(g.c.start, g.c.n_seg, n);
  k.c = g.c.ptr[cs] + (n - g.c.start[cs]);
  k.ldc = g.c.ld[cs];
  k.pre = nullptr; k.ldpre = 0;
  if (g.pre.n_seg > 0) {
    const int ps = find_seg(g.pre.start, g.pre.n_seg, n);
    if (g.pre.ptr[ps] != nullptr) { k.pre = g.pre.ptr[ps] + (n - g.pre.start[ps]); k.ldpre = g.pre.ld[ps]; }
  }
#pragma unroll
  for (int r = 0; r < kEpiPrefetchResid; ++r) {
    const bool on = r < g.n_resid && n < g.resid_cols[r];
    k.resid[r] = on ? g.resid[r] + n : nullptr;
    k.ldr[r] = on ? g.ld_resid[r] : 0;
  }
  const bool d_on = g.dact != AX2D_ACT_NONE && n < g.dact_cols;
  k.dpre = d_on ? g.dact_pre + n : nullptr;
  k.lddp = d_on ? g.ld_dact : 0;
  k.mask = g.mask != nullptr ? g.mask + n : nullptr;
  k.ldm = g.ld_mask;
  k.bias = g.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(g.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
  k.act_on = n < g.act_cols;
  k.accumulate = g.accumulate != 0;
  k.more_resid = g.n_resid > kEpiPrefetchResid;
  k.n = n;
  k.ncols = static_cast<uint32_t>(g.N);
  return k;
}
struct EpiOperands {       // the operands worth issuing early: the first two residuals and the saved pre-activation
  float4 resid[kEpiPrefetchResid];
  float4 dpre;
};
__device__ __forceinline__ void epi_prefetch(const EpiArgs& g, const EpiCol& k, int64_t m, EpiOperands& o) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < kEpiPrefetchResid; ++r)
    o.resid[r] = k.resid[r] != nullptr ? __ldg(reinterpret_cast<const float4*>(k.resid[r] + m * k.ldr[r])) : z;
  o.dpre = k.dpre != nullptr ? __ldg(reinterpret_cast<const float4*>(k.dpre + m * k.lddp)) : z;
}
// `m` addresses rows relative to the pointers in `k`; `row` (default m) is the absolute row, used for the dropout
// counter and the rarely used operands that are read through `g`.
template <int ACT, int DACT, bool DROP>
__device__ __forceinline__ void epi_finish(const EpiArgs& g, const EpiCtx& cx, const EpiCol& k, int64_t m,
                                           const EpiOperands& o, float a0, float a1, float a2, float a3,
                                           int64_t row = -1) {
  if (row < 0) row = m;
  // (fma with acc_scale == 1 is the plain sum, bit for bit)
  float v[4] = {fmaf(a0, cx.acc_scale, k.bias.x), fmaf(a1, cx.acc_scale, k.bias.y), fmaf(a2, cx.acc_scale, k.bias.z),
                fmaf(a3, cx.acc_scale, k.bias.w)};
  if (k.pre != nullptr) *reinterpret_cast<float4*>(k.pre + m * k.ldpre) = make_float4(v[0], v[1], v[2], v[3]);
  float drop[4] = {1.f, 1.f, 1.f, 1.f};
  if constexpr (DROP) {
    if (k.mask != nullptr) {
      const float4 mk = __ldg(reinterpret_cast<const float4*>(k.mask + m * k.ldm));
      drop[0] = mk.x; drop[1] = mk.y; drop[2] = mk.z; drop[3] = mk.w;
    } else {
      drop_scale4(cx, static_cast<uint64_t>(row) * k.ncols + static_cast<uint64_t>(k.n), drop);
    }
  }
  if constexpr (ACT != AX2D_ACT_NONE) {
    if (k.act_on) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = act_fwd_t<ACT>(v[j]);
    }
  }
  if constexpr (DROP && DACT == AX2D_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= drop[j];
  }
#pragma unroll
  for (int r = 0; r < kEpiPrefetchResid; ++r) {     // zeros where the residual does not apply
    v[0] += o.resid[r].x; v[1] += o.resid[r].y; v[2] += o.resid[r].z; v[3] += o.resid[r].w;
  }
  if (k.more_resid) {
    for (int r = kEpiPrefetchResid; r < g.n_resid; ++r) {
      if (k.n >= g.resid_cols[r]) continue;
      const float4 rv = __ldg(reinterpret_cast<const float4*>(g.resid[r] + row * g.ld_resid[r] + k.n));
      v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
    }
  }
  if constexpr (DACT != AX2D_ACT_NONE) {
    if (k.dpre != nullptr) {
      v[0] *= act_bwd_t<DACT>(o.dpre.x) * drop[0];
      v[1] *= act_bwd_t<DACT>(o.dpre.y) * drop[1];
      v[2] *= act_bwd_t<DACT>(o.dpre.z) * drop[2];
      v[3] *= act_bwd_t<DACT>(o.dpre.w) * drop[3];
    }
  }
  float4* dst = reinterpret_cast<float4*>(k.c + m * k.ldc);
  if (k.accumulate) {
    const float4 cold = *dst;
    v[0] += cold.x; v[1] += cold.y; v[2] += cold.z; v[3] += cold.w;
  }
  *dst = make_float4(v[0], v[1], v[2], v[3]);
}

// Calls f(std::integral_constant<int, ACT>, std::integral_constant<int, DACT>, std::bool_constant<DROP>) for the
// instance matching the runtime epilogue (forward: DACT == NONE; backward: ACT == NONE).
#define AX2D_EPI_CASE(A, D, ...)                                                    \
  if (dropping) { constexpr int ACT = A; constexpr int DACT = D; constexpr bool DROP = true; __VA_ARGS__ } \
  else { constexpr int ACT = A; constexpr int DACT = D; constexpr bool DROP = false; __VA_ARGS__ }
#define AX2D_EPI_DISPATCH(act, dact, dropping, ...)                                 \
  do {                                                                              \
    if ((dact) == AX2D_ACT_NONE) {                                                  \
      switch (act) {                                                                \
        case AX2D_ACT_RELU: { AX2D_EPI_CASE(AX2D_ACT_RELU, AX2D_ACT_NONE, __VA_ARGS__) } break;           \
        case AX2D_ACT_LEAKYRELU: { AX2D_EPI_CASE(AX2D_ACT_LEAKYRELU, AX2D_ACT_NONE, __VA_ARGS__) } break; \
        case AX2D_ACT_ELU: { AX2D_EPI_CASE(AX2D_ACT_ELU, AX2D_ACT_NONE, __VA_ARGS__) } break;             \
        case AX2D_ACT_GELU: { AX2D_EPI_CASE(AX2D_ACT_GELU, AX2D_ACT_NONE, __VA_ARGS__) } break;           \
        case AX2D_ACT_SILU: { AX2D_EPI_CASE(AX2D_ACT_SILU, AX2D_ACT_NONE, __VA_ARGS__) } break;           \
        default: { AX2D_EPI_CASE(AX2D_ACT_NONE, AX2D_ACT_NONE, __VA_ARGS__) } break;                      \
      }                                                                             \
    } else {                                                                        \
      switch (dact) {                                                               \
        case AX2D_ACT_RELU: { AX2D_EPI_CASE(AX2D_ACT_NONE, AX2D_ACT_RELU, __VA_ARGS__) } break;           \
        case AX2D_ACT_LEAKYRELU: { AX2D_EPI_CASE(AX2D_ACT_NONE, AX2D_ACT_LEAKYRELU, __VA_ARGS__) } break; \
        case AX2D_ACT_ELU: { AX2D_EPI_CASE(AX2D_ACT_NONE, AX2D_ACT_ELU, __VA_ARGS__) } break;             \
        case AX2D_ACT_GELU: { AX2D_EPI_CASE(AX2D_ACT_NONE, AX2D_ACT_GELU, __VA_ARGS__) } break;           \
        default: { AX2D_EPI_CASE(AX2D_ACT_NONE, AX2D_ACT_SILU, __VA_ARGS__) } break;                      \
      }                                                                             \
    }                                                                               \
  } while (0)

// host helpers (gemm.cu)
int to_view(const ax2d_cmat* m, SegView* v, int64_t total, const char* what);
int to_out(const ax2d_mat* m, SegOut* v, int64_t total, const char* what, bool allow_null);
int fill_epilogue(const ax2d_epilogue* ep, int64_t M, int64_t N, EpiArgs* e);
int splitk_reduce(const float* ws, int split, int64_t M, int64_t N, const SegOut& c, int accumulate, const float* vec_ws,
                  float* vec_out, cudaStream_t st);   // validates + copies; c is set by the caller

}  // namespace ax2d
