// bf16 dense projections on the 5th-generation tensor cores (tcgen05 kind::f16 / TMEM / TMA) -- the mixed-precision
// configuration of BASELINE.json (configs[3]: bf16, fp32 accumulation, fp32 master weights; tolerance 2e-2).
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )     A: bf16 activations, column-segmented (cat is never materialised)
//                                               B: bf16 weights, K-major (the packed cache keeps W and W^T in bf16)
//   dW[No,Ki] = sum_m G[m,o] X[m,i]            G, X: bf16 activations, contraction over the rows (weight gradients)
//
// Unlike the fp32-faithful kernels of gemm_tc.cu there is no operand split: BOTH operands go from global memory to shared
// memory by TMA (128-byte swizzle) and from there straight into tcgen05.mma -- no splitter warps, no TMEM A ring.  With the
// tensor work gone (one kind::f16 MMA per product at twice the TF32 rate instead of three) these kernels are bound by
// their output traffic, so the epilogue is built around the TMA engine in BOTH directions:
//   * every epilogue warp owns one TMEM lane quarter (32 rows) and walks 32-column chunks: tcgen05.ld (lane = row) ->
//     bias / activation / counter-based dropout / residuals / activation-backward in registers -> the finished chunk is
//     written to a swizzled shared-memory staging tile and leaves as ONE bulk tensor store (cp.async.bulk.tensor ...
//     global.shared::cta) -- no per-thread global addressing, no transpose pass, full 32-byte sectors;
//   * the chunk's input operands (residuals, saved pre-activation) arrive the same way: bulk tensor loads into staging
//     tiles, issued one chunk ahead (also across tile boundaries) and awaited on a per-warp mbarrier.
// Persistent, one CTA per SM: warp 0 = TMA producer of the k-block ring, warp 1 = MMA issuer + TMEM owner (accumulators
// double-buffered: the epilogue of tile i overlaps the main loop of tile i + 1), warps 2..9 = epilogue.
#include <cuda_bf16.h>

#include "tc_ptx.cuh"

namespace ax2d {

constexpr int BF_BM = 128;
constexpr int BF_BK = 64;                 // bf16 per 128-byte swizzle row
constexpr int BF_MAX_BN = 256;            // two accumulator buffers share the 512 TMEM columns
constexpr int BF_EPI_WARPS = 12;          // three per TMEM lane quarter
constexpr int BF_THREADS = 32 * (2 + BF_EPI_WARPS);
constexpr int BF_MAX_STAGES = 8;
constexpr int BF_MAX_CT = 32;             // column tiles per row tile
constexpr int BF_MAX_CSEG = 4;            // output / pre-activation segments
constexpr int BF_MAX_RES = 3;
constexpr int BF_CHUNK = 32;              // columns per epilogue chunk (= one tcgen05.ld 32x32b.x32)
constexpr int BF_BIAS_SMEM = 2048;        // bias entries staged in shared memory

struct BfMaps {
  CUtensorMap a[AX2D_MAX_SEG];
  CUtensorMap b;
  CUtensorMap c[BF_MAX_CSEG];
  CUtensorMap pre[BF_MAX_CSEG];
  CUtensorMap res[BF_MAX_RES];
  CUtensorMap dp;
};
struct BfColTile {
  int n0, w;             // absolute first column, valid width
  int cseg, cn0;         // output segment and column inside it
  int pseg, pn0;         // pre-activation segment (-1: none) and column inside it
  uint32_t flags;        // bit 0: activation, bit 1: activation-backward, bits 2..4: residual r covers the tile
};
struct BfArgs {
  unsigned long long* dbg;               // development aid: %globaltimer stamps of CTA 0 (ax2d_debug_bf16_timing); normally null
  int64_t M;
  int N;
  const float* bias;
  int act, dact;
  float drop_p;
  uint64_t drop_seed;
  const uint64_t* drop_tick;
  int c_f32;                             // output dtype: 1 = fp32, 0 = bf16
  int n_seg, num_kb;
  int seg_kb_start[AX2D_MAX_SEG + 1];    // first k-block of every A segment
  int seg_k0[AX2D_MAX_SEG];              // first K column of every A segment inside B
  int n_ct, total_tiles, stages;
  int bn_box;                            // rows of the B box (>= every tile's MMA N)
  int in_double;                         // 1: the input staging tiles are double-buffered (next chunk's loads fly during this one)
  int n_in_max;                          // staging tiles per chunk: inputs (residuals + dact_pre) ...
  int n_out;                             // ... and outputs (c, pre)
  BfColTile ct[BF_MAX_CT];
};

// ---------------------------------------------------------------------------------------------- PTX (bf16 specific)
// One k-block = four kind::f16 k-steps (K = 16 each) with both operands in shared memory, and the commit that frees the
// stage, under one election.  `step`: descriptor start-address advance per k-step in 16-byte units (K-major 128-byte
// swizzle: 32 bytes = 2; MN-major 64-byte swizzle: two 8-row groups of 512 bytes = 64).
__device__ __forceinline__ void umma_kblock_ss_w(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc_first,
                                                 uint64_t step, uint64_t* free_bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e, pf, pt;\n\t"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %4, 0;\n\t"
      "setp.eq.b32 pt, %3, %3;\n\t"
      "add.u64 a1, %1, %5;\n\t add.u64 a2, a1, %5;\n\t add.u64 a3, a2, %5;\n\t"
      "add.u64 b1, %2, %5;\n\t add.u64 b2, b1, %5;\n\t add.u64 b3, b2, %5;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pf;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, pt;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "}\n" ::"r"(d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc_first), "l"(step), "r"(smem_u32(free_bar))
      : "memory");
}
// the same plus a fifth chain: the N = 16 "ones" operand whose accumulator column 0 is the column sum of A (bias gradient)
__device__ __forceinline__ void umma_kblock_ss_ones_w(uint32_t d, uint32_t d_ones, uint64_t da, uint64_t db, uint64_t dones,
                                                      uint32_t idesc, uint32_t idesc_ones, uint32_t acc_first, uint64_t step,
                                                      uint64_t* free_bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e, pf, pt;\n\t"
      ".reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %4, 0;\n\t"
      "setp.eq.b32 pt, %3, %3;\n\t"
      "add.u64 a1, %1, %5;\n\t add.u64 a2, a1, %5;\n\t add.u64 a3, a2, %5;\n\t"
      "add.u64 b1, %2, %5;\n\t add.u64 b2, b1, %5;\n\t add.u64 b3, b2, %5;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pf;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, pt;\n\t"
      // every k-step reads the same 16 rows of ones (the tile is constant), so the descriptor does not advance
      "@e tcgen05.mma.cta_group::1.kind::f16 [%7], %1, %8, %9, pf;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%7], a1, %8, %9, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%7], a2, %8, %9, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%7], a3, %8, %9, pt;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
      "}\n" ::"r"(d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc_first), "l"(step), "r"(smem_u32(free_bar)), "r"(d_ones), "l"(dones),
      "r"(idesc_ones)
      : "memory");
}
// expect_tx + the two tensor-map loads of one k-block (A, B), one election
__device__ __forceinline__ void tma_kblock_ab_w(uint64_t* bar, uint32_t bytes, void* dst_a, const CUtensorMap* map_a, int ka, int m0,
                                                void* dst_b, const CUtensorMap* map_b, int kb, int n0) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3, {%4, %5}], [%0];\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%6], [%7, {%8, %9}], [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(bytes), "r"(smem_u32(dst_a)), "l"(reinterpret_cast<uint64_t>(map_a)), "r"(ka), "r"(m0), "r"(smem_u32(dst_b)),
      "l"(reinterpret_cast<uint64_t>(map_b)), "r"(kb), "r"(n0)
      : "memory");
}
// single-lane forms (called by lane 0 of an epilogue warp)
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   dst_smem),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int c_inner, int c_outer) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(src_smem), "r"(c_inner), "r"(c_outer)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// Instruction descriptor, kind::f16: D = F32 (bit 4), A = B = BF16 (format 1 at bits 7 and 10), M = 128, N = bn;
// bits 15 / 16: A / B are MN-major (weight gradients).
__device__ __forceinline__ uint32_t idesc_bf16(int bn, bool mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | (static_cast<uint32_t>(bn >> 3) << 17) |
         (static_cast<uint32_t>(BF_BM >> 4) << 24);
}
// MN-major operand, 64-byte swizzle (cute: Layout_MN_SW64_Atom<bf16> = Swizzle<2,4,3> o ((32,n),(8,k)):((1,LBO),(32,SBO))
// in elements): a TMA box [32 columns x KB rows] (64-byte rows) is one column chunk; the next 32-column chunk follows
// `chunk_bytes` later (LBO), the next group of 8 contraction rows 512 bytes later (SBO).  Layout type 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t smem_desc_mn_sw64(uint32_t smem_addr, uint32_t chunk_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(chunk_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------- epilogue of one chunk
// Lane = accumulator row.  v[0..31]: the chunk's 32 columns of that row.  Staging tiles: bf16 [32 rows x 64 B], the 16-byte
// unit u of row r lives at u ^ ((r >> 1) & 3) (TMA SWIZZLE_64B); fp32 [32 rows x 128 B], unit u at u ^ (r & 7) (SWIZZLE_128B).
// Both patterns make a warp's 128-bit accesses conflict-free (8 consecutive rows cover all 32 banks).
__device__ __forceinline__ void stage_read_bf16(uint32_t tile, int lane, float* v) {
  const uint32_t row = tile + static_cast<uint32_t>(lane) * 64u;
  const uint32_t sw = (static_cast<uint32_t>(lane) >> 1) & 3u;
#pragma unroll
  for (uint32_t u = 0; u < 4; ++u) {
    const uint4 q = lds128u(row + ((u ^ sw) << 4));
    v[8 * u + 0] = bf_lo(q.x); v[8 * u + 1] = bf_hi(q.x); v[8 * u + 2] = bf_lo(q.y); v[8 * u + 3] = bf_hi(q.y);
    v[8 * u + 4] = bf_lo(q.z); v[8 * u + 5] = bf_hi(q.z); v[8 * u + 6] = bf_lo(q.w); v[8 * u + 7] = bf_hi(q.w);
  }
}
__device__ __forceinline__ void stage_write_bf16(uint32_t tile, int lane, const float* v) {
  const uint32_t row = tile + static_cast<uint32_t>(lane) * 64u;
  const uint32_t sw = (static_cast<uint32_t>(lane) >> 1) & 3u;
#pragma unroll
  for (uint32_t u = 0; u < 4; ++u)
    sts128u(row + ((u ^ sw) << 4), pack_bf16(v[8 * u], v[8 * u + 1]), pack_bf16(v[8 * u + 2], v[8 * u + 3]),
            pack_bf16(v[8 * u + 4], v[8 * u + 5]), pack_bf16(v[8 * u + 6], v[8 * u + 7]));
}
__device__ __forceinline__ void stage_write_f32(uint32_t tile, int lane, const float* v) {
  const uint32_t row = tile + static_cast<uint32_t>(lane) * 128u;
  const uint32_t sw = static_cast<uint32_t>(lane) & 7u;
#pragma unroll
  for (uint32_t u = 0; u < 8; ++u)
    sts128u(row + ((u ^ sw) << 4), __float_as_uint(v[4 * u]), __float_as_uint(v[4 * u + 1]), __float_as_uint(v[4 * u + 2]),
            __float_as_uint(v[4 * u + 3]));
}

constexpr uint32_t BF_IN_TILE = 2048;     // bytes of a bf16 staging tile

// SiLU and its derivative through ONE special-function instruction (tanh.approx, ~2^-11 relative -- four times finer than
// the bf16 rounding of the result): sigmoid(x) = 0.5 + 0.5 tanh(0.5 x).  The fp32 epilogues use ex2 + rcp (two MUFU ops per
// element); in this kernel the 12 epilogue warps are throughput-bound on the MUFU and integer pipes.
__device__ __forceinline__ float bf_sigmoid(float v) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * v));
  return fmaf(0.5f, t, 0.5f);
}
template <int ACT>
__device__ __forceinline__ float bf_act_fwd(float v) {
  if constexpr (ACT == AX2D_ACT_SILU) return v * bf_sigmoid(v);
  else return act_fwd_t<ACT>(v);
}
template <int ACT>
__device__ __forceinline__ float bf_act_bwd(float v) {
  if constexpr (ACT == AX2D_ACT_SILU) {
    const float sg = bf_sigmoid(v);
    return sg * fmaf(v, 1.f - sg, 1.f);
  } else return act_bwd_t<ACT>(v);
}

// v <- epilogue(v) for one chunk; `in_base`: this chunk's input staging tiles in the order residuals.., dact_pre.
// Returns with v = final values; pre-activation values (before activation) are written to pre_tile when has_pre.
template <int ACT, int DACT, bool DROP>
__device__ __forceinline__ void bf_chunk_math(const BfArgs& g, const EpiCtx& cx, const BfColTile& t, int col0, int64_t row, int lane,
                                              uint32_t in_base, uint32_t pre_tile, bool has_pre, float* v, const float* s_bias) {
  // bias (fp32, same 32 values for every lane): broadcast reads of the copy staged in shared memory at kernel start (a
  // global load here cost every chunk an exposed L2 round trip right at the top of its dependency chain)
  if (g.bias != nullptr) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int n = col0 + 4 * u;
      if (n + 3 < g.N) {
        const float4 b4 = s_bias != nullptr ? *reinterpret_cast<const float4*>(s_bias + n)
                                            : __ldg(reinterpret_cast<const float4*>(g.bias + n));
        v[4 * u] += b4.x; v[4 * u + 1] += b4.y; v[4 * u + 2] += b4.z; v[4 * u + 3] += b4.w;
      }
    }
  }
  if (has_pre) stage_write_bf16(pre_tile, lane, v);
  float drop[32];
  if constexpr (DROP) {
    // Counter-based like drop_scale4 (the backward pass regenerates the forward decisions from (seed, row, column)), but
    // one avalanche hash per (row, 32-column chunk) and then one Weyl-step / multiply / xorshift word per TWO elements:
    // ~5 instead of ~9 instructions per element in an epilogue that is bound by its instruction stream.
    const uint32_t h0 = mix32(mix32(static_cast<uint32_t>(row) ^ cx.base0) + static_cast<uint32_t>(col0) * 0x9E3779B1u);
    const uint32_t t16 = cx.thresh << 16;          // decisions compare the HIGH 16 bits of a word: no field extraction
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      uint32_t w = h0 + static_cast<uint32_t>(u) * 0x9E3779B9u;      // Weyl step (compile-time addend)
      w ^= w >> 15;
      w *= 0x2C1B3C6Du;
      w ^= w >> 12;
      drop[2 * u] = w >= t16 ? cx.inv_keep : 0.f;
      drop[2 * u + 1] = (w << 16) >= t16 ? cx.inv_keep : 0.f;
    }
  }
  if constexpr (ACT != AX2D_ACT_NONE) {
    if (t.flags & 1u) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = bf_act_fwd<ACT>(v[j]);
    }
  }
  if constexpr (DROP && DACT == AX2D_ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= drop[j];
  }
  uint32_t in_tile = in_base;
#pragma unroll
  for (int r = 0; r < BF_MAX_RES; ++r) {
    if (t.flags & (4u << r)) {
      float x[32];
      stage_read_bf16(in_tile, lane, x);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += x[j];
      in_tile += BF_IN_TILE;
    }
  }
  if constexpr (DACT != AX2D_ACT_NONE) {
    if (t.flags & 2u) {
      float x[32];
      stage_read_bf16(in_tile, lane, x);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float s = bf_act_bwd<DACT>(x[j]);
        if constexpr (DROP) s *= drop[j];
        v[j] *= s;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- projection kernel
__global__ void __launch_bounds__(BF_THREADS, 1) gemm_bf16_kernel(const __grid_constant__ BfMaps maps, const __grid_constant__ BfArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[BF_MAX_STAGES], empty_bar[BF_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ __align__(8) uint64_t in_bar[BF_EPI_WARPS][2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = g.stages;
  unsigned long long* dbg = (g.dbg != nullptr && blockIdx.x == 0) ? g.dbg : nullptr;
  if (dbg != nullptr && threadIdx.x == 0) dbg[0] = gtime();
  constexpr uint32_t A_BYTES = BF_BM * BF_BK * 2;                                   // 16 KB
  const uint32_t b_bytes = static_cast<uint32_t>(g.bn_box) * BF_BK * 2;
  const uint32_t stage_bytes = A_BYTES + b_bytes;                                   // multiple of 1024 (bn_box % 16 == 0 -> 2 KB units)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t out_tile_bytes = g.c_f32 ? 4096u : 2048u;
  const uint32_t in_bufs = g.in_double ? 2u : 1u;
  const uint32_t warp_stage_bytes = in_bufs * static_cast<uint32_t>(g.n_in_max) * BF_IN_TILE + 2u * static_cast<uint32_t>(g.n_out) * out_tile_bytes;
  unsigned char* epi_base = smem + static_cast<size_t>(S) * stage_bytes;
  // bias copy behind the epilogue staging tiles (N <= BF_BIAS_SMEM floats; larger N reads it from global memory)
  float* s_bias = (g.bias != nullptr && g.N <= BF_BIAS_SMEM)
                      ? reinterpret_cast<float*>(epi_base + static_cast<size_t>(BF_EPI_WARPS) * warp_stage_bytes) : nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], BF_EPI_WARPS);
    }
    for (int w = 0; w < BF_EPI_WARPS; ++w) {
      mbar_init(&in_bar[w][0], 1);
      mbar_init(&in_bar[w][1], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();      // nothing above touches global memory: the prologue overlaps the previous kernel's tail
  if (s_bias != nullptr) {                       // CTA-uniform
    for (int i = threadIdx.x; i < g.N; i += blockDim.x) s_bias[i] = __ldg(g.bias + i);
    __syncthreads();
  }
  const uint32_t tmem_base = tmem_base_smem;
  const int n_ct = g.n_ct, total = g.total_tiles, num_kb = g.num_kb;
  if (dbg != nullptr && threadIdx.x == 0) dbg[1] = gtime();

  if (warp == 0) {
    // ===================================================================== TMA producer (whole warp, elected issue)
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int m0 = (tile / n_ct) * BF_BM;
      const int n0 = g.ct[tile % n_ct].n0;
      int seg = 0;
      for (int kb = 0; kb < num_kb; ++kb, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
        mbar_wait(&empty_bar[s], ph ^ 1u);
        __syncwarp();
        while (seg + 1 < g.n_seg && kb >= g.seg_kb_start[seg + 1]) ++seg;
        const int ka = (kb - g.seg_kb_start[seg]) * BF_BK;          // column inside the A segment (the tail of a segment
        unsigned char* st = smem + static_cast<size_t>(s) * stage_bytes;   // reads zeros: out-of-bounds fill)
        tma_kblock_ab_w(&full_bar[s], stage_bytes, st, &maps.a[seg], ka, m0, st + A_BYTES, &maps.b, g.seg_k0[seg] + ka, n0);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp, elected issue)
    int s = 0, j = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++j) {
      const int buf = j & 1;
      const int w = g.ct[tile % n_ct].w;
      const uint32_t idesc = idesc_bf16((w + 15) & ~15, false);
      mbar_wait(&acc_empty[buf], (static_cast<uint32_t>(j >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(buf * BF_MAX_BN);
      for (int kb = 0; kb < num_kb; ++kb, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        __syncwarp();
        const uint32_t sa = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
        if (dbg != nullptr && lane == 0 && kb == 0 && j < 2) dbg[2 + j] = gtime();      // first k-block of tile j has landed
        umma_kblock_ss_w(d, smem_desc_k_sw128(sa), smem_desc_k_sw128(sa + A_BYTES), idesc, kb != 0 ? 1u : 0u, 2ull, &empty_bar[s]);
      }
      umma_commit_w(&acc_full[buf]);
      if (dbg != nullptr && lane == 0 && j < 2) dbg[4 + j] = gtime();                    // all MMAs of tile j issued
    }
  } else {
    // ===================================================================== epilogue warps
    const int ew = warp - 2;
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int cp = ew >> 2;                  // this warp takes chunks cp, cp + BF_EPI_WARPS / 4, ...
    const uint32_t stg = smem_u32(epi_base + static_cast<size_t>(ew) * warp_stage_bytes);
    const uint32_t in_bytes = static_cast<uint32_t>(g.n_in_max) * BF_IN_TILE;
    const uint32_t out_bytes = static_cast<uint32_t>(g.n_out) * out_tile_bytes;
    // staging of item k: inputs at stg + ib(k) * in_bytes, outputs at stg + in_bufs * in_bytes + (k & 1) * out_bytes,
    // ib(k) = k & 1 with double-buffered inputs, 0 otherwise
    const uint32_t dbl = g.in_double ? 1u : 0u;
    const EpiCtx cx = epi_ctx_seed(g.drop_p, g.drop_seed, g.drop_tick);
    const bool dropping = g.drop_p > 0.f;
    uint64_t* bars = in_bar[ew];

    // Work items of this warp: (tile, chunk) pairs in tile order, chunks cp, cp + BF_EPI_WARPS / 4, ...  A position keeps the
    // tile's row-tile / column-tile indices incrementally: no integer division in the loop (each costs ~50 instructions
    // on a warp whose instruction stream is the bottleneck).
    struct Pos { int tile, rt, ct, ci; };
    const int step_rt = static_cast<int>(gridDim.x) / n_ct, step_ct = static_cast<int>(gridDim.x) % n_ct;
    auto next_tile = [&](Pos& p) {
      p.tile += gridDim.x;
      p.rt += step_rt;
      p.ct += step_ct;
      if (p.ct >= n_ct) { p.ct -= n_ct; ++p.rt; }
    };
    auto first_item = [&](Pos& p) {          // first (tile, chunk) of this warp at or after p
      while (p.tile < total) {
        if (p.ci * BF_CHUNK < g.ct[p.ct].w) return;
        next_tile(p);
        p.ci = cp;
      }
    };
    // bulk loads of one chunk's input tiles (lane 0)
    auto issue_inputs = [&](const Pos& p, uint32_t k) {
      const BfColTile& t = g.ct[p.ct];
      const int row0 = p.rt * BF_BM + q * 32;
      const int col = t.n0 + p.ci * BF_CHUNK;
      uint32_t dst = stg + (k & dbl) * in_bytes;
      uint32_t n = 0;
      for (int r = 0; r < BF_MAX_RES; ++r) n += (t.flags >> (2 + r)) & 1u;
      n += (t.flags >> 1) & 1u;
      uint64_t* bar = &bars[k & dbl];
      if (n == 0) {                       // keep the barrier protocol uniform: a plain arrival completes the phase
        mbar_arrive(bar);
        return;
      }
      mbar_expect_tx(bar, n * BF_IN_TILE);
      for (int r = 0; r < BF_MAX_RES; ++r)
        if (t.flags & (4u << r)) {
          tma_load_2d(dst, &maps.res[r], bar, col, row0);
          dst += BF_IN_TILE;
        }
      if (t.flags & 2u) tma_load_2d(dst, &maps.dp, bar, col, row0);
    };

    const bool use_in = g.n_in_max > 0;
    Pos cur{static_cast<int>(blockIdx.x), static_cast<int>(blockIdx.x) / n_ct, static_cast<int>(blockIdx.x) % n_ct, cp};
    Pos tp = cur;                            // the tile loop's own position
    first_item(cur);
    uint32_t k = 0;
    if (use_in && cur.tile < total && lane == 0) issue_inputs(cur, 0);
    int j = 0;
    for (; tp.tile < total; next_tile(tp), ++j) {
      const int tile = tp.tile;
      const int buf = j & 1;
      const BfColTile& t = g.ct[tp.ct];
      const int m0 = tp.rt * BF_BM;
      mbar_wait(&acc_full[buf], static_cast<uint32_t>(j >> 1) & 1u);
      tc_fence_after();
      if (dbg != nullptr && ew == 0 && lane == 0 && j < 2) dbg[6 + 4 * j] = gtime();     // accumulator of tile j ready
      const uint32_t acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * BF_MAX_BN);
      const bool has_pre = t.pseg >= 0;
      while (cur.tile == tile) {
        const int ci = cur.ci;
        Pos nxt = cur;
        nxt.ci += BF_EPI_WARPS / 4;
        first_item(nxt);
        if (use_in) {
          if (dbl && nxt.tile < total && lane == 0) issue_inputs(nxt, k + 1);
          mbar_wait(&bars[k & dbl], dbl ? ((k >> 1) & 1u) : (k & 1u));
        }
        if (dbg != nullptr && ew == 0 && lane == 0 && j < 2 && ci == cp) dbg[7 + 4 * j] = gtime();   // inputs of the first chunk here
        uint32_t r[32];
        tmem_ld32(acc + static_cast<uint32_t>(ci * BF_CHUNK), r);
        if (dbg != nullptr && ew == 0 && lane == 0 && j < 2 && ci == cp) dbg[8 + 4 * j] = gtime();   // ... accumulators in registers
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        const uint32_t in_base = stg + (k & dbl) * in_bytes;
        const uint32_t out_base = stg + in_bufs * in_bytes + (k & 1u) * out_bytes;
        // the staging tiles of item k - 2 must have been read by their bulk stores before they are overwritten
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        const int64_t row = static_cast<int64_t>(m0) + q * 32 + lane;
        const int col0 = t.n0 + ci * BF_CHUNK;
        AX2D_EPI_DISPATCH(g.act, g.dact, dropping, {
          bf_chunk_math<ACT, DACT, DROP>(g, cx, t, col0, row, lane, in_base, out_base + out_tile_bytes, has_pre, v, s_bias);
        });
        if (g.c_f32) stage_write_f32(out_base, lane, v);
        else stage_write_bf16(out_base, lane, v);
        fence_proxy_async_smem();
        __syncwarp();
        // single-buffered inputs: every lane has consumed this chunk's tiles (syncwarp above), the next chunk's may land
        if (use_in && !dbl && nxt.tile < total && lane == 0) issue_inputs(nxt, k + 1);
        if (lane == 0) {
          const int row0 = m0 + q * 32;
          tma_store_2d(&maps.c[t.cseg], out_base, t.cn0 + ci * BF_CHUNK, row0);
          if (has_pre) tma_store_2d(&maps.pre[t.pseg], out_base + out_tile_bytes, t.pn0 + ci * BF_CHUNK, row0);
          bulk_commit();
          if (dbg != nullptr && ew == 0 && j < 2 && ci == cp) dbg[9 + 4 * j] = gtime();               // first chunk stored
        }
        cur = nxt;
        ++k;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    if (dbg != nullptr && ew == 0 && lane == 0) dbg[14] = gtime();                                    // last chunk of this warp issued
    if (lane == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (dbg != nullptr && threadIdx.x == 0) dbg[15] = gtime();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ---------------------------------------------------------------------------------------------- weight gradients
//   dW[No, Ki] = sum_m G[m, o] X[m, i]   (+ db[o] = sum_m G[m, o])
// Both operands are MN-major for the tensor core (the contraction index is the ROW of a row-major matrix): TMA boxes
// [32 columns x 64 rows] with the 64-byte swizzle land as the canonical MN-major SW64 layout, one box per 32-column chunk.
// CTA tile = 128 (o) x BN (i), the rows are split over blockIdx.z; every CTA leaves its fp32 partial tile in the
// workspace plane of its split (bulk tensor stores), summed in split order by ax2d_unpack_grads / splitk_reduce.
// The bias gradient rides along as one more N = 16 MMA per k-step against a constant tile of ones.
constexpr int BW_KB = 64;                 // contraction rows per k-block
constexpr int BW_CHUNK_BYTES = BW_KB * 64;        // one [32 columns x 64 rows] bf16 box
constexpr int BW_MAX_STAGES = 6;
struct BwMaps {
  CUtensorMap a[AX2D_MAX_SEG];
  CUtensorMap b[AX2D_MAX_SEG];
  CUtensorMap out;                        // fp32 [(splits * No) rows, Ki columns]
};
struct BwArgs {
  int No, Ki;
  int64_t rows;                           // contraction length
  int a_start[AX2D_MAX_SEG + 1], a_nseg;
  int b_start[AX2D_MAX_SEG + 1], b_nseg;
  int num_kb, kb_per_split;
  int BN, stages;
  float* db;                              // [splits][No] partial bias gradients or nullptr
};

__global__ void __launch_bounds__(BF_THREADS, 1) gemm_bf16_wgrad_kernel(const __grid_constant__ BwMaps maps, const __grid_constant__ BwArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[BW_MAX_STAGES], empty_bar[BW_MAX_STAGES], acc_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = g.BN, S = g.stages;
  const int a_chunks = BF_BM / 32, b_chunks = (BN + 31) / 32;
  const uint32_t a_bytes = static_cast<uint32_t>(a_chunks) * BW_CHUNK_BYTES, b_bytes = static_cast<uint32_t>(b_chunks) * BW_CHUNK_BYTES;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* ones = smem + static_cast<size_t>(S) * stage_bytes;                 // [64 rows x 64 B] of bf16 1.0
  unsigned char* epi_base = ones + BW_CHUNK_BYTES;                                    // 8 warps x 2 x 4 KB fp32 staging tiles

  const int m0 = blockIdx.x * BF_BM;
  const int n0 = blockIdx.y * BN;
  const int kb0 = blockIdx.z * g.kb_per_split;
  int kb1 = kb0 + g.kb_per_split;
  kb1 = kb1 < g.num_kb ? kb1 : g.num_kb;
  const int nkb = kb1 - kb0;
  const bool want_db = g.db != nullptr && blockIdx.y == 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    mbar_fence_init();
  }
  if (want_db) {            // constant operand of the bias-gradient MMAs
    for (int i = threadIdx.x; i < BW_CHUNK_BYTES / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();      // nothing above touches global memory: the prologue overlaps the previous kernel's tail
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t ones_col = static_cast<uint32_t>((BN + 15) & ~15);              // TMEM column of the bias accumulator

  if (warp == 0) {
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
      mbar_wait(&empty_bar[s], ph ^ 1u);
      __syncwarp();
      mbar_expect_tx_w(&full_bar[s], a_bytes + b_bytes);
      const int row = (kb0 + it) * BW_KB;
      unsigned char* dst = smem + static_cast<size_t>(s) * stage_bytes;
      for (int c = 0; c < a_chunks; ++c) {          // columns beyond the last segment: fully out-of-bounds box -> zeros
        const int o = m0 + 32 * c;
        const int sg = find_seg(g.a_start, g.a_nseg, o);
        tma_load_2d_w(dst + c * BW_CHUNK_BYTES, &maps.a[sg], &full_bar[s], o - g.a_start[sg], row);
      }
      dst += a_bytes;
      for (int c = 0; c < b_chunks; ++c) {
        const int i = n0 + 32 * c;
        const int sg = find_seg(g.b_start, g.b_nseg, i);
        tma_load_2d_w(dst + c * BW_CHUNK_BYTES, &maps.b[sg], &full_bar[s], i - g.b_start[sg], row);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_bf16((BN + 15) & ~15, true);
    const uint32_t idesc_ones = idesc_bf16(16, true);
    const uint64_t d_ones = smem_desc_mn_sw64(smem_u32(ones), BW_CHUNK_BYTES);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      __syncwarp();
      const uint32_t sa = smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
      const uint64_t da = smem_desc_mn_sw64(sa, BW_CHUNK_BYTES), db = smem_desc_mn_sw64(sa + a_bytes, BW_CHUNK_BYTES);
      // 16 contraction rows per k-step = two 8-row groups of 512 bytes = 64 descriptor address units
      if (want_db)
        umma_kblock_ss_ones_w(tmem_base, tmem_base + ones_col, da, db, d_ones, idesc, idesc_ones, it != 0 ? 1u : 0u, 64ull,
                              &empty_bar[s]);
      else
        umma_kblock_ss_w(tmem_base, da, db, idesc, it != 0 ? 1u : 0u, 64ull, &empty_bar[s]);
    }
    umma_commit_w(&acc_bar);
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;
    const int cp = ew >> 2;
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    const int o0 = m0 + q * 32;
    if (o0 < g.No) {                               // No % 32 == 0: a lane quarter is valid as a whole
      const uint32_t stg = smem_u32(epi_base + static_cast<size_t>(ew) * 8192u);
      const uint32_t acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
      const int n_ch = (BN + 31) / 32;
      uint32_t k = 0;
      for (int ci = cp; ci < n_ch; ci += BF_EPI_WARPS / 4, ++k) {
        if (n0 + ci * 32 >= g.Ki) break;
        uint32_t r[32];
        tmem_ld32(acc + static_cast<uint32_t>(ci * 32), r);
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
        const uint32_t tile = stg + (k & 1u) * 4096u;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        stage_write_f32(tile, lane, v);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&maps.out, tile, n0 + ci * 32, static_cast<int>(blockIdx.z) * g.No + o0);
          bulk_commit();
        }
      }
      if (want_db && cp == 0) {
        uint32_t b;
        tmem_ld1(acc + ones_col, b);
        g.db[static_cast<int64_t>(blockIdx.z) * g.No + o0 + lane] = __uint_as_float(b);
      }
      if (lane == 0) bulk_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ---------------------------------------------------------------------------------------------- small elementwise kernels
// out[m, n] = cast(in[m, n]) between fp32 and bf16 (n < width, width % 8 == 0; 16-byte accesses on both sides)
template <bool TO_BF16>
__global__ void __launch_bounds__(256) convert_kernel(const void* __restrict__ in, int64_t ldi, void* __restrict__ out, int64_t ldo,
                                                      int64_t M, int w8) {
  pdl_enter();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M * w8) return;
  const int64_t m = i / w8;
  const int c = static_cast<int>(i - m * w8) * 8;
  if constexpr (TO_BF16) {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(in) + m * ldi + c);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(out) + m * ldo + c) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  } else {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(in) + m * ldi + c));
    float4* o = reinterpret_cast<float4*>(static_cast<float*>(out) + m * ldo + c);
    o[0] = make_float4(bf_lo(u.x), bf_hi(u.x), bf_lo(u.y), bf_hi(u.y));
    o[1] = make_float4(bf_lo(u.z), bf_hi(u.z), bf_lo(u.w), bf_hi(u.w));
  }
}
// g_pre[m, n] = g[m, n] * act'(pre[m, n]), all bf16 (fp32 arithmetic)
__global__ void __launch_bounds__(256) act_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ gr, int64_t ldg,
                                                           const __nv_bfloat16* __restrict__ pre, int64_t ldp,
                                                           __nv_bfloat16* __restrict__ out, int64_t ldo, int64_t M, int w8, int act) {
  pdl_enter();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= M * w8) return;
  const int64_t m = i / w8;
  const int c = static_cast<int>(i - m * w8) * 8;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(gr + m * ldg + c));
  const uint4 p = __ldg(reinterpret_cast<const uint4*>(pre + m * ldp + c));
  const uint32_t av[4] = {a.x, a.y, a.z, a.w}, pv[4] = {p.x, p.y, p.z, p.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    o[j] = pack_bf16(bf_lo(av[j]) * act_bwd(act, bf_lo(pv[j])), bf_hi(av[j]) * act_bwd(act, bf_hi(pv[j])));
  *reinterpret_cast<uint4*>(out + m * ldo + c) = make_uint4(o[0], o[1], o[2], o[3]);
}

static const CUtensorMapDataType kBF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;

}  // namespace ax2d

using namespace ax2d;

static unsigned long long* g_bf_dbg = nullptr;
// development aid (not part of include/ax2d.h): 16 x u64 device buffer receiving %globaltimer stamps of CTA 0
extern "C" void ax2d_debug_bf16_timing(unsigned long long* buf) { g_bf_dbg = buf; }

extern "C" int ax2d_gemm_bf16_supported(const ax2d_cmat* a, int64_t M, int64_t N, int64_t K) {
  if (a == nullptr || M < 1 || N < 1 || K < 8) return 0;
  if (a->n_seg < 1 || a->n_seg > AX2D_MAX_SEG) return 0;
  int64_t acc = 0;
  for (int s = 0; s < a->n_seg; ++s) {
    if (a->width[s] <= 0 || a->width[s] % 8 != 0 || a->ld[s] % 8 != 0 || (reinterpret_cast<uintptr_t>(a->ptr[s]) & 15u)) return 0;
    acc += a->width[s];
  }
  return acc == K ? 1 : 0;
}

// C = epilogue(A B^T): a = bf16 segments, b = bf16 [N, K] (ldb), c = bf16 or fp32 segments (c_dtype); in `ep` the pre /
// resid / dact_pre members point to bf16 data, bias to fp32.  Not supported here: explicit masks, accumulate.
extern "C" int ax2d_gemm_bf16(const ax2d_cmat* a, const void* b, int64_t ldb, const ax2d_mat* c, int c_dtype, int64_t M, int64_t N,
                              int64_t K, const ax2d_epilogue* ep, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(a != nullptr && b != nullptr && c != nullptr, "ax2d_gemm_bf16: null operand");
  AX2D_CHECK_ARG(ax2d_gemm_bf16_supported(a, M, N, K), "ax2d_gemm_bf16: unsupported operands M=%lld N=%lld K=%lld (bf16 segment "
                 "widths and leading dimensions must be multiples of 8, pointers 16-byte aligned)", (long long)M, (long long)N, (long long)K);
  AX2D_CHECK_ARG(c_dtype == AX2D_BF16 || c_dtype == AX2D_F32, "ax2d_gemm_bf16: bad output dtype %d", c_dtype);
  AX2D_CHECK_ARG(ldb % 8 == 0 && ldb >= K, "ax2d_gemm_bf16: bad ldb");
  AX2D_CHECK_ALIGN(b);
  AX2D_CHECK_ARG(c->n_seg >= 1 && c->n_seg <= BF_MAX_CSEG, "ax2d_gemm_bf16: at most %d output segments", BF_MAX_CSEG);
  AX2D_CHECK_ARG(M < (1ll << 31) && N < (1 << 24), "ax2d_gemm_bf16: matrix too large");
  const int out_elem = c_dtype == AX2D_F32 ? 4 : 2;
  const int out_align = 16 / out_elem;
  BfArgs g;
  memset(&g, 0, sizeof(g));
  BfMaps maps;
  memset(&maps, 0, sizeof(maps));
  g.dbg = g_bf_dbg;
  g.M = M;
  g.N = static_cast<int>(N);
  g.c_f32 = c_dtype == AX2D_F32 ? 1 : 0;
  int rc;
  // ---- epilogue description
  int act_cols = static_cast<int>(N), dact_cols = static_cast<int>(N);
  int n_res = 0, res_cols[BF_MAX_RES] = {0, 0, 0};
  const ax2d_mat* pre = nullptr;
  bool has_dp = false;
  if (ep != nullptr) {
    AX2D_CHECK_ARG(ep->mask == nullptr && ep->accumulate == 0, "ax2d_gemm_bf16: explicit masks / accumulate are not supported");
    AX2D_CHECK_ARG(ep->drop_p >= 0.f && ep->drop_p < 1.f, "ax2d_gemm_bf16: dropout p=%f", ep->drop_p);
    AX2D_CHECK_ALIGN(ep->bias);
    g.bias = ep->bias;
    g.act = ep->act;
    act_cols = ep->act != AX2D_ACT_NONE ? ep->act_cols : static_cast<int>(N);
    g.drop_p = ep->drop_p; g.drop_seed = ep->drop_seed; g.drop_tick = ep->drop_tick;
    AX2D_CHECK_ARG(ep->resid.n_seg >= 0 && ep->resid.n_seg <= BF_MAX_RES, "ax2d_gemm_bf16: at most %d residuals", BF_MAX_RES);
    n_res = ep->resid.n_seg;
    for (int r = 0; r < n_res; ++r) {
      res_cols[r] = ep->resid.width[r] > 0 ? ep->resid.width[r] : static_cast<int>(N);
      AX2D_CHECK_ARG(ep->resid.ld[r] % 8 == 0 && (reinterpret_cast<uintptr_t>(ep->resid.ptr[r]) & 15u) == 0, "ax2d_gemm_bf16: residual alignment");
      if ((rc = make_map(&maps.res[r], ep->resid.ptr[r], res_cols[r], M, ep->resid.ld[r], 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, kBF, 2)) != AX2D_OK) return rc;
    }
    if (ep->dact_pre != nullptr && ep->dact != AX2D_ACT_NONE) {
      has_dp = true;
      g.dact = ep->dact;
      dact_cols = ep->dact_cols > 0 ? ep->dact_cols : static_cast<int>(N);
      AX2D_CHECK_ARG(ep->ld_dact % 8 == 0 && (reinterpret_cast<uintptr_t>(ep->dact_pre) & 15u) == 0, "ax2d_gemm_bf16: dact_pre alignment");
      if ((rc = make_map(&maps.dp, ep->dact_pre, dact_cols, M, ep->ld_dact, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, kBF, 2)) != AX2D_OK) return rc;
    }
    if (ep->pre.n_seg > 0) pre = &ep->pre;
  }
  AX2D_CHECK_ARG(!(g.act != AX2D_ACT_NONE && g.dact != AX2D_ACT_NONE), "ax2d_gemm_bf16: act and dact together");
  // ---- output / pre segments
  int c_start[BF_MAX_CSEG + 1], p_start[BF_MAX_CSEG + 1];
  int acc = 0;
  for (int s = 0; s < c->n_seg; ++s) {
    c_start[s] = acc;
    AX2D_CHECK_ARG(c->width[s] > 0 && c->ld[s] % out_align == 0 && c->ptr[s] != nullptr && (reinterpret_cast<uintptr_t>(c->ptr[s]) & 15u) == 0,
                   "ax2d_gemm_bf16: output segment %d: width / ld / alignment", s);
    if (g.c_f32) {
      if ((rc = make_map(&maps.c[s], c->ptr[s], c->width[s], M, c->ld[s], 32, 32, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4)) != AX2D_OK) return rc;
    } else {
      if ((rc = make_map(&maps.c[s], c->ptr[s], c->width[s], M, c->ld[s], 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, kBF, 2)) != AX2D_OK) return rc;
    }
    acc += c->width[s];
  }
  c_start[c->n_seg] = acc;
  AX2D_CHECK_ARG(acc == N, "ax2d_gemm_bf16: C segments cover %d columns, expected %lld", acc, (long long)N);
  int p_nseg = 0;
  bool p_on[BF_MAX_CSEG] = {false, false, false, false};
  if (pre != nullptr) {
    AX2D_CHECK_ARG(pre->n_seg <= BF_MAX_CSEG, "ax2d_gemm_bf16: at most %d pre-activation segments", BF_MAX_CSEG);
    acc = 0;
    for (int s = 0; s < pre->n_seg; ++s) {
      p_start[s] = acc;
      AX2D_CHECK_ARG(pre->width[s] > 0, "ax2d_gemm_bf16: pre segment width");
      if (pre->ptr[s] != nullptr) {
        AX2D_CHECK_ARG(pre->ld[s] % 8 == 0 && (reinterpret_cast<uintptr_t>(pre->ptr[s]) & 15u) == 0, "ax2d_gemm_bf16: pre segment alignment");
        if ((rc = make_map(&maps.pre[s], pre->ptr[s], pre->width[s], M, pre->ld[s], 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, kBF, 2)) != AX2D_OK) return rc;
        p_on[s] = true;
      }
      acc += pre->width[s];
    }
    p_nseg = pre->n_seg;
    p_start[p_nseg] = acc;
    AX2D_CHECK_ARG(acc == N, "ax2d_gemm_bf16: pre segments cover %d columns, expected %lld", acc, (long long)N);
  }
  // ---- column tiles: every output segment is cut into tiles of at most bn_max columns (multiples of 32 except the last
  // tile of a segment, which ends with the segment = the end of its tensor map, where bulk stores clip)
  const int64_t m_tiles = (M + BF_BM - 1) / BF_BM;
  int bn_max = BF_MAX_BN;
  if (m_tiles * ((N + BF_MAX_BN - 1) / BF_MAX_BN) < kNumSMs) {          // few row tiles: narrower tiles until ~one per SM
    const int64_t want = (kNumSMs + m_tiles - 1) / m_tiles;
    int bn = static_cast<int>((N + want - 1) / want);
    bn = (bn + 31) / 32 * 32;
    bn_max = bn < 32 ? 32 : (bn > BF_MAX_BN ? BF_MAX_BN : bn);
  }
  int n_ct = 0, bn_box = 16;
  for (int s = 0; s < c->n_seg; ++s) {
    const int w = c->width[s];
    const int parts = (w + bn_max - 1) / bn_max;
    int per = ((w + parts - 1) / parts + 31) / 32 * 32;
    for (int off = 0; off < w; off += per) {
      AX2D_CHECK_ARG(n_ct < BF_MAX_CT, "ax2d_gemm_bf16: more than %d column tiles", BF_MAX_CT);
      BfColTile& t = g.ct[n_ct++];
      t.n0 = c_start[s] + off;
      t.w = w - off < per ? w - off : per;
      t.cseg = s;
      t.cn0 = off;
      t.pseg = -1;
      t.pn0 = 0;
      const int n1 = t.n0 + t.w;
      if (pre != nullptr) {
        int ps = 0;
        while (ps + 1 < p_nseg && t.n0 >= p_start[ps + 1]) ++ps;
        AX2D_CHECK_ARG(n1 <= p_start[ps + 1], "ax2d_gemm_bf16: a pre-activation segment boundary cuts an output tile");
        if (p_on[ps]) { t.pseg = ps; t.pn0 = t.n0 - p_start[ps]; }
      }
      t.flags = 0;
      if (g.act != AX2D_ACT_NONE) {
        AX2D_CHECK_ARG(n1 <= act_cols || t.n0 >= act_cols, "ax2d_gemm_bf16: act_cols cuts an output tile");
        if (n1 <= act_cols) t.flags |= 1u;
      }
      if (has_dp) {
        AX2D_CHECK_ARG(n1 <= dact_cols || t.n0 >= dact_cols, "ax2d_gemm_bf16: dact_cols cuts an output tile");
        if (n1 <= dact_cols) t.flags |= 2u;
      }
      for (int r = 0; r < n_res; ++r) {
        AX2D_CHECK_ARG(n1 <= res_cols[r] || t.n0 >= res_cols[r], "ax2d_gemm_bf16: a residual width cuts an output tile");
        if (n1 <= res_cols[r]) t.flags |= 4u << r;
      }
      const int nm = (t.w + 15) & ~15;
      bn_box = nm > bn_box ? nm : bn_box;
    }
  }
  g.n_ct = n_ct;
  g.bn_box = bn_box;
  AX2D_CHECK_ARG(m_tiles * n_ct < (1ll << 30), "ax2d_gemm_bf16: too many tiles");
  g.total_tiles = static_cast<int>(m_tiles * n_ct);
  // ---- A segments (k-blocks of 64; the tail of a segment is zero-filled by the TMA unit) and B
  int kb = 0, k0 = 0;
  g.n_seg = a->n_seg;
  for (int s = 0; s < a->n_seg; ++s) {
    g.seg_kb_start[s] = kb;
    g.seg_k0[s] = k0;
    if ((rc = make_map(&maps.a[s], a->ptr[s], a->width[s], M, a->ld[s], BF_BM, BF_BK, CU_TENSOR_MAP_SWIZZLE_128B, kBF, 2)) != AX2D_OK) return rc;
    kb += (a->width[s] + BF_BK - 1) / BF_BK;
    k0 += a->width[s];
  }
  for (int s = a->n_seg; s <= AX2D_MAX_SEG; ++s) g.seg_kb_start[s] = kb;
  g.num_kb = kb;
  if ((rc = make_map(&maps.b, b, K, N, ldb, bn_box, BF_BK, CU_TENSOR_MAP_SWIZZLE_128B, kBF, 2)) != AX2D_OK) return rc;
  // ---- shared memory: epilogue staging first, the k-block ring takes the rest
  int n_in_max = 0;
  for (int i = 0; i < n_ct; ++i) {
    int n = 0;
    for (int r = 0; r < BF_MAX_RES; ++r) n += (g.ct[i].flags >> (2 + r)) & 1u;
    n += (g.ct[i].flags >> 1) & 1u;
    n_in_max = n > n_in_max ? n : n_in_max;
  }
  g.n_in_max = n_in_max;
  g.n_out = pre != nullptr ? 2 : 1;
  const size_t out_tile = g.c_f32 ? 4096 : 2048;
  const size_t stage_bytes = static_cast<size_t>(BF_BM + bn_box) * BF_BK * 2;
  g.in_double = 1;
  size_t epi_bytes = static_cast<size_t>(BF_EPI_WARPS) * (2 * n_in_max * BF_IN_TILE + 2 * g.n_out * out_tile);
  if (227 * 1024 - 2048 - BF_BIAS_SMEM * 4 < epi_bytes + 3 * stage_bytes) {       // keep >= 3 ring stages: single-buffer the inputs instead
    g.in_double = 0;
    epi_bytes = static_cast<size_t>(BF_EPI_WARPS) * (n_in_max * BF_IN_TILE + 2 * g.n_out * out_tile);
  }
  epi_bytes += BF_BIAS_SMEM * 4;
  const size_t budget = 227 * 1024 - 1024 - 1024 - epi_bytes;
  int stages = static_cast<int>(budget / stage_bytes);
  stages = stages > BF_MAX_STAGES ? BF_MAX_STAGES : stages;
  AX2D_CHECK_ARG(stages >= 2, "ax2d_gemm_bf16: shared memory: %zu B of epilogue staging leave room for %d stage(s)", epi_bytes, stages);
  g.stages = stages;
  size_t smem = stages * stage_bytes + epi_bytes + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;    // never two CTAs on an SM: each allocates all 512 TMEM columns
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_error("ax2d_gemm_bf16: cannot raise the dynamic shared memory limit to %zu: %s", smem, cudaGetErrorString(e));
      return AX2D_ERR_LAUNCH;
    }
    configured = smem;
  }
  dim3 grid(static_cast<unsigned>(g.total_tiles < kNumSMs ? g.total_tiles : kNumSMs));
  launch_k(gemm_bf16_kernel, dim3(grid), dim3(BF_THREADS), smem, reinterpret_cast<cudaStream_t>(stream), maps, g);
  return launch_status("ax2d_gemm_bf16");
}

extern "C" int ax2d_gemm_bf16_wgrad_supported(const ax2d_cmat* a, const ax2d_cmat* b, int64_t M, int64_t N, int64_t K) {
  // a: G [K rows, M columns (segmented)], b: X [K rows, N columns (segmented)]
  if (a == nullptr || b == nullptr || M < 32 || N < 32 || M % 32 != 0 || N % 8 != 0 || K < 1) return 0;
  const ax2d_cmat* ops_[2] = {a, b};
  for (const ax2d_cmat* m : ops_) {
    if (m->n_seg < 1 || m->n_seg > AX2D_MAX_SEG) return 0;
    for (int s = 0; s < m->n_seg; ++s)
      if (m->width[s] <= 0 || m->width[s] % 32 != 0 || m->ld[s] % 8 != 0 || (reinterpret_cast<uintptr_t>(m->ptr[s]) & 15u)) return 0;
  }
  return 1;
}

static int bw_n_tiles(int64_t N) { return static_cast<int>((N + BF_MAX_BN - 1) / BF_MAX_BN); }
static int bw_split(int64_t tiles, int64_t num_kb) {
  int64_t hi = 4 * ((kNumSMs + tiles - 1) / tiles);
  hi = hi > num_kb ? num_kb : hi;
  int64_t best = 1, best_cost = -1;
  for (int64_t sp = 1; sp <= hi; ++sp) {
    const int64_t per = (num_kb + sp - 1) / sp;
    const int64_t real = (num_kb + per - 1) / per;
    const int64_t waves = (tiles * real + kNumSMs - 1) / kNumSMs;
    const int64_t cost = waves * (per + 4) * 64 + real;             // prologue + epilogue ~ 4 k-blocks; ties: fewer partials
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = real; }
  }
  return static_cast<int>(best);
}
extern "C" int ax2d_gemm_bf16_wgrad_splits(int64_t M, int64_t N, int64_t K) {
  return bw_split(((M + BF_BM - 1) / BF_BM) * bw_n_tiles(N), (K + BW_KB - 1) / BW_KB);
}
extern "C" int64_t ax2d_gemm_bf16_wgrad_workspace(int64_t M, int64_t N, int64_t K) {
  return static_cast<int64_t>(ax2d_gemm_bf16_wgrad_splits(M, N, K)) * (M * N + M) * 4;      // partial tiles + partial bias vectors
}

// C[M,N] (fp32, contiguous rows of ldc) (+)= A^T B; accumulate: 0 write, 1 add, 2 leave the partials [splits][M][N] (+ the
// partial bias vectors [splits][M]) in the workspace for ax2d_unpack_grads.  The workspace is always required.
extern "C" int ax2d_gemm_bf16_wgrad(const ax2d_cmat* a, const ax2d_cmat* b, const ax2d_mat* c, int64_t M, int64_t N, int64_t K,
                                    int accumulate, float* bias_grad, void* workspace, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(ax2d_gemm_bf16_wgrad_supported(a, b, M, N, K), "ax2d_gemm_bf16_wgrad: unsupported operands M=%lld N=%lld K=%lld "
                 "(segment widths must be multiples of 32)", (long long)M, (long long)N, (long long)K);
  AX2D_CHECK_ARG(workspace != nullptr, "ax2d_gemm_bf16_wgrad: workspace required (ax2d_gemm_bf16_wgrad_workspace)");
  AX2D_CHECK_ALIGN(workspace);
  AX2D_CHECK_ARG(accumulate == 2 || c != nullptr, "ax2d_gemm_bf16_wgrad: null output");
  BwArgs g;
  memset(&g, 0, sizeof(g));
  BwMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc, acc = 0;
  g.No = static_cast<int>(M);
  g.Ki = static_cast<int>(N);
  g.rows = K;
  g.a_nseg = a->n_seg;
  for (int s = 0; s < a->n_seg; ++s) {
    g.a_start[s] = acc;
    if ((rc = make_map(&maps.a[s], a->ptr[s], a->width[s], K, a->ld[s], BW_KB, 32, CU_TENSOR_MAP_SWIZZLE_64B, kBF, 2)) != AX2D_OK) return rc;
    acc += a->width[s];
  }
  for (int s = a->n_seg; s <= AX2D_MAX_SEG; ++s) g.a_start[s] = acc;
  AX2D_CHECK_ARG(acc == M, "ax2d_gemm_bf16_wgrad: A segments cover %d columns, expected %lld", acc, (long long)M);
  acc = 0;
  g.b_nseg = b->n_seg;
  for (int s = 0; s < b->n_seg; ++s) {
    g.b_start[s] = acc;
    if ((rc = make_map(&maps.b[s], b->ptr[s], b->width[s], K, b->ld[s], BW_KB, 32, CU_TENSOR_MAP_SWIZZLE_64B, kBF, 2)) != AX2D_OK) return rc;
    acc += b->width[s];
  }
  for (int s = b->n_seg; s <= AX2D_MAX_SEG; ++s) g.b_start[s] = acc;
  AX2D_CHECK_ARG(acc == N, "ax2d_gemm_bf16_wgrad: B segments cover %d columns, expected %lld", acc, (long long)N);
  const int n_tiles = bw_n_tiles(N);
  int BN = static_cast<int>((N + n_tiles - 1) / n_tiles);
  BN = (BN + 31) / 32 * 32;
  g.BN = BN;
  const int m_tiles = static_cast<int>((M + BF_BM - 1) / BF_BM);
  g.num_kb = static_cast<int>((K + BW_KB - 1) / BW_KB);
  int split = bw_split(static_cast<int64_t>(m_tiles) * n_tiles, g.num_kb);
  g.kb_per_split = (g.num_kb + split - 1) / split;
  split = (g.num_kb + g.kb_per_split - 1) / g.kb_per_split;
  float* ws = static_cast<float*>(workspace);
  g.db = bias_grad != nullptr || accumulate == 2 ? ws + static_cast<int64_t>(split) * M * N : nullptr;
  if ((rc = make_map(&maps.out, ws, N, static_cast<int64_t>(split) * M, N, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4)) != AX2D_OK) return rc;
  const size_t stage_bytes = static_cast<size_t>(BF_BM / 32 + (BN + 31) / 32) * BW_CHUNK_BYTES;
  const size_t fixed = BW_CHUNK_BYTES + static_cast<size_t>(BF_EPI_WARPS) * 8192 + 1024;
  int stages = static_cast<int>((226 * 1024 - fixed) / stage_bytes);
  stages = stages > BW_MAX_STAGES ? BW_MAX_STAGES : stages;
  stages = stages > g.kb_per_split ? g.kb_per_split : stages;
  if (stages < 1) stages = 1;
  g.stages = stages;
  size_t smem = stages * stage_bytes + fixed;
  if (smem < 120 * 1024) smem = 120 * 1024;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_error("ax2d_gemm_bf16_wgrad: cannot raise the dynamic shared memory limit to %zu: %s", smem, cudaGetErrorString(e));
      return AX2D_ERR_LAUNCH;
    }
    configured = smem;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>(m_tiles), static_cast<unsigned>(n_tiles), static_cast<unsigned>(split));
  launch_k(gemm_bf16_wgrad_kernel, dim3(grid), dim3(BF_THREADS), smem, st, maps, g);
  rc = launch_status("ax2d_gemm_bf16_wgrad");
  if (rc != AX2D_OK || accumulate == 2) return rc;
  SegOut out;
  if ((rc = to_out(c, &out, N, "C", false)) != AX2D_OK) return rc;
  return splitk_reduce(ws, split, M, N, out, accumulate == 1, bias_grad != nullptr ? g.db : nullptr, bias_grad, st);
}

extern "C" int ax2d_convert(const void* in, int64_t ldi, int in_dtype, void* out, int64_t ldo, int out_dtype, int64_t M, int width,
                            ax2d_stream_t stream) {
  AX2D_CHECK_ARG(in != nullptr && out != nullptr && M >= 0 && width > 0 && width % 8 == 0 && ldi % 8 == 0 && ldo % 8 == 0,
                 "ax2d_convert: width and leading dimensions must be multiples of 8");
  AX2D_CHECK_ARG((in_dtype == AX2D_F32 && out_dtype == AX2D_BF16) || (in_dtype == AX2D_BF16 && out_dtype == AX2D_F32),
                 "ax2d_convert: fp32 <-> bf16 only");
  AX2D_CHECK_ALIGN(in);
  AX2D_CHECK_ALIGN(out);
  if (M == 0) return AX2D_OK;
  const int64_t total = M * (width / 8);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (out_dtype == AX2D_BF16) launch_k(convert_kernel<true>, dim3(blocks), dim3(256), 0, st, in, ldi, out, ldo, M, width / 8);
  else launch_k(convert_kernel<false>, dim3(blocks), dim3(256), 0, st, in, ldi, out, ldo, M, width / 8);
  return launch_status("ax2d_convert");
}

extern "C" int ax2d_act_bwd_bf16(const void* g, int64_t ldg, const void* pre, int64_t ldp, void* out, int64_t ldo, int64_t M,
                                 int width, int act, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(width > 0 && width % 8 == 0 && ldg % 8 == 0 && ldp % 8 == 0 && ldo % 8 == 0, "ax2d_act_bwd_bf16: widths %% 8");
  AX2D_CHECK_ALIGN(g);
  AX2D_CHECK_ALIGN(pre);
  AX2D_CHECK_ALIGN(out);
  if (M <= 0) return AX2D_OK;
  const int64_t total = M * (width / 8);
  launch_k(act_bwd_bf16_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(g), ldg, static_cast<const __nv_bfloat16*>(pre), ldp, static_cast<__nv_bfloat16*>(out), ldo, M,
      width / 8, act);
  return launch_status("ax2d_act_bwd_bf16");
}
