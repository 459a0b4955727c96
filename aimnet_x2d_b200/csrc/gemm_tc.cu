// Dense projections on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), fp32-faithful.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T )       A: activations (column-segmented, K-major, fp32)
//                                                 B: weights, K-major, pre-split into (hi, lo) fp32 arrays
//
// The fp32 configuration must match the reference within 1e-5 (BASELINE.json), which rules out plain TF32
// (2^-11 operand rounding).  Every operand is therefore split into two TF32-representable terms,
//   a = a_hi + a_lo,  a_hi = a rounded to 10 explicit mantissa bits,  a_lo = a - a_hi (exact in fp32),
// (a_lo additionally rounded to TF32: <= 2^-23 |a| dropped), and the product is accumulated in fp32 TMEM as
//   a_lo*b_hi + a_hi*b_lo + a_hi*b_hi   (+ a_lo*b_lo with -DAX2D_TC_TERMS=4)
// three kind::tf32 MMAs per k-step ("3xTF32").  Every partial product is exact in fp32 (11 x 11 significant bits).
//
// Persistent kernel, one CTA per SM walking 128 x BN output tiles (BN <= 192), 18 warps:
//   warp 0      : TMA producer -- per k-block (16 fp32 = one 64-byte swizzle row) one tensor-map load of the raw
//                 A tile [128 x 16] and of B_hi / B_lo [BN x 16] into a ring of shared-memory stages (SWIZZLE_64B),
//                 completion on an mbarrier (complete_tx).
//   warps 2..5  : splitters -- thread = row of the A tile: read the raw row from shared memory, split it and store
//                 (a_hi, a_lo) into a ring of TENSOR-MEMORY slots (tcgen05.st; lane = row, column = k).
//   warp 1      : issues the tcgen05.mma instructions (A from TMEM, B from shared-memory descriptors, accumulators in
//                 TMEM, double-buffered); tcgen05.commit releases each stage back to the producer and hands finished
//                 accumulators to the epilogue.  The warp also owns the TMEM allocation.
//   warps 6..17 : epilogue -- tcgen05.ld (32 lanes x 32 columns per warp) -> swizzled shared-memory transpose -> the
//                 fused epilogue of gemm_common.cuh with row-contiguous 128-bit global accesses (bias,
//                 pre-activation copy, activation, counter-based dropout, residuals, activation backward, segmented
//                 outputs) -- of tile i while the other warps already run the main loop of tile i + 1.
#include "tc_ptx.cuh"

namespace ax2d {

struct TcMaps {
  CUtensorMap a[AX2D_MAX_SEG];
  CUtensorMap b_hi;
  CUtensorMap b_lo;
};

struct TcArgs {
  unsigned long long* dbg;             // development aid: phase timestamps of CTA (0,0) (ax2d_debug_timing); normally null
  EpiArgs e;
  int seg_kb_start[AX2D_MAX_SEG + 1];   // first k-block of every A segment
  int n_seg;
  int num_kb;
  int BN;
  int stages;
  int tmem_cols;
  int acc2;                            // column offset of the small-term accumulator (0: single accumulator)
  int n_tiles, total_tiles;            // projection kernel: column tiles per row tile, all tiles
  float acc_scale;                     // truncation compensation of the main accumulator (tc_acc_scale)
  int acc_stride;                      // TMEM columns between the two accumulator buffers
  int a_base;                          // first TMEM column of the A ring
};


__device__ __forceinline__ float round_tf32(float v) {
  // round-to-nearest (half away from zero) onto 10 explicit mantissa bits
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
  // v - hi is exact in fp32; lo is rounded to TF32 as well so that the tensor core sees exactly representable
  // inputs whatever its own fp32 -> tf32 conversion does (dropped: <= 2^-23 |v|)
  hi = round_tf32(v);
  lo = round_tf32(v - hi);
}

// CTA-uniform description of the epilogue over the CTA's whole column range [n0, n1): present when every optional
// operand either covers the range completely or not at all and the 128 rows are all valid.  Everything in it derives
// from kernel parameters and blockIdx only, so the branches on it are uniform branches (no divergence bookkeeping),
// and the row loop below needs no per-row validity or null tests: ~3x fewer instructions than the general path.
struct TileEpi {
  bool ok, act, has_pre, has_dp;
  int nres;
  float* c; int ldc;
  float* pre; int ldpre;
  const float* res0; int ldres0;
  const float* res1; int ldres1;
  const float* res2; int ldres2;
  const float* dp; int lddp;
};
__device__ __forceinline__ TileEpi tile_epi(const EpiArgs& e, int64_t m0, int n0, int n1, const float* ws) {
  TileEpi t;
  t.ok = (m0 + TC_BM <= e.M) && e.mask == nullptr && !e.accumulate && e.n_resid <= 3;
  const int cs = find_seg(e.c.start, e.c.n_seg, n0);
  t.ok = t.ok && find_seg(e.c.start, e.c.n_seg, n1 - 1) == cs;
  t.c = e.c.ptr[cs] - e.c.start[cs];           // indexed by absolute column
  t.ldc = static_cast<int>(e.c.ld[cs]);
  if (ws != nullptr) {
    t.c = const_cast<float*>(ws);
    t.ldc = static_cast<int>(e.N);
  }
  t.has_pre = false; t.pre = nullptr; t.ldpre = 0;
  if (e.pre.n_seg > 0) {
    const int ps = find_seg(e.pre.start, e.pre.n_seg, n0);
    t.ok = t.ok && find_seg(e.pre.start, e.pre.n_seg, n1 - 1) == ps;
    if (e.pre.ptr[ps] != nullptr) {
      t.has_pre = true;
      t.pre = e.pre.ptr[ps] - e.pre.start[ps];
      t.ldpre = static_cast<int>(e.pre.ld[ps]);
    }
  }
  t.act = e.act != AX2D_ACT_NONE && n1 <= e.act_cols;
  t.ok = t.ok && (e.act == AX2D_ACT_NONE || n1 <= e.act_cols || n0 >= e.act_cols);
  t.has_dp = e.dact != AX2D_ACT_NONE && n1 <= e.dact_cols;
  t.ok = t.ok && (e.dact == AX2D_ACT_NONE || n1 <= e.dact_cols || n0 >= e.dact_cols);
  t.dp = e.dact_pre; t.lddp = static_cast<int>(e.ld_dact);
  // the (at most three) residuals either cover the whole column range or none of it; kept in their original order
  // (the last ShellConv projection adds h, the skip projection and the layer input: three)
  t.res0 = t.res1 = t.res2 = nullptr;
  t.ldres0 = t.ldres1 = t.ldres2 = 0;
  t.nres = 0;
  for (int r = 0; r < 3; ++r) {
    if (r >= e.n_resid) break;
    const bool covers = n1 <= e.resid_cols[r];
    t.ok = t.ok && (covers || n0 >= e.resid_cols[r]);
    if (!covers) continue;
    const float* p = e.resid[r];
    const int ld = static_cast<int>(e.ld_resid[r]);
    if (t.nres == 0) { t.res0 = p; t.ldres0 = ld; }
    else if (t.nres == 1) { t.res1 = p; t.ldres1 = ld; }
    else { t.res2 = p; t.ldres2 = ld; }
    ++t.nres;
  }
  return t;
}

// 128-bit shared-memory load from a 32-bit shared address.  The address comes out of an opaque asm statement (pin32), so
// the compiler keeps it in a register instead of re-deriving it from the thread index in every iteration, which at
// the kernel's 96-register cap it otherwise does (~20 instructions per two rows).
__device__ __forceinline__ uint32_t pin32(uint32_t v) {
  asm volatile("mov.u32 %0, %0;" : "+r"(v));
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// The eight rows (r0w + 4 i) of one 32 x 32 chunk that a lane finishes on the CTA-uniform path: NRES residuals.
// s_even / s_odd: shared addresses of this lane's float4 in rows r0w and r0w + 4 (row r0w + 4 i is 1024 (i / 2) bytes
// further: 8 rows of 128 bytes; the swizzle term (row & 7) only depends on the parity of i).
template <int ACT, int DACT, bool DROP, int NRES>
__device__ __forceinline__ void lean_rows(const TileEpi& te, const EpiCtx& cx, uint32_t s_even, uint32_t s_odd, float4 b4,
                                          float* pc, float* ppre, const float* pr0, const float* pr1, const float* pr2,
                                          const float* pdp, uint64_t didx, int N) {
#pragma unroll 2      // measured on the C2 shapes: 2 beats 1 (too little in flight) and 4 / 8 (instruction-cache pressure)
  for (int i = 0; i < 8; ++i) {          // rows mrow + 4 i
    float4 r0, r1, r2, dp = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (NRES > 0) r0 = __ldg(reinterpret_cast<const float4*>(pr0));
    if constexpr (NRES > 1) r1 = __ldg(reinterpret_cast<const float4*>(pr1));
    if constexpr (NRES > 2) r2 = __ldg(reinterpret_cast<const float4*>(pr2));
    if constexpr (DACT != AX2D_ACT_NONE) {
      if (te.has_dp) dp = __ldg(reinterpret_cast<const float4*>(pdp));
    }
    const float4 s4 = lds128(((i & 1) ? s_odd : s_even) + static_cast<uint32_t>(i >> 1) * 1024u);
    float v[4] = {fmaf(s4.x, cx.acc_scale, b4.x), fmaf(s4.y, cx.acc_scale, b4.y), fmaf(s4.z, cx.acc_scale, b4.z),
                  fmaf(s4.w, cx.acc_scale, b4.w)};
    if (te.has_pre) *reinterpret_cast<float4*>(ppre) = make_float4(v[0], v[1], v[2], v[3]);
    float drop[4] = {1.f, 1.f, 1.f, 1.f};
    if constexpr (DROP) drop_scale4(cx, didx, drop);
    if constexpr (ACT != AX2D_ACT_NONE) {
      if (te.act) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = act_fwd_t<ACT>(v[j]);
      }
    }
    if constexpr (DROP && DACT == AX2D_ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] *= drop[j];
    }
    if constexpr (NRES > 0) { v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; }
    if constexpr (NRES > 1) { v[0] += r1.x; v[1] += r1.y; v[2] += r1.z; v[3] += r1.w; }
    if constexpr (NRES > 2) { v[0] += r2.x; v[1] += r2.y; v[2] += r2.z; v[3] += r2.w; }
    if constexpr (DACT != AX2D_ACT_NONE) {
      if (te.has_dp) {
        v[0] *= act_bwd_t<DACT>(dp.x) * drop[0];
        v[1] *= act_bwd_t<DACT>(dp.y) * drop[1];
        v[2] *= act_bwd_t<DACT>(dp.z) * drop[2];
        v[3] *= act_bwd_t<DACT>(dp.w) * drop[3];
      }
    }
    *reinterpret_cast<float4*>(pc) = make_float4(v[0], v[1], v[2], v[3]);
    pc += 4 * te.ldc;
    ppre += 4 * te.ldpre;
    if constexpr (NRES > 0) pr0 += 4 * te.ldres0;
    if constexpr (NRES > 1) pr1 += 4 * te.ldres1;
    if constexpr (NRES > 2) pr2 += 4 * te.ldres2;
    if constexpr (DACT != AX2D_ACT_NONE) pdp += 4 * te.lddp;
    if constexpr (DROP) didx += 4ull * static_cast<uint32_t>(N);
  }
}

// Epilogue of one warp: its 32 accumulator rows (TMEM lanes 32 q .. 32 q + 31), 32 columns at a time:
// tcgen05.ld (lane = row, registers = columns) -> shared-memory transpose -> 8 lanes per row, float4 per lane, so
// every global access of the fused epilogue is a row-contiguous 128-byte segment.
// ws != nullptr: raw partial sums go to the split-K workspace slice [M, N] at ws instead of the output segments.
// Inlined into the kernel on purpose: `e` is then known to live in the (immutable, constant-cached) kernel parameter
// space; behind a real call it degrades to generic loads that must be re-issued after every global store.
// The tensor core adds into its fp32 accumulator with TRUNCATION: every MMA that touches an accumulator of magnitude |c|
// loses on average a fixed fraction of ulp(c), always towards zero.  Measured on B200 against float64 (tools/tc_bias.py,
// profiles/): zero-mean operands -> mean signed relative error of the result = -0.27 x 2^-24 per MMA into the accumulator,
// the same coefficient for K = 160 ... 1024 and every tile layout (-9.6e-7 at K = 160, -3.3e-6 at K = 544); the SIMT
// kernel's is < 1e-9.  It is a BIAS, not noise: through the ~20 products between the embeddings and the loss it adds up
// (the full-size C2 step was 1.4e-5 low in the loss, C3 outputs 2.5e-5 of their scale off -- over the 1e-5 bar) instead
// of averaging out.  The epilogue therefore scales the main accumulator by 1 + 0.27 n 2^-24 (n = MMAs accumulated into
// it): one FFMA that replaces the FADD of the bias / small-term sum, no extra instruction.  What remains is zero-mean.
__host__ __device__ __forceinline__ float tc_acc_scale(int n_mma) { return 1.f + 0.27f * static_cast<float>(n_mma) * 5.9604645e-8f; }

template <int ACT, int DACT, bool DROP>
__device__ __forceinline__ void tc_epilogue(const EpiArgs& e, int BN, float* ws, const EpiCtx& cx_in, uint32_t tmem_base,
                                         float* stg, int m0, int n0, int q, int lane, int first_chunk, uint32_t acc2,
                                         float acc_scale, unsigned long long* dbg = nullptr, int chunk_step = 64) {
  // with a separate small-term accumulator the scale is applied to the main one where the two are summed; otherwise where
  // the bias is added (lean_rows / epi_finish)
  EpiCtx cx = cx_in;
  cx.acc_scale = acc2 != 0 ? 1.f : acc_scale;
  const int64_t M = e.M;
  const int N = static_cast<int>(e.N);
  const int cg = lane & 7;
  const TileEpi te = tile_epi(e, m0, n0, (n0 + BN < N ? n0 + BN : N), ws);
  // the warps that share a TMEM lane quarter take the 32-column chunks round-robin
  for (int c0 = 32 * first_chunk; c0 < BN; c0 += chunk_step) {
    if (n0 + c0 >= N) break;
    uint32_t r[32];
    if (dbg != nullptr && c0 == 0) dbg[8] = gtime();
    tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0), r);
    if (dbg != nullptr && c0 == 0) dbg[9] = gtime();
    if (acc2 != 0) {      // main (hi*hi) + small-term accumulator, summed here in round-to-nearest
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r2[16];
        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc2 + static_cast<uint32_t>(c0 + 16 * hh), r2);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          r[16 * hh + j] = __float_as_uint(fmaf(__uint_as_float(r[16 * hh + j]), acc_scale, __uint_as_float(r2[j])));
      }
    }
    // transpose through shared memory with 128-bit accesses: row `lane` stores its eight float4 column groups at
    // slots (group ^ (lane & 7)) of its 128-byte row -- conflict-free for the stores (a quarter-warp hits 8 different
    // slots) and for the row-contiguous reads below (a quarter-warp reads one whole row)
    float4* stg4 = reinterpret_cast<float4*>(stg);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      stg4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                      __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
    __syncwarp();
    if (dbg != nullptr && c0 == 0) dbg[10] = gtime();
    const int n = n0 + c0 + cg * 4;
    if (te.ok) {
      if (n < N) {
        const float4 b4 = e.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(e.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int64_t mrow = static_cast<int64_t>(m0) + q * 32 + (lane >> 3);
        float* pc = te.c + mrow * te.ldc + n;
        float* ppre = te.pre + mrow * te.ldpre + n;
        const float* pr0 = te.res0 + mrow * te.ldres0 + n;
        const float* pr1 = te.res1 + mrow * te.ldres1 + n;
        const float* pr2 = te.res2 + mrow * te.ldres2 + n;
        const float* pdp = te.dp + mrow * te.lddp + n;
        const int r0w = lane >> 3;               // this lane's rows: r0w + 4 i
        uint64_t didx = static_cast<uint64_t>(mrow) * static_cast<uint32_t>(N) + static_cast<uint32_t>(n);
        // the row loop is compiled once per residual count: without it every row carried the zero-initialisation,
        // predicated loads, adds and pointer bookkeeping of three optional residuals (~25 of ~87 instructions per
        // float4 row of the plain SiLU epilogue)
        const int sw0 = cg ^ r0w;
        const uint32_t s_even = pin32(smem_u32(stg4 + r0w * 8 + sw0));
        const uint32_t s_odd = pin32(smem_u32(stg4 + (r0w + 4) * 8 + (sw0 ^ 4)));
        switch (te.nres) {
          case 0: lean_rows<ACT, DACT, DROP, 0>(te, cx, s_even, s_odd, b4, pc, ppre, pr0, pr1, pr2, pdp, didx, N); break;
          case 1: lean_rows<ACT, DACT, DROP, 1>(te, cx, s_even, s_odd, b4, pc, ppre, pr0, pr1, pr2, pdp, didx, N); break;
          case 2: lean_rows<ACT, DACT, DROP, 2>(te, cx, s_even, s_odd, b4, pc, ppre, pr0, pr1, pr2, pdp, didx, N); break;
          default: lean_rows<ACT, DACT, DROP, 3>(te, cx, s_even, s_odd, b4, pc, ppre, pr0, pr1, pr2, pdp, didx, N); break;
        }
      }
    } else if (n < N) {
      EpiCol col = epi_col(e, n);
      if (ws != nullptr) {
        col.c = ws + n;
        col.ldc = N;
      }
      if (dbg != nullptr && c0 == 0) dbg[11] = gtime();
      // rolled on purpose (a fully unrolled body is ~2.7 k instructions per template instance and thrashes the 32 KB
      // instruction cache); two rows per iteration, their loads issued before either is consumed.  All row pointers
      // advance by 8 rows per iteration instead of being recomputed from 64-bit products.
      EpiCol ca = col;
      const int64_t mfirst = static_cast<int64_t>(m0) + q * 32 + (lane >> 3);
      ca.c += mfirst * col.ldc;
      if (ca.pre != nullptr) ca.pre += mfirst * col.ldpre;
#pragma unroll
      for (int rr = 0; rr < kEpiPrefetchResid; ++rr)
        if (ca.resid[rr] != nullptr) ca.resid[rr] += mfirst * col.ldr[rr];
      if (ca.dpre != nullptr) ca.dpre += mfirst * col.lddp;
      if (ca.mask != nullptr) ca.mask += mfirst * col.ldm;
      const int r0w = lane >> 3;
#pragma unroll 1
      for (int i = 0; i < 8; i += 2) {
        const int ra = r0w + 4 * i, rb = ra + 4;
        const float4 sa = stg4[ra * 8 + (cg ^ (ra & 7))], sb = stg4[rb * 8 + (cg ^ (rb & 7))];
        const int64_t ma = mfirst + 4 * i, mb = ma + 4;
        EpiOperands oa, ob;
        if (ma < M) epi_prefetch(e, ca, 0, oa);
        if (mb < M) epi_prefetch(e, ca, 4, ob);
        if (ma < M) epi_finish<ACT, DACT, DROP>(e, cx, ca, 0, oa, sa.x, sa.y, sa.z, sa.w, ma);
        if (mb < M) epi_finish<ACT, DACT, DROP>(e, cx, ca, 4, ob, sb.x, sb.y, sb.z, sb.w, mb);
        if (dbg != nullptr && c0 == 0 && i == 0) dbg[12] = gtime();
        ca.c += 8 * col.ldc;
        if (ca.pre != nullptr) ca.pre += 8 * col.ldpre;
#pragma unroll
        for (int rr = 0; rr < kEpiPrefetchResid; ++rr)
          if (ca.resid[rr] != nullptr) ca.resid[rr] += 8 * col.ldr[rr];
        if (ca.dpre != nullptr) ca.dpre += 8 * col.lddp;
        if (ca.mask != nullptr) ca.mask += 8 * col.ldm;
      }
    }
    if (dbg != nullptr && c0 == 0) dbg[13] = gtime();
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- kernel
// Persistent: one CTA per SM walks the output tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (column tiles of one
// row tile are adjacent in that order, so the CTAs that share an A tile run at the same time and it is read from HBM
// once).
//
// Two things bound the first version of this kernel (one tile per CTA, two CTAs per SM, both operands from shared
// memory), both measured with per-CTA %globaltimer stamps:
//  * all CTAs moved through "main loop" (tensor-bound) and "epilogue" (HBM-bound) in lockstep, so neither resource
//    was busy for more than half of the time.  Here the accumulator is double-buffered in TMEM and the epilogue warps
//    drain tile i while the producer / splitter / MMA warps run the main loop of tile i + 1;
//  * shared-memory bandwidth: four MMAs per k-step, each re-reading its A and B tiles from shared memory, plus the
//    splitters' read-modify-write of A came to ~124 KB per k-block = ~970 cycles at 128 B/cycle, against 640 cycles
//    of tensor time.  Here the split A operand goes to TENSOR MEMORY (tcgen05.st, lane = row, column = k) and the
//    MMAs take it from there (tcgen05.mma with a TMEM A operand); shared memory only carries the raw A tile once
//    and the B tiles: ~76 KB per k-block.
//
// TMEM columns (W = acc_stride >= BN, at most 192): [0, W) accumulator 0, [W, 2 W) accumulator 1 (small-term
// accumulators at +96 when BN <= 96), [2 W, 512) the A ring: 32 columns (hi 16 | lo 16) per pipeline stage.
constexpr int TC_MAX_BN = 192;
template <int BK>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const __grid_constant__ TcMaps maps, const __grid_constant__ TcArgs g) {
  static_assert(BK == 16 || BK == 32, "k-block = one 64-byte or one 128-byte swizzle row of fp32");
  constexpr int TC_BK = BK;                         // shadows the namespace constant inside this kernel
  constexpr uint32_t TC_A_BYTES = TC_BM * BK * 4;   // raw A tile of a stage: 8 or 16 KB
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[TC_MAX_STAGES], split_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = g.BN, S = g.stages;
  const uint32_t b_bytes = static_cast<uint32_t>(BN) * TC_BK * 4;
  const uint32_t stage_bytes = TC_A_BYTES + 2 * b_bytes;
  // align the dynamic region to 1024 B (swizzle atoms)
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the address space and
  // emits LDS / STS instead of generic loads and stores for everything derived from it)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  auto stage_a = [&](int s) { return smem + static_cast<size_t>(s) * stage_bytes; };
  auto stage_bhi = [&](int s) { return stage_a(s) + TC_A_BYTES; };
  auto stage_blo = [&](int s) { return stage_a(s) + TC_A_BYTES + b_bytes; };
  float* stg_base = reinterpret_cast<float*>(smem + static_cast<size_t>(S) * stage_bytes);     // epilogue transposes

  // development aid, all CTAs: dbg[16 + 8 cta + {0..5}] = start, end, SM id, main loops issued, first accumulator
  // ready, last epilogue done
  unsigned long long* cta_dbg = g.dbg != nullptr ? g.dbg + 16 + 32 * blockIdx.x : nullptr;
  if (cta_dbg != nullptr && threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    cta_dbg[0] = gtime();
    cta_dbg[2] = smid;
    cta_dbg[11] = tc_clock();
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], TC_SPLIT_WARPS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], TC_EPI_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();      // nothing above touches global memory: the prologue overlaps the previous kernel's tail
  const uint32_t tmem_base = tmem_base_smem;
  const int n_tiles = g.n_tiles, total = g.total_tiles, num_kb = g.num_kb;

  if (warp == 0) {
    // ===================================================================== TMA producer (whole warp, elected issue)
    int s = 0;              // stage and phase advance by increments: an integer division by the runtime stage count
    uint32_t ph = 0;        // costs this single, latency-exposed warp ~200 cycles per k-block
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * BN;
      int seg = 0;
      for (int kb = 0; kb < num_kb; ++kb, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
        mbar_wait(&empty_bar[s], ph ^ 1u);      // the MMAs that read this stage (smem B and TMEM A slot) are done
        __syncwarp();
        while (seg + 1 < g.n_seg && kb >= g.seg_kb_start[seg + 1]) ++seg;
        tma_kblock_w(&full_bar[s], TC_A_BYTES + 2 * b_bytes, stage_a(s), &maps.a[seg], (kb - g.seg_kb_start[seg]) * TC_BK, m0,
                     stage_bhi(s), &maps.b_hi, stage_blo(s), &maps.b_lo, kb * TC_BK, n0);
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp, elected issue)
    const uint32_t idesc = idesc_tf32(BN);
    int s = 0, j = 0;
    uint32_t ph = 0;
    long long t_issue = 0, t_wait = 0, t_acc = 0;
    const long long t_loop0 = tc_clock();
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++j) {
      const int buf = j & 1;
      const long long ca = tc_clock();
      mbar_wait(&acc_empty[buf], (static_cast<uint32_t>(j >> 1) & 1u) ^ 1u);   // epilogue of tile j - 2 has drained it
      tc_fence_after();
      t_acc += tc_clock() - ca;
      const uint32_t t_main = tmem_base + static_cast<uint32_t>(buf * g.acc_stride);
      const uint32_t t_small = t_main + static_cast<uint32_t>(g.acc2);
      for (int kb = 0; kb < num_kb; ++kb, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
        const long long c0 = tc_clock();
        mbar_wait(&full_bar[s], ph);     // B_hi / B_lo have landed
        mbar_wait(&split_bar[s], ph);    // (a_hi, a_lo) are in TMEM
        tc_fence_after();
        __syncwarp();
        const long long c1 = tc_clock();
        t_wait += c1 - c0;
        const uint32_t a_hi = tmem_base + static_cast<uint32_t>(g.a_base + s * 2 * TC_BK);
        const uint32_t a_lo = a_hi + TC_BK;
        const uint64_t db_hi = BK == 16 ? smem_desc_k_sw64(smem_u32(stage_bhi(s))) : smem_desc_k_sw128(smem_u32(stage_bhi(s)));
        const uint64_t db_lo = BK == 16 ? smem_desc_k_sw64(smem_u32(stage_blo(s))) : smem_desc_k_sw128(smem_u32(stage_blo(s)));
        // The tensor core's fp32 accumulation truncates, so every MMA into a large accumulator costs up to one ulp
        // of it.  When TMEM allows (acc2 != 0) the three small terms go to their own accumulator (2^-11 of the
        // magnitude, so their truncation is negligible) and only hi*hi touches the main one.  Per k-step the
        // descriptors advance by 32 bytes inside the swizzle row (B) and by 8 TMEM columns (A).
        const uint32_t acc_first = kb != 0 ? 1u : 0u;
        if constexpr (BK == 16)
          umma_kblock_ts_w(t_main, t_small, a_hi, a_lo, db_hi, db_lo, idesc, acc_first, g.acc2 != 0 ? acc_first : 1u,
                           &empty_bar[s]);    // the commit frees the stage once these MMAs have read it
        else
          umma_kblock_ts32_w(t_main, t_small, a_hi, a_lo, db_hi, db_lo, idesc, acc_first, g.acc2 != 0 ? acc_first : 1u,
                             &empty_bar[s]);
        t_issue += tc_clock() - c1;
      }
      umma_commit_w(&acc_full[buf]);       // accumulator of this tile complete
    }
    if (cta_dbg != nullptr && lane == 0) { cta_dbg[3] = gtime(); cta_dbg[6] = t_issue; cta_dbg[7] = t_wait; cta_dbg[8] = t_acc; cta_dbg[13] = tc_clock() - t_loop0; }
  } else if (warp < 2 + TC_SPLIT_WARPS) {
    // ===================================================================== splitters: thread = row of the A tile
    // raw row (64 / 128 bytes in the swizzled stage: 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3) for the
    // 64-byte swizzle, c ^ (r & 7) for the 128-byte one, which also makes the 8 rows of a quarter-warp hit 8 different
    // bank groups) -> (hi, lo) -> TMEM lane r, BK columns each
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int sw = BK == 16 ? (row >> 1) & 3 : row & 7;
    int s = 0, prev = -1;   // software pipeline: the TMEM stores of k-block i complete while k-block i + 1 is read and split
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      for (int kb = 0; kb < num_kb; ++kb, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
        mbar_wait(&full_bar[s], ph);
        const float4* a = reinterpret_cast<const float4*>(stage_a(s)) + row * (TC_BK / 4);
        uint32_t hi[TC_BK], lo[TC_BK];
#pragma unroll
        for (int c = 0; c < TC_BK / 4; ++c) {
          const float4 v = a[c ^ sw];
          float h, o;
          split_tf32(v.x, h, o); hi[4 * c + 0] = __float_as_uint(h); lo[4 * c + 0] = __float_as_uint(o);
          split_tf32(v.y, h, o); hi[4 * c + 1] = __float_as_uint(h); lo[4 * c + 1] = __float_as_uint(o);
          split_tf32(v.z, h, o); hi[4 * c + 2] = __float_as_uint(h); lo[4 * c + 2] = __float_as_uint(o);
          split_tf32(v.w, h, o); hi[4 * c + 3] = __float_as_uint(h); lo[4 * c + 3] = __float_as_uint(o);
        }
        if (prev >= 0) {
          tmem_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&split_bar[prev]);
        }
        const uint32_t slot = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g.a_base + s * 2 * TC_BK);
        if constexpr (BK == 16) {
          tmem_st16(slot, hi);
          tmem_st16(slot + TC_BK, lo);
        } else {
          tmem_st32(slot, hi);
          tmem_st32(slot + TC_BK, lo);
        }
        prev = s;
      }
    }
    if (prev >= 0) {
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&split_bar[prev]);
    }
  } else {
    // ===================================================================== epilogue warps
    // TMEM lanes [32 q, 32 q + 32) can be read by the warps with (warp id % 4) == q; two warps per quarter take
    // alternate 32-column chunks.
    const int ew = warp - (2 + TC_SPLIT_WARPS);
    float* stg = stg_base + ew * (32 * 32);
    const EpiCtx cx = epi_ctx(g.e);
    const bool dropping = cx.dropping;
    int j = 0;
    long long e_wait = 0, e_work = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++j) {
      const int m0 = (tile / n_tiles) * TC_BM, n0 = (tile % n_tiles) * BN;
      const int buf = j & 1;
      {
        // while this tile's main loop runs: pull the rows the epilogue will read (residuals, activation-backward
        // operand) into L2, so that its loads do not each pay an HBM round trip with only two warps per scheduler
        const int n1 = n0 + BN < static_cast<int>(g.e.N) ? n0 + BN : static_cast<int>(g.e.N);
        const TileEpi te = tile_epi(g.e, m0, n0, n1, nullptr);
        if (te.ok && (te.nres > 0 || te.has_dp)) {
          const int lines = ((n1 - n0) * 4 + 127) / 128;            // 128-byte lines per row of the tile (<= 6)
          const int et = threadIdx.x - 32 * (2 + TC_SPLIT_WARPS);
          const int l = et & 7;                                      // 8 lanes per row, one line each: no divisions
          if (l < lines) {
            const int col = n0 + l * 32;
            for (int r = et >> 3; r < TC_BM; r += 4 * TC_EPI_WARPS) {
              const int64_t row = m0 + r;
              if (te.nres > 0) prefetch_l2(te.res0 + row * te.ldres0 + col);
              if (te.nres > 1) prefetch_l2(te.res1 + row * te.ldres1 + col);
              if (te.nres > 2) prefetch_l2(te.res2 + row * te.ldres2 + col);
              if (te.has_dp) prefetch_l2(te.dp + row * te.lddp + col);
            }
          }
        }
      }
      const long long ce0 = tc_clock();
      mbar_wait(&acc_full[buf], static_cast<uint32_t>(j >> 1) & 1u);
      tc_fence_after();
      const long long ce1 = tc_clock();
      e_wait += ce1 - ce0;
      if (cta_dbg != nullptr && j == 0 && ew == 0 && lane == 0) cta_dbg[4] = gtime();
      AX2D_EPI_DISPATCH(g.e.act, g.e.dact, dropping, {
        tc_epilogue<ACT, DACT, DROP>(g.e, BN, nullptr, cx, tmem_base + static_cast<uint32_t>(buf * g.acc_stride), stg, m0, n0,
                                     warp & 3, lane, ew >> 2, static_cast<uint32_t>(g.acc2), g.acc_scale,
                                     (cta_dbg != nullptr && ew == 0 && lane == 0 && j == 0) ? cta_dbg + 8 : nullptr,
                                     32 * (TC_EPI_WARPS / 4));
      });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
      e_work += tc_clock() - ce1;
    }
    if (cta_dbg != nullptr && ew == 0 && lane == 0) { cta_dbg[5] = gtime(); cta_dbg[9] = e_wait; cta_dbg[10] = e_work; }
  }
  tc_fence_before();
  __syncthreads();
  if (cta_dbg != nullptr && threadIdx.x == 0) { cta_dbg[1] = gtime(); cta_dbg[12] = tc_clock(); }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ---------------------------------------------------------------------------------------------- weight gradients
//   dW[No, Ki] = sum_m G[m, o] * X[m, i]        (G, X column-segmented activations, contraction over the rows m)
// Both operands are "MN-major" for the tensor core: the contraction index is the ROW of a row-major matrix.  For
// 32-bit MN-major operands the only shared-memory layout the tensor core accepts is the 128-byte swizzle with a
// 32-byte base (cute: Layout_MN_SW128_32B_Atom = Swizzle<2,5,2> o [32 elements x 4 rows], descriptor layout type 1;
// TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  A TMA box [32 rows x 32 columns] (128-byte rows) lands as one column
// chunk of that canonical layout: 4-row groups of 512 B (SBO), column chunks of 32 elements 4096 B apart (LBO).  One k-block = 32 rows of G and X; the CTA's output tile is 128 (o) x BN (i); the
// row range is split over blockIdx.z and the partial tiles are reduced in a fixed order by splitk_reduce_kernel.
constexpr int WG_KB = 32;
// The tensor core accumulates in fp32 with truncation (measured: a systematic ~3e-8 relative loss per accumulation
// step on same-sign sums), so one TMEM accumulator covers at most 32 k-blocks = 1024 rows = 128 accumulation steps;
// longer contractions are cut into more splits, which are then summed by splitk_reduce_kernel in round-to-nearest.
constexpr int WG_MAX_KB_PER_SPLIT = 32;
constexpr int WG_CHUNK_BYTES = WG_KB * 128;      // one [32 x 32] fp32 box
struct WgMaps {
  CUtensorMap a[AX2D_MAX_SEG];
  CUtensorMap b[AX2D_MAX_SEG];
  CUtensorMap a3;                  // merged kernel: 3-D chunk view of the G segment that holds columns 0..159 ...
  CUtensorMap b3[AX2D_MAX_SEG];    // ... and of every X segment
};
struct WgArgs {
  EpiArgs e;                       // M = No, N = Ki; used directly when there is a single split
  int a_start[AX2D_MAX_SEG + 1], a_nseg;
  int b_start[AX2D_MAX_SEG + 1], b_nseg;
  int num_kb, kb_per_split;
  int BN, stages, tmem_cols;
  int a3_chunk0, b3_on;            // merged kernel: first chunk of the G tile in maps.a3 (-1: per-chunk boxes); X tiles that
                                   // lie inside one segment go through maps.b3[segment] (0: per-chunk boxes)
  int acc2;                        // column offset of the small-term accumulator
  float* ws;                       // [splits][No][Ki] or nullptr
  float* db;                       // optional fused bias gradient: column sums of G, [splits][No] partials (or the
                                   // final [No] vector when there is a single split); nullptr = not requested
};

__device__ __forceinline__ uint64_t smem_desc_mn_sw128_32b(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(WG_CHUNK_BYTES >> 4) << 16;   // LBO: next 32-element chunk along M / N
  d |= static_cast<uint64_t>(512 >> 4) << 32;              // SBO: next group of 4 contraction rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;                     // SWIZZLE_128B_BASE32B
  return d;
}

// TMEM columns of the weight-gradient kernel: [0, 160) main accumulator, [160, 320) small-term accumulator,
// [320, 512) the A ring: three slots of 64 columns (32 contraction rows: hi | lo).
constexpr int WG_ACC2 = 160;
constexpr int WG_A_TMEM = 320;
constexpr int WG_MAX_STAGES = 3;
constexpr int WG_MAX_BN = 160;

// One k-block (4 k-steps x 4 split terms) with the A operand in tensor memory, and the commit, under one election.
__device__ __forceinline__ void umma_kblock_wg_w(uint32_t t_main, uint32_t t_small, uint32_t a_hi, uint32_t a_lo, uint64_t dbh,
                                                 uint64_t dbl, uint32_t idesc, uint32_t acc_first, uint64_t* free_bar) {
  static_assert(WG_KB == 32, "four k-steps of 8 per k-block");
#if AX2D_TC_TERMS == 4
#define AX2D_WG_STEP(AH, AL, BH, BL, PF)                                                  \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [" AL "], " BL ", %6, " PF ";\n\t"        \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [" AL "], " BH ", %6, pt;\n\t"            \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [" AH "], " BL ", %6, pt;\n\t"            \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [" AH "], " BH ", %6, " PF ";\n\t"
#else
#define AX2D_WG_STEP(AH, AL, BH, BL, PF)                                                  \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [" AL "], " BH ", %6, " PF ";\n\t"        \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [" AH "], " BL ", %6, pt;\n\t"            \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [" AH "], " BH ", %6, " PF ";\n\t"
#endif
  asm volatile(
      "{\n\t"
      ".reg .pred e, pf, pt;\n\t"
      ".reg .b32 ah1, al1, ah2, al2, ah3, al3;\n\t"
      ".reg .b64 bh1, bl1, bh2, bl2, bh3, bl3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %7, 0;\n\t"
      "setp.eq.b32 pt, %6, %6;\n\t"
      "add.u32 ah1, %2, 8;\n\t add.u32 al1, %3, 8;\n\t add.u64 bh1, %4, 64;\n\t add.u64 bl1, %5, 64;\n\t"
      "add.u32 ah2, %2, 16;\n\t add.u32 al2, %3, 16;\n\t add.u64 bh2, %4, 128;\n\t add.u64 bl2, %5, 128;\n\t"
      "add.u32 ah3, %2, 24;\n\t add.u32 al3, %3, 24;\n\t add.u64 bh3, %4, 192;\n\t add.u64 bl3, %5, 192;\n\t"
      AX2D_WG_STEP("%2", "%3", "%4", "%5", "pf")
      AX2D_WG_STEP("ah1", "al1", "bh1", "bl1", "pt")
      AX2D_WG_STEP("ah2", "al2", "bh2", "bl2", "pt")
      AX2D_WG_STEP("ah3", "al3", "bh3", "bl3", "pt")
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
      "}\n" ::"r"(t_main),
      "r"(t_small), "r"(a_hi), "r"(a_lo), "l"(dbh), "l"(dbl), "r"(idesc), "r"(acc_first), "r"(smem_u32(free_bar))
      : "memory");
#undef AX2D_WG_STEP
}

// Roles (one CTA per SM, 10 warps): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..5 A-splitters
// (thread = output feature o = TMEM lane: reads its column of the raw G chunk, splits it and stores (hi, lo) into the
// TMEM A ring; accumulates the fused bias gradient on the way), warps 6..9 B-splitters (rewrite the raw X chunks in
// place as hi and emit lo beside them; the MN-major swizzled layout is preserved because the split is element-wise).
// Keeping A out of shared memory removes ~1/3 of the shared-memory traffic that bounded the first version
// (both operands from shared memory: 288 KB per k-block against 1280 cycles of tensor time).
__global__ void __launch_bounds__(TC_WG_THREADS, 1) gemm_tc_wgrad_kernel(const __grid_constant__ WgMaps maps,
                                                                      const __grid_constant__ WgArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG_MAX_STAGES], split_bar[WG_MAX_STAGES], empty_bar[WG_MAX_STAGES], acc_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = g.BN, S = g.stages;
  const int a_chunks = TC_BM / 32, b_chunks = BN / 32;
  const uint32_t a_bytes = static_cast<uint32_t>(a_chunks) * WG_CHUNK_BYTES, b_bytes = static_cast<uint32_t>(b_chunks) * WG_CHUNK_BYTES;
  const uint32_t stage_bytes = a_bytes + 2 * b_bytes;
  // (pointer arithmetic on the __shared__ array, not an integer round trip: the compiler keeps the address space and
  // emits LDS / STS instead of generic loads and stores for everything derived from it)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  auto stage_a = [&](int s) { return smem + static_cast<size_t>(s) * stage_bytes; };     // raw G chunks
  auto stage_bhi = [&](int s) { return stage_a(s) + a_bytes; };                          // raw X chunks -> hi
  auto stage_blo = [&](int s) { return stage_a(s) + a_bytes + b_bytes; };

  const int m0 = blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * BN;
  const int kb0 = blockIdx.z * g.kb_per_split;
  int kb1 = kb0 + g.kb_per_split;
  kb1 = kb1 < g.num_kb ? kb1 : g.num_kb;
  const int nkb = kb1 - kb0;       // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], TC_WORKERS / 32);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();      // nothing above touches global memory: the prologue overlaps the previous kernel's tail
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // whole warp, elected issue (see "Warp-collective issue" above)
    // The segment / column of every 32-column chunk is fixed for the CTA: resolved once, not per k-block (the producer's
    // issue path is on the critical path of the three-stage ring: nine find_seg + TMA issues per k-block measured ~0.45 us).
    // (Requesting the k-blocks beyond the ring into L2 ahead of time with cp.async.bulk.prefetch.tensor was measured and
    // made the kernel SLOWER, 30 -> 38 us on 160 x 160: nine more issues per k-block on this warp, tools/bench_wgrad.py.)
    constexpr int A_CH = TC_BM / 32, B_CH = WG_MAX_BN / 32;
    const CUtensorMap* amap[A_CH];
    const CUtensorMap* bmap[B_CH];
    int acol[A_CH], bcol[B_CH];
#pragma unroll
    for (int c = 0; c < A_CH; ++c) {                // columns beyond the last segment: fully out-of-bounds box -> zeros
      const int o = m0 + 32 * c;
      const int sg = find_seg(g.a_start, g.a_nseg, o);
      amap[c] = &maps.a[sg];
      acol[c] = o - g.a_start[sg];
    }
#pragma unroll
    for (int c = 0; c < B_CH; ++c) {
      const int i = n0 + 32 * c;
      const int sg = find_seg(g.b_start, g.b_nseg, i);
      bmap[c] = &maps.b[sg];
      bcol[c] = i - g.b_start[sg];
    }
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
      mbar_wait(&empty_bar[s], ph ^ 1u);
      __syncwarp();
      mbar_expect_tx_w(&full_bar[s], a_bytes + b_bytes);
      const int row = (kb0 + it) * WG_KB;
      unsigned char* dst = stage_a(s);
#pragma unroll
      for (int c = 0; c < A_CH; ++c) tma_load_2d_w(dst + c * WG_CHUNK_BYTES, amap[c], &full_bar[s], acol[c], row);
      dst = stage_bhi(s);
#pragma unroll
      for (int c = 0; c < B_CH; ++c)
        if (c < b_chunks) tma_load_2d_w(dst + c * WG_CHUNK_BYTES, bmap[c], &full_bar[s], bcol[c], row);
    }
  } else if (warp == 1) {
    // D = F32, A = B = TF32, A K-major from TMEM, B MN-major (bit 16) from shared memory, M = 128, N = BN
    const uint32_t idesc = idesc_tf32(BN) | (1u << 16);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
      mbar_wait(&full_bar[s], ph);
      mbar_wait(&split_bar[s], ph);
      tc_fence_after();
      __syncwarp();
      const uint32_t a_hi = tmem_base + static_cast<uint32_t>(WG_A_TMEM + s * 2 * WG_KB);
      // 8 contraction rows per k-step = two 4-row swizzle atoms = 1024 bytes (64 descriptor address units); the three
      // small terms go to their own accumulator (see gemm_tc_kernel)
      umma_kblock_wg_w(tmem_base, tmem_base + WG_ACC2, a_hi, a_hi + WG_KB, smem_desc_mn_sw128_32b(smem_u32(stage_bhi(s))),
                       smem_desc_mn_sw128_32b(smem_u32(stage_blo(s))), idesc, it != 0 ? 1u : 0u, &empty_bar[s]);
    }
    umma_commit_w(&acc_bar);
  } else {
    const int t = threadIdx.x - 64;
    const bool a_side = warp < 6;                      // warps 2..5: A-splitters, 6..9: B-splitters
    const bool want_db = g.db != nullptr && blockIdx.y == 0;
    float colsum = 0.f;                                // fused bias gradient: db[o] = sum_m G[m, o]
    int s = 0;
    uint32_t ph = 0;
    if (a_side) {
      const int q = warp & 3;                          // TMEM lane quarter of this warp = chunk of the G tile
      const float* chunk0 = reinterpret_cast<const float*>(stage_a(0)) + q * (WG_CHUNK_BYTES / 4);
      for (int it = 0; it < nkb; ++it, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
        mbar_wait(&full_bar[s], ph);
        const float* chunk = chunk0 + static_cast<size_t>(s) * (stage_bytes / 4);
        const uint32_t slot = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(WG_A_TMEM + s * 2 * WG_KB);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            // logical column `lane` of row r: 16-byte slot 2 * ((lane / 8) ^ (r & 3)) + ((lane / 4) & 1)   (Swizzle<2,5,2>);
            // the 32 lanes of a warp hit 32 different banks
            const int r = 16 * half + rr;
            const float v = chunk[r * 32 + (2 * ((lane >> 3) ^ (r & 3)) + ((lane >> 2) & 1)) * 4 + (lane & 3)];
            colsum += v;
            float h, o;
            split_tf32(v, h, o);
            hi[rr] = __float_as_uint(h);
            lo[rr] = __float_as_uint(o);
          }
          tmem_st16(slot + 16 * half, hi);
          tmem_st16(slot + WG_KB + 16 * half, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
      }
      if (want_db) {
        const int64_t o = static_cast<int64_t>(m0) + q * 32 + lane;
        if (o < g.e.M) g.db[static_cast<int64_t>(blockIdx.z) * g.e.M + o] = colsum;
      }
    } else {
      const int tb = t - 128;
      const int n4 = static_cast<int>(b_bytes / 16);
      for (int it = 0; it < nkb; ++it, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
        mbar_wait(&full_bar[s], ph);
        float4* h4 = reinterpret_cast<float4*>(stage_bhi(s));
        float4* l4 = reinterpret_cast<float4*>(stage_blo(s));
#pragma unroll 5
        for (int i = tb; i < n4; i += 128) {
          const float4 v = h4[i];
          float4 h, o;
          split_tf32(v.x, h.x, o.x);
          split_tf32(v.y, h.y, o.y);
          split_tf32(v.z, h.z, o.z);
          split_tf32(v.w, h.w, o.w);
          h4[i] = h;
          l4[i] = o;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
      }
    }
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 32);     // stage memory is free now
    const EpiCtx cx = epi_ctx(g.e);
    float* ws = g.ws != nullptr ? g.ws + static_cast<int64_t>(blockIdx.z) * g.e.M * g.e.N : nullptr;
    // main accumulator: one hi*hi MMA per k-step, four k-steps per k-block
    tc_epilogue<AX2D_ACT_NONE, AX2D_ACT_NONE, false>(g.e, BN, ws, cx, tmem_base, stg, m0, n0, warp & 3, lane, (warp - 2) >> 2,
                                                     static_cast<uint32_t>(WG_ACC2), tc_acc_scale(4 * nkb));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// ---------------------------------------------------------------------------------------------- weight gradients, 160 rows
// Most weight matrices of the model have 160 output features = one 128-row UMMA tile + a 32-row strip.  gemm_tc_wgrad_kernel
// gives the strip its own CTAs, which run the whole k-loop (all of X again, a full M = 128 MMA per term) for a quarter of a
// tile of output: half of the CTA-k-blocks of a 160 x 160 launch.  Here ONE CTA owns all 160 rows of its column tile:
//   * rows 0..127: as before -- G columns 0..127 split into the TMEM A ring by warps 2..5, MMAs with the A operand from
//     tensor memory;
//   * rows 128..159 are computed TRANSPOSED, strip^T[i, o] = sum_m X[m, i] G[m, 128 + o]: the M side of the MMA is the
//     column tile of X (already in shared memory, split into hi / lo by warps 6..9), the N side the strip's G chunk
//     [32 contraction rows x 32 columns] -- one more TMA box per k-block, split in shared memory by warps 2..5 -- so a strip
//     MMA has N = 32 and costs a fifth of a main-tile MMA instead of as much (an M = 128 MMA over the strip's 32 valid rows
//     costs the same as a full tile).  Two such products per term: X columns 0..127 and 128..159 (the second reads past the
//     tile's X chunks into the following shared memory: finite garbage in accumulator lanes the epilogue never stores);
//   * TMEM: [0, 160) accumulator of rows 0..127, [160, 192) / [192, 224) the two transposed strip accumulators, [320, 512)
//     the A ring.  The three terms of a k-step go to ONE accumulator (like the projection kernel: 128 + 64 columns more for
//     separate small-term accumulators do not fit next to the ring), compensated by tc_acc_scale;
//   * the strip's epilogue needs no transpose: lane = column i of X, register j = row 128 + j, so every store instruction
//     of a warp writes 32 consecutive floats of one output row.
// Half the CTAs per split means twice the splits on the same machine: 8 instead of 16 k-blocks per CTA on 160 x 160.
// Where a tile lies inside one column segment its chunks arrive as ONE 3-D TMA box per operand (make_map_chunks) instead of
// one box per 32-column chunk.  Measured (tools/bench_wgrad.py): ~1.15 us per k-block whatever the number of TMA
// instructions (2 or 10) and with or without an L2 prefetch of the k-blocks beyond the ring -- what is left is the MMA
// stream itself: 12 MMAs of N = 160 (0.5 us) + 24 of N = 32, which cost far more than a fifth of a wide one.
constexpr int WGM_ACC_SA = 160, WGM_ACC_SB = 192;
constexpr int WGM_STRIP_BYTES = WG_CHUNK_BYTES;       // one [32 x 32] fp32 box

// One k-block of the merged kernel: per k-step three MMAs into the main tile (A from TMEM, N = BN) and 2 x 3 into the
// transposed strip tiles (A = X from shared memory, B = the strip chunk, N = 32), and the commit, under one election.
__device__ __forceinline__ void umma_kblock_wgm_w(uint32_t t0, uint32_t ts, uint32_t a_hi, uint32_t a_lo, uint64_t dsh, uint64_t dsl,
                                                  uint64_t dbh, uint64_t dbl, uint32_t idesc_ts, uint32_t idesc_s,
                                                  uint32_t acc_first, uint64_t* free_bar) {
  static_assert(WG_KB == 32, "four k-steps of 8 per k-block");
  // %0 main accumulator, %1 strip accumulator A (B = %1 + 32), %2 / %3 TMEM A (hi / lo), %4 / %5 strip chunk (hi / lo),
  // %6 / %7 X (hi / lo), %8 / %9 instruction descriptors, %10 accumulate flag of the first k-step, %11 barrier
#define AX2D_WGM_STEP(AH, AL, SH, SL, BH, BL, BH2, BL2, PF)                                \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [" AL "], " BH ", %8, " PF ";\n\t"         \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [" AH "], " BL ", %8, pt;\n\t"             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [" AH "], " BH ", %8, pt;\n\t"             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], " BL ", " SH ", %9, " PF ";\n\t"           \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], " BH ", " SL ", %9, pt;\n\t"               \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], " BH ", " SH ", %9, pt;\n\t"               \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [tsb], " BL2 ", " SH ", %9, " PF ";\n\t"         \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [tsb], " BH2 ", " SL ", %9, pt;\n\t"             \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [tsb], " BH2 ", " SH ", %9, pt;\n\t"
  asm volatile(
      "{\n\t"
      ".reg .pred e, pf, pt;\n\t"
      ".reg .b32 ah1, al1, ah2, al2, ah3, al3, tsb;\n\t"
      ".reg .b64 sh1, sl1, sh2, sl2, sh3, sl3, bh1, bl1, bh2, bl2, bh3, bl3;\n\t"
      ".reg .b64 ch0, cl0, ch1, cl1, ch2, cl2, ch3, cl3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %10, 0;\n\t"
      "setp.eq.b32 pt, %8, %8;\n\t"
      "add.u32 tsb, %1, 32;\n\t"
      "add.u32 ah1, %2, 8;\n\t add.u32 al1, %3, 8;\n\t add.u64 bh1, %6, 64;\n\t add.u64 bl1, %7, 64;\n\t"
      "add.u32 ah2, %2, 16;\n\t add.u32 al2, %3, 16;\n\t add.u64 bh2, %6, 128;\n\t add.u64 bl2, %7, 128;\n\t"
      "add.u32 ah3, %2, 24;\n\t add.u32 al3, %3, 24;\n\t add.u64 bh3, %6, 192;\n\t add.u64 bl3, %7, 192;\n\t"
      "add.u64 sh1, %4, 64;\n\t add.u64 sl1, %5, 64;\n\t add.u64 sh2, %4, 128;\n\t add.u64 sl2, %5, 128;\n\t"
      "add.u64 sh3, %4, 192;\n\t add.u64 sl3, %5, 192;\n\t"
      // X columns 128 ..: four 4 KB chunks further on (descriptor addresses are in 16-byte units)
      "add.u64 ch0, %6, 1024;\n\t add.u64 cl0, %7, 1024;\n\t add.u64 ch1, bh1, 1024;\n\t add.u64 cl1, bl1, 1024;\n\t"
      "add.u64 ch2, bh2, 1024;\n\t add.u64 cl2, bl2, 1024;\n\t add.u64 ch3, bh3, 1024;\n\t add.u64 cl3, bl3, 1024;\n\t"
      AX2D_WGM_STEP("%2", "%3", "%4", "%5", "%6", "%7", "ch0", "cl0", "pf")
      AX2D_WGM_STEP("ah1", "al1", "sh1", "sl1", "bh1", "bl1", "ch1", "cl1", "pt")
      AX2D_WGM_STEP("ah2", "al2", "sh2", "sl2", "bh2", "bl2", "ch2", "cl2", "pt")
      AX2D_WGM_STEP("ah3", "al3", "sh3", "sl3", "bh3", "bl3", "ch3", "cl3", "pt")
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%11];\n\t"
      "}\n" ::"r"(t0),
      "r"(ts), "r"(a_hi), "r"(a_lo), "l"(dsh), "l"(dsl), "l"(dbh), "l"(dbl), "r"(idesc_ts), "r"(idesc_s), "r"(acc_first),
      "r"(smem_u32(free_bar))
      : "memory");
#undef AX2D_WGM_STEP
}

__global__ void __launch_bounds__(TC_WG_THREADS, 1) gemm_tc_wgrad160_kernel(const __grid_constant__ WgMaps maps,
                                                                         const __grid_constant__ WgArgs g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[WG_MAX_STAGES], split_bar[WG_MAX_STAGES], empty_bar[WG_MAX_STAGES], acc_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_strip_sum[4][32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = g.BN, S = g.stages;
  constexpr int a_chunks = TC_BM / 32;
  const int b_chunks = BN / 32;
  constexpr uint32_t a_bytes = a_chunks * WG_CHUNK_BYTES;
  const uint32_t b_bytes = static_cast<uint32_t>(b_chunks) * WG_CHUNK_BYTES;
  // stage: [raw G chunks 0..3 | strip hi | X hi chunks | strip lo | X lo chunks]
  const uint32_t stage_bytes = a_bytes + 2 * WGM_STRIP_BYTES + 2 * b_bytes;
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  auto stage_a = [&](int s) { return smem + static_cast<size_t>(s) * stage_bytes; };
  auto stage_shi = [&](int s) { return stage_a(s) + a_bytes; };
  auto stage_bhi = [&](int s) { return stage_a(s) + a_bytes + WGM_STRIP_BYTES; };
  auto stage_slo = [&](int s) { return stage_a(s) + a_bytes + WGM_STRIP_BYTES + b_bytes; };
  auto stage_blo = [&](int s) { return stage_a(s) + a_bytes + 2 * WGM_STRIP_BYTES + b_bytes; };

  const int n0 = blockIdx.y * BN;
  const int kb0 = blockIdx.z * g.kb_per_split;
  int kb1 = kb0 + g.kb_per_split;
  kb1 = kb1 < g.num_kb ? kb1 : g.num_kb;
  const int nkb = kb1 - kb0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&split_bar[s], TC_WORKERS / 32);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_enter();      // nothing above touches global memory: the prologue overlaps the previous kernel's tail
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    constexpr int A_CH = a_chunks + 1, B_CH = WG_MAX_BN / 32;          // four chunks of the main tile + the strip
    const CUtensorMap* amap[A_CH];
    const CUtensorMap* bmap[B_CH];
    int acol[A_CH], bcol[B_CH];
#pragma unroll
    for (int c = 0; c < A_CH; ++c) {
      const int o = 32 * c;
      const int sg = find_seg(g.a_start, g.a_nseg, o);
      amap[c] = &maps.a[sg];
      acol[c] = o - g.a_start[sg];
    }
#pragma unroll
    for (int c = 0; c < B_CH; ++c) {
      const int i = n0 + 32 * c;
      const int sg = find_seg(g.b_start, g.b_nseg, i);
      bmap[c] = &maps.b[sg];
      bcol[c] = i - g.b_start[sg];
    }
    // this CTA's X tile through the 3-D view of its segment: only if it lies inside ONE segment
    const int b_sg0 = find_seg(g.b_start, g.b_nseg, n0);
    const bool b3_ok = g.b3_on != 0 && find_seg(g.b_start, g.b_nseg, n0 + BN - 1) == b_sg0 &&
                       n0 + BN <= g.b_start[b_sg0 + 1];
    const int b3_chunk0 = (n0 - g.b_start[b_sg0]) >> 5;
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
      mbar_wait(&empty_bar[s], ph ^ 1u);
      __syncwarp();
      mbar_expect_tx_w(&full_bar[s], a_bytes + WGM_STRIP_BYTES + b_bytes);
      const int row = (kb0 + it) * WG_KB;
      unsigned char* dst = stage_a(s);
      if (g.a3_chunk0 >= 0) {
        // the four chunks of the main tile and the strip chunk are contiguous in shared memory: one 3-D box
        tma_load_3d_w(dst, &maps.a3, &full_bar[s], 0, row, g.a3_chunk0);
      } else {
#pragma unroll
        for (int c = 0; c < a_chunks; ++c) tma_load_2d_w(dst + c * WG_CHUNK_BYTES, amap[c], &full_bar[s], acol[c], row);
        tma_load_2d_w(stage_shi(s), amap[a_chunks], &full_bar[s], acol[a_chunks], row);
      }
      dst = stage_bhi(s);
      if (b3_ok) {
        tma_load_3d_w(dst, &maps.b3[b_sg0], &full_bar[s], 0, row, b3_chunk0);
      } else {
#pragma unroll
        for (int c = 0; c < B_CH; ++c)
          if (c < b_chunks) tma_load_2d_w(dst + c * WG_CHUNK_BYTES, bmap[c], &full_bar[s], bcol[c], row);
      }
    }
  } else if (warp == 1) {
    // main tile: A K-major from TMEM, B MN-major (bit 16), N = BN; strip tiles: A = X, MN-major from shared memory (bit 15),
    // B = the strip chunk, MN-major, N = 32
    const uint32_t idesc_ts = idesc_tf32(BN) | (1u << 16);
    const uint32_t idesc_s = idesc_tf32(32) | (1u << 15) | (1u << 16);
    int s = 0;
    uint32_t ph = 0;
    for (int it = 0; it < nkb; ++it, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
      mbar_wait(&full_bar[s], ph);
      mbar_wait(&split_bar[s], ph);
      tc_fence_after();
      __syncwarp();
      const uint32_t a_hi = tmem_base + static_cast<uint32_t>(WG_A_TMEM + s * 2 * WG_KB);
      umma_kblock_wgm_w(tmem_base, tmem_base + WGM_ACC_SA, a_hi, a_hi + WG_KB, smem_desc_mn_sw128_32b(smem_u32(stage_shi(s))),
                        smem_desc_mn_sw128_32b(smem_u32(stage_slo(s))), smem_desc_mn_sw128_32b(smem_u32(stage_bhi(s))),
                        smem_desc_mn_sw128_32b(smem_u32(stage_blo(s))), idesc_ts, idesc_s, it != 0 ? 1u : 0u, &empty_bar[s]);
    }
    umma_commit_w(&acc_bar);
  } else {
    const int t = threadIdx.x - 64;
    const bool a_side = warp < 6;
    const bool want_db = g.db != nullptr && blockIdx.y == 0;
    float colsum = 0.f, strip_sum = 0.f;
    int s = 0;
    uint32_t ph = 0;
    if (a_side) {
      const int q = warp & 3;
      const float* chunk0 = reinterpret_cast<const float*>(stage_a(0)) + q * (WG_CHUNK_BYTES / 4);
      for (int it = 0; it < nkb; ++it, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
        mbar_wait(&full_bar[s], ph);
        const float* chunk = chunk0 + static_cast<size_t>(s) * (stage_bytes / 4);
        const uint32_t slot = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(WG_A_TMEM + s * 2 * WG_KB);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int rr = 0; rr < 16; ++rr) {
            const int r = 16 * half + rr;
            const float v = chunk[r * 32 + (2 * ((lane >> 3) ^ (r & 3)) + ((lane >> 2) & 1)) * 4 + (lane & 3)];
            colsum += v;
            float h, o;
            split_tf32(v, h, o);
            hi[rr] = __float_as_uint(h);
            lo[rr] = __float_as_uint(o);
          }
          tmem_st16(slot + 16 * half, hi);
          tmem_st16(slot + WG_KB + 16 * half, lo);
        }
        // the strip chunk: this warp splits contraction rows 8 q .. 8 q + 7 of it in shared memory (lane = column)
        float* shi = reinterpret_cast<float*>(stage_shi(s));
        float* slo = reinterpret_cast<float*>(stage_slo(s));
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
          const int r = 8 * q + rr;
          const int idx = r * 32 + (2 * ((lane >> 3) ^ (r & 3)) + ((lane >> 2) & 1)) * 4 + (lane & 3);
          const float v = shi[idx];
          strip_sum += v;
          float h, o;
          split_tf32(v, h, o);
          shi[idx] = h;
          slo[idx] = o;
        }
        fence_proxy_async_smem();
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
      }
      if (want_db) {
        const int64_t o = static_cast<int64_t>(q) * 32 + lane;
        g.db[static_cast<int64_t>(blockIdx.z) * g.e.M + o] = colsum;
      }
      s_strip_sum[q][lane] = strip_sum;
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (want_db && q == 0 && 128 + lane < g.e.M)
        g.db[static_cast<int64_t>(blockIdx.z) * g.e.M + 128 + lane] =
            ((s_strip_sum[0][lane] + s_strip_sum[1][lane]) + s_strip_sum[2][lane]) + s_strip_sum[3][lane];
    } else {
      const int tb = t - 128;
      const int n4 = static_cast<int>(b_bytes / 16);
      for (int it = 0; it < nkb; ++it, ph ^= (s + 1 == S ? 1u : 0u), s = (s + 1 == S ? 0 : s + 1)) {
        mbar_wait(&full_bar[s], ph);
        float4* h4 = reinterpret_cast<float4*>(stage_bhi(s));
        float4* l4 = reinterpret_cast<float4*>(stage_blo(s));
#pragma unroll 5
        for (int i = tb; i < n4; i += 128) {
          const float4 v = h4[i];
          float4 h, o;
          split_tf32(v.x, h.x, o.x);
          split_tf32(v.y, h.y, o.y);
          split_tf32(v.z, h.z, o.z);
          split_tf32(v.w, h.w, o.w);
          h4[i] = h;
          l4[i] = o;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&split_bar[s]);
      }
    }
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 32);     // stage memory is free now
    const EpiCtx cx = epi_ctx(g.e);
    float* ws = g.ws != nullptr ? g.ws + static_cast<int64_t>(blockIdx.z) * g.e.M * g.e.N : nullptr;
    const float scale = tc_acc_scale(4 * nkb * 3);                 // three terms per k-step into one accumulator
    tc_epilogue<AX2D_ACT_NONE, AX2D_ACT_NONE, false>(g.e, BN, ws, cx, tmem_base, stg, 0, n0, warp & 3, lane, (warp - 2) >> 2, 0u,
                                                     scale);
    // the strip, transposed in tensor memory: lane = X column i, register j = output row 128 + j
    {
      const int q = warp & 3, second = (warp - 2) >> 2;            // warps 2..5: X columns 0..127, warp 8: columns 128..159
      if (second == 0 || q == 0) {
        const int i = second == 0 ? q * 32 + lane : TC_BM + lane;
        uint32_t r[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(second == 0 ? WGM_ACC_SA : WGM_ACC_SB), r);
        const int64_t N = g.e.N;
        const int rows = static_cast<int>(g.e.M) - TC_BM;          // valid strip rows (<= 32)
        if (i < BN && n0 + i < N) {
          float* dst = ws + static_cast<int64_t>(TC_BM) * N + n0 + i;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < rows) dst[static_cast<int64_t>(j) * N] = __uint_as_float(r[j]) * scale;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// elementwise (hi, lo) split of a weight matrix, optionally transposed: out[r, c] = split(w[c, r]) if transpose.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ w, int64_t ldw, int rows, int cols,
                                                         int transpose, float* __restrict__ hi, float* __restrict__ lo,
                                                         int64_t ldo) {
  pdl_enter();
  // out is [rows, cols]; source is w[rows, cols] or, transposed, w[cols, rows]
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  if (!transpose) {
    for (int r = by + ty; r < by + 32 && r < rows; r += 8) {
      const int c = bx + tx;
      if (c < cols) {
        float h, l;
        split_tf32(w[static_cast<int64_t>(r) * ldw + c], h, l);
        hi[static_cast<int64_t>(r) * ldo + c] = h;
        lo[static_cast<int64_t>(r) * ldo + c] = l;
      }
    }
    return;
  }
  // transposed: read w[c, r] coalesced along r (source columns), write out[r, c] coalesced along c
  for (int c = bx + ty; c < bx + 32; c += 8) {
    const int r = by + tx;
    tile[c - bx][tx] = (c < cols && r < rows) ? w[static_cast<int64_t>(c) * ldw + r] : 0.f;
  }
  __syncthreads();
  for (int r = by + ty; r < by + 32 && r < rows; r += 8) {
    const int c = bx + tx;
    if (c < cols) {
      float h, l;
      split_tf32(tile[tx][r - by], h, l);
      hi[static_cast<int64_t>(r) * ldo + c] = h;
      lo[static_cast<int64_t>(r) * ldo + c] = l;
    }
  }
}

// ---------------------------------------------------------------------------------------------- host
}  // namespace ax2d

using namespace ax2d;

static unsigned long long* g_tc_dbg = nullptr;
// development aid (not part of include/ax2d.h): 8 x u64 device buffer receiving %globaltimer stamps of CTA (0,0) of
// every following ax2d_gemm_tc launch: start, setup done, first tile landed, first tile split, MMAs issued,
// accumulator ready, epilogue done, CTA done.  nullptr switches it off.
extern "C" void ax2d_debug_timing(unsigned long long* buf) { g_tc_dbg = buf; }


extern "C" int ax2d_split_tf32(const float* w, int64_t ldw, int rows, int cols, int transpose, float* hi, float* lo,
                               int64_t ldo, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(w != nullptr && hi != nullptr && lo != nullptr && rows > 0 && cols > 0 && ldo >= cols,
                 "ax2d_split_tf32: bad arguments");
  dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>((rows + 31) / 32));
  launch_k(split_tf32_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), w, ldw, rows, cols, transpose, hi, lo, ldo);
  return launch_status("ax2d_split_tf32");
}

extern "C" int ax2d_gemm_tc_supported(const ax2d_cmat* a, int64_t M, int64_t N, int64_t K) {
  if (a == nullptr || M < 1 || N < 4 || N % 4 != 0 || K < TC_BK || K % TC_BK != 0) return 0;
  if (a->n_seg < 1 || a->n_seg > AX2D_MAX_SEG) return 0;
  for (int s = 0; s < a->n_seg; ++s)
    if (a->width[s] <= 0 || a->width[s] % TC_BK != 0 || a->ld[s] % 4 != 0 || (reinterpret_cast<uintptr_t>(a->ptr[s]) & 15u)) return 0;
  return 1;
}

extern "C" int ax2d_gemm_tc(const ax2d_cmat* a, const float* b_hi, const float* b_lo, int64_t ldb, const ax2d_mat* c,
                            int64_t M, int64_t N, int64_t K, const ax2d_epilogue* ep, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(a != nullptr && b_hi != nullptr && b_lo != nullptr && c != nullptr, "ax2d_gemm_tc: null operand");
  AX2D_CHECK_ARG(ax2d_gemm_tc_supported(a, M, N, K), "ax2d_gemm_tc: unsupported shape M=%lld N=%lld K=%lld (see ax2d.h)",
                 (long long)M, (long long)N, (long long)K);
  AX2D_CHECK_ARG(ldb % 4 == 0 && ldb >= K, "ax2d_gemm_tc: bad ldb");
  AX2D_CHECK_ALIGN(b_hi);
  AX2D_CHECK_ALIGN(b_lo);
  TcArgs g;
  memset(&g, 0, sizeof(g));
  g.dbg = g_tc_dbg;
  TcMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if ((rc = to_out(c, &g.e.c, N, "C", false)) != AX2D_OK) return rc;
  if ((rc = fill_epilogue(ep, M, N, &g.e)) != AX2D_OK) return rc;
  // tile shape: column tiles of at most 192 accumulator columns (two accumulator buffers + the A ring share the 512
  // TMEM columns), as few as possible without padding N by more than ~6 % -- unless the tiles cannot fill the machine
  // (the [B, 512] head products: 16 row tiles), in which case N is cut into narrower tiles until there is about one
  // per SM (narrow tiles also get the second accumulator)
  const int64_t m_tiles = (M + TC_BM - 1) / TC_BM;
  int n_tiles = static_cast<int>((N + TC_MAX_BN - 1) / TC_MAX_BN);
  for (int extra = 0; extra < 2; ++extra) {
    const int bn = static_cast<int>((N + n_tiles - 1) / n_tiles + 31) / 32 * 32;
    if (static_cast<int64_t>(bn) * n_tiles * 16 <= N * 17) break;
    ++n_tiles;
  }
  if (m_tiles * n_tiles < kNumSMs) {
    const int want = static_cast<int>((kNumSMs + m_tiles - 1) / m_tiles);
    const int most = static_cast<int>((N + 31) / 32);
    n_tiles = want < most ? want : most;
  }
  int BN = static_cast<int>((N + n_tiles - 1) / n_tiles);
  BN = (BN + 31) / 32 * 32;
  n_tiles = static_cast<int>((N + BN - 1) / BN);
  g.BN = BN;
  g.n_tiles = n_tiles;
  AX2D_CHECK_ARG(m_tiles * n_tiles < (1ll << 30), "ax2d_gemm_tc: too many tiles");
  g.total_tiles = static_cast<int>(m_tiles * n_tiles);
  g.tmem_cols = 512;
  // k-blocks of 32 fp32 (128-byte swizzle rows) when every A segment allows it, of 16 otherwise: every k-block costs
  // the producer -> splitter -> MMA chain a fixed ~0.2 us on top of its MMAs, which dominated the [B, 512] head products
  // (measured: 17.3 vs 19.9 us on [2048, 512] x [512, 512]).  Wide k-blocks halve the number of ring slots, which costs
  // the big products more than the halved overhead gains (107 -> 130 us on 544 -> 512), so they are used for the
  // narrow tiles of the small-M products only.
  bool wide = K % TC_BK_WIDE == 0 && BN <= 64;
  for (int s = 0; s < a->n_seg; ++s) wide = wide && a->width[s] % TC_BK_WIDE == 0;
  const int BK = wide ? TC_BK_WIDE : TC_BK;
  // TMEM layouts (accumulator stride, small-term accumulator offset, first column of the A ring; a slot of the ring
  // is 2 BK columns).  Only layouts that were run on B200 are used: a ring of six 32-column slots starting at column
  // 256 or 288 hung the kernel in testing (cause unknown; the same ring with <= 5 slots, or 8 slots under 64-wide
  // tiles, or starting at 320, runs).
  if (BN <= 64 && wide) {
    g.acc_stride = 128; g.acc2 = 64; g.a_base = 256;       // 4 slots of 64
  } else if (BN <= 96 && !wide) {
    g.acc_stride = 192; g.acc2 = 96; g.a_base = 384;       // 4 slots of 32
  } else if (BN <= 160) {
    g.acc_stride = 160; g.acc2 = 0; g.a_base = 320;        // 6 slots of 32 / 3 slots of 64
  } else {
    g.acc_stride = 192; g.acc2 = 0; g.a_base = 384;        // 4 slots of 32 / 2 slots of 64
  }
  int acc = 0, kb = 0;
  g.n_seg = a->n_seg;
  const CUtensorMapSwizzle swz = wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  for (int s = 0; s < a->n_seg; ++s) {
    g.seg_kb_start[s] = kb;
    if ((rc = make_map(&maps.a[s], a->ptr[s], a->width[s], M, a->ld[s], TC_BM, BK, swz)) != AX2D_OK) return rc;
    kb += a->width[s] / BK;
    acc += a->width[s];
  }
  for (int s = a->n_seg; s <= AX2D_MAX_SEG; ++s) g.seg_kb_start[s] = kb;
  AX2D_CHECK_ARG(acc == K, "ax2d_gemm_tc: A segments cover %d columns, expected %lld", acc, (long long)K);
  g.num_kb = kb;
  // MMAs accumulated into the main accumulator per output: K / 8 k-steps x (hi*hi only | all split terms)
  g.acc_scale = tc_acc_scale(static_cast<int>(K / 8) * (g.acc2 != 0 ? 1 : AX2D_TC_TERMS));
  if ((rc = make_map(&maps.b_hi, b_hi, K, N, ldb, BN, BK, swz)) != AX2D_OK) return rc;
  if ((rc = make_map(&maps.b_lo, b_lo, K, N, ldb, BN, BK, swz)) != AX2D_OK) return rc;
  const size_t stage_bytes = static_cast<size_t>(TC_BM + 2 * BN) * BK * 4;
  // one CTA per SM: the stage ring takes what the 227 KB leave after the epilogue's transpose buffers
  const size_t stg_bytes = static_cast<size_t>(TC_EPI_WARPS) * 32 * 32 * 4;
  const size_t budget = 227 * 1024 - 1024 - stg_bytes - 512;         // alignment slack, static barriers
  int stages = static_cast<int>(budget / stage_bytes);
  const int a_slots = (512 - g.a_base) / (2 * BK);                    // TMEM A ring: one slot per stage
  stages = stages > a_slots ? a_slots : stages;
  stages = stages > TC_MAX_STAGES ? TC_MAX_STAGES : stages;
  if (stages < 1) stages = 1;
  g.stages = stages;
  size_t smem = stages * stage_bytes + stg_bytes + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;    // never two CTAs on an SM: each allocates all 512 TMEM columns
  static size_t configured[2] = {0, 0};
  if (smem > configured[wide]) {
    cudaError_t e = wide ? cudaFuncSetAttribute(gemm_tc_kernel<TC_BK_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                         : cudaFuncSetAttribute(gemm_tc_kernel<TC_BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_error("ax2d_gemm_tc: cannot raise the dynamic shared memory limit to %zu: %s", smem, cudaGetErrorString(e));
      return AX2D_ERR_LAUNCH;
    }
    configured[wide] = smem;
  }
  dim3 grid(static_cast<unsigned>(g.total_tiles < kNumSMs ? g.total_tiles : kNumSMs));
  if (wide) launch_k(gemm_tc_kernel<TC_BK_WIDE>, dim3(grid), dim3(TC_THREADS), smem, reinterpret_cast<cudaStream_t>(stream), maps, g);
  else launch_k(gemm_tc_kernel<TC_BK>, dim3(grid), dim3(TC_THREADS), smem, reinterpret_cast<cudaStream_t>(stream), maps, g);
  return launch_status("ax2d_gemm_tc");
}

extern "C" int ax2d_gemm_tc_wgrad_supported(const ax2d_cmat* a, const ax2d_cmat* b, int64_t M, int64_t N, int64_t K) {
  // a: G [K rows, M columns (segmented)], b: X [K rows, N columns (segmented)]
  if (a == nullptr || b == nullptr || M < 32 || N < 32 || M % 4 != 0 || N % 4 != 0 || K < 1) return 0;
  const ax2d_cmat* ops_[2] = {a, b};
  for (const ax2d_cmat* m : ops_) {
    if (m->n_seg < 1 || m->n_seg > AX2D_MAX_SEG) return 0;
    for (int s = 0; s < m->n_seg; ++s)
      if (m->width[s] <= 0 || m->width[s] % 32 != 0 || m->ld[s] % 4 != 0 || (reinterpret_cast<uintptr_t>(m->ptr[s]) & 15u)) return 0;
  }
  return 1;
}

// How many ways to cut the contraction: at least `by_chain` (accumulation chains of at most 1024 rows), and among the
// candidates up to 4x that the one that minimises waves x (k-blocks per CTA + fixed cost): 6 tiles x 37 splits = 222
// CTAs would run 1.5 waves on 148 SMs, 6 x 49 = 294 runs two full ones with 24 instead of 32 k-blocks each.
static int wgrad_split(int64_t tiles, int64_t num_kb) {
  const int64_t by_chain = (num_kb + WG_MAX_KB_PER_SPLIT - 1) / WG_MAX_KB_PER_SPLIT;
  int64_t hi = by_chain * 4 > num_kb ? num_kb : by_chain * 4;
  const int64_t fill = (kNumSMs + tiles - 1) / tiles;       // enough CTAs for one wave
  if (hi < fill) hi = fill < num_kb ? fill : num_kb;
  int64_t best = by_chain > num_kb ? num_kb : by_chain, best_cost = -1;
  for (int64_t sp = best; sp <= hi; ++sp) {
    const int64_t per = (num_kb + sp - 1) / sp;
    const int64_t real = (num_kb + per - 1) / per;                 // no empty split
    const int64_t waves = (tiles * real + kNumSMs - 1) / kNumSMs;
    const int64_t cost = waves * (per + 5) * 64 + real;            // prologue + epilogue ~ 5 k-blocks; ties: fewer partials
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = real; }
  }
  return static_cast<int>(best);
}

// 128 < M <= 160 output features: one CTA per column tile owns all rows (gemm_tc_wgrad160_kernel) instead of a full tile
// plus a strip tile.  Needs the three-term split (one accumulator per tile, no room for a fourth term's bookkeeping).
static int g_wg_merged = 1, g_wg_box3 = 1;
// development aid (not part of include/ax2d.h): 0 = per-chunk 2-D boxes in the merged kernel
extern "C" void ax2d_debug_wgrad_box3(int on) { g_wg_box3 = on; }
// development aid (not part of include/ax2d.h): 0 = always the two-tile kernel
extern "C" void ax2d_debug_wgrad_merged(int on) { g_wg_merged = on; }
// (only with several splits, i.e. long contractions: its strip epilogue writes partial tiles)
static bool wgrad_merged(int64_t M, int64_t K) {
  return g_wg_merged != 0 && AX2D_TC_TERMS == 3 && M > TC_BM && M <= TC_BM + 32 && K >= 64 * WG_KB;
}
static int64_t wgrad_m_tiles(int64_t M, int64_t K) { return wgrad_merged(M, K) ? 1 : (M + TC_BM - 1) / TC_BM; }

extern "C" int ax2d_gemm_tc_wgrad_splits(int64_t M, int64_t N, int64_t K) {
  const int64_t n_tiles = (N + WG_MAX_BN - 1) / WG_MAX_BN;
  return wgrad_split(wgrad_m_tiles(M, K) * n_tiles, (K + WG_KB - 1) / WG_KB);
}

extern "C" int64_t ax2d_gemm_tc_wgrad_workspace(int64_t M, int64_t N, int64_t K) {
  const int64_t n_tiles = (N + WG_MAX_BN - 1) / WG_MAX_BN;
  const int64_t tiles = wgrad_m_tiles(M, K) * n_tiles;
  const int64_t num_kb = (K + WG_KB - 1) / WG_KB;
  const int64_t split = wgrad_split(tiles, num_kb);
  return split > 1 ? split * (M * N + M) * 4 : 0;       // partial tiles + partial bias-gradient vectors
}

// C[M,N] (+)= A^T B with A = a [K, M], B = b [K, N] (both column-segmented, contraction over the K rows).
extern "C" int ax2d_gemm_tc_wgrad(const ax2d_cmat* a, const ax2d_cmat* b, const ax2d_mat* c, int64_t M, int64_t N, int64_t K,
                                  int accumulate, float* bias_grad, void* workspace, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(c != nullptr && ax2d_gemm_tc_wgrad_supported(a, b, M, N, K),
                 "ax2d_gemm_tc_wgrad: unsupported operands M=%lld N=%lld K=%lld (segment widths must be multiples of 32)",
                 (long long)M, (long long)N, (long long)K);
  WgArgs g;
  memset(&g, 0, sizeof(g));
  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if ((rc = to_out(c, &g.e.c, N, "C", false)) != AX2D_OK) return rc;
  ax2d_epilogue ep;
  memset(&ep, 0, sizeof(ep));
  ep.accumulate = accumulate;
  if ((rc = fill_epilogue(&ep, M, N, &g.e)) != AX2D_OK) return rc;
  int acc = 0;
  g.a_nseg = a->n_seg;
  for (int s = 0; s < a->n_seg; ++s) {
    g.a_start[s] = acc;
    if ((rc = make_map(&maps.a[s], a->ptr[s], a->width[s], K, a->ld[s], WG_KB, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != AX2D_OK) return rc;
    acc += a->width[s];
  }
  for (int s = a->n_seg; s <= AX2D_MAX_SEG; ++s) g.a_start[s] = acc;
  AX2D_CHECK_ARG(acc == M, "ax2d_gemm_tc_wgrad: A segments cover %d columns, expected %lld", acc, (long long)M);
  acc = 0;
  g.b_nseg = b->n_seg;
  for (int s = 0; s < b->n_seg; ++s) {
    g.b_start[s] = acc;
    if ((rc = make_map(&maps.b[s], b->ptr[s], b->width[s], K, b->ld[s], WG_KB, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) != AX2D_OK) return rc;
    acc += b->width[s];
  }
  for (int s = b->n_seg; s <= AX2D_MAX_SEG; ++s) g.b_start[s] = acc;
  AX2D_CHECK_ARG(acc == N, "ax2d_gemm_tc_wgrad: B segments cover %d columns, expected %lld", acc, (long long)N);
  const int n_tiles = static_cast<int>((N + WG_MAX_BN - 1) / WG_MAX_BN);
  int BN = static_cast<int>((N + n_tiles - 1) / n_tiles);
  BN = (BN + 31) / 32 * 32;
  g.BN = BN;
  g.tmem_cols = 512;        // one CTA per SM: main + small-term accumulator + the A ring
  g.acc2 = WG_ACC2;
  const bool merged = wgrad_merged(M, K);
  g.a3_chunk0 = -1;
  g.b3_on = 0;
  if (merged && g_wg_box3 != 0) {
    // one 3-D box per operand and k-block where the tile lies inside one segment (see make_map_chunks)
    if (a->width[0] >= TC_BM + 32) {
      if ((rc = make_map_chunks(&maps.a3, a->ptr[0], a->width[0], K, a->ld[0], WG_KB, TC_BM / 32 + 1)) != AX2D_OK) return rc;
      g.a3_chunk0 = 0;
    }
    for (int s = 0; s < b->n_seg; ++s)
      if ((rc = make_map_chunks(&maps.b3[s], b->ptr[s], b->width[s], K, b->ld[s], WG_KB, BN / 32)) != AX2D_OK) return rc;
    g.b3_on = 1;
  }
  const int m_tiles = static_cast<int>(wgrad_m_tiles(M, K));
  g.num_kb = static_cast<int>((K + WG_KB - 1) / WG_KB);
  int split = wgrad_split(static_cast<int64_t>(m_tiles) * n_tiles, g.num_kb);
  g.kb_per_split = (g.num_kb + split - 1) / split;
  split = (g.num_kb + g.kb_per_split - 1) / g.kb_per_split;      // no empty split
  g.e.accumulate = (split == 1 && accumulate == 1) ? 1 : 0;      // partial tiles are plain stores; the reduce accumulates
  if (split > 1) {
    AX2D_CHECK_ARG(workspace != nullptr, "ax2d_gemm_tc_wgrad: workspace required (ax2d_gemm_tc_wgrad_workspace)");
    AX2D_CHECK_ALIGN(workspace);
    g.ws = static_cast<float*>(workspace);
    g.db = bias_grad != nullptr ? g.ws + static_cast<int64_t>(split) * M * N : nullptr;
  } else {
    g.db = bias_grad;
  }
  const size_t stage_bytes = static_cast<size_t>(TC_BM / 32 + (merged ? 2 : 0) + 2 * (BN / 32)) * WG_CHUNK_BYTES;
  int stages = static_cast<int>((224 * 1024) / stage_bytes);
  stages = stages > WG_MAX_STAGES ? WG_MAX_STAGES : stages;
  stages = stages > g.kb_per_split ? g.kb_per_split : stages;
  if (stages < 1) stages = 1;
  g.stages = stages;
  size_t smem = stages * stage_bytes;
  const size_t epi_bytes = 8 * 32 * 32 * 4;        // epilogue transpose staging
  if (smem < epi_bytes) smem = epi_bytes;
  if (smem < 120 * 1024) smem = 120 * 1024;        // never two CTAs on an SM: each allocates all 512 TMEM columns
  smem += 1024;
  if (merged) smem += static_cast<size_t>(8 - BN / 32) * WG_CHUNK_BYTES;   // the second strip product reads X chunks 4..7:
                                                                           // up to here past the last stage's X lo
  static size_t configured[2] = {0, 0};
  if (smem > configured[merged]) {
    cudaError_t e = merged ? cudaFuncSetAttribute(gemm_tc_wgrad160_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                           : cudaFuncSetAttribute(gemm_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_error("ax2d_gemm_tc_wgrad: cannot raise the dynamic shared memory limit to %zu: %s", smem, cudaGetErrorString(e));
      return AX2D_ERR_LAUNCH;
    }
    configured[merged] = smem;
  }
  AX2D_CHECK_ARG(accumulate != 2 || split > 1, "ax2d_gemm_tc_wgrad: accumulate == 2 (leave the partials) needs more than one split "
                                              "(ax2d_gemm_tc_wgrad_splits)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>(m_tiles), static_cast<unsigned>(n_tiles), static_cast<unsigned>(split));
  if (merged) launch_k(gemm_tc_wgrad160_kernel, dim3(grid), dim3(TC_WG_THREADS), smem, st, maps, g);
  else launch_k(gemm_tc_wgrad_kernel, dim3(grid), dim3(TC_WG_THREADS), smem, st, maps, g);
  rc = launch_status("ax2d_gemm_tc_wgrad");
  if (rc != AX2D_OK || split == 1 || accumulate == 2) return rc;
  return splitk_reduce(g.ws, split, M, N, g.e.c, accumulate == 1, g.db, bias_grad, st);
}
