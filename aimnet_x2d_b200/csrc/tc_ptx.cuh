// tcgen05 / TMEM / TMA building blocks shared by the tensor-core kernels (gemm_tc.cu: 3xTF32, gemm_bf16.cu: bf16).
#pragma once
#include <cuda.h>

#include "gemm_common.cuh"

namespace ax2d {

constexpr int TC_BM = 128;
constexpr int TC_BK = 16;                        // fp32 per 64-byte swizzle row (k-block of the generic instance)
constexpr int TC_BK_WIDE = 32;                   // fp32 per 128-byte swizzle row (instance for K segments % 32 == 0)
constexpr int TC_SPLIT_WARPS = 4;                // warps 2..5
constexpr int TC_EPI_WARPS = 12;                 // warps 6..17: three per TMEM lane quarter
constexpr int TC_THREADS = 32 * (2 + TC_SPLIT_WARPS + TC_EPI_WARPS);   // warp 0 TMA, warp 1 MMA + TMEM
constexpr int TC_WORKERS = 256;                  // weight-gradient kernel: splitter / epilogue threads
constexpr int TC_WG_THREADS = 64 + TC_WORKERS;
constexpr int TC_MAX_STAGES = 6;
#ifndef AX2D_TC_TERMS
#define AX2D_TC_TERMS 3     // 3: classic 3xTF32 (drops a_lo*b_lo, 2^-22 relative per product); 4: keeps it.  Measured on
                            // the shapes of the model the maximum error against float64 is IDENTICAL with 3 and 4 terms
                            // (it comes from the tensor core's truncating fp32 accumulation, see acc2), so the default
                            // spends 25 % fewer MMAs; -DAX2D_TC_TERMS=4 restores the fourth term.
#endif

// ---------------------------------------------------------------------------------------------- PTX
// cycle counter of the development stamps (tools/tc_phases.py); compiled out unless built with -DAX2D_TC_PROFILE
__device__ __forceinline__ long long tc_clock() {
#ifdef AX2D_TC_PROFILE
  return clock64();
#else
  return 0;
#endif
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// Warp-collective issue: every function below with the suffix _w is called by ALL lanes of a converged warp and
// elects one lane inside the statement.  Inside an `if (lane == 0)` branch the compiler cannot use the uniform
// datapath that tcgen05.mma / TMA instructions need and wraps every one of them in an ELECT / branch loop with
// R2UR moves (measured: ~75 cycles per MMA issued -- as long as a 128 x 160 x 8 MMA occupies the tensor pipe).
// One k-block of the projection kernel in ONE statement: the (TC_BK / 8 = 2) k-steps x 4 split terms and the commit
// that frees the stage, under a single election.  Eight separately elected MMAs cost ~50 cycles of issue each
// (predicate + R2UR traffic per statement); for the small-M head products that was the whole main loop.
//   small accumulator: a_lo*b_lo (+)= , a_lo*b_hi, a_hi*b_lo      main accumulator: a_hi*b_hi
__device__ __forceinline__ void umma_kblock_ts_w(uint32_t t_main, uint32_t t_small, uint32_t a_hi, uint32_t a_lo, uint64_t db_hi,
                                                 uint64_t db_lo, uint32_t idesc, uint32_t acc_small_first, uint32_t acc_main_first,
                                                 uint64_t* free_bar) {
  static_assert(TC_BK == 16, "two k-steps of 8 per k-block");
  asm volatile(
      "{\n\t"
      ".reg .pred e, pf, pm, pt;\n\t"
      ".reg .b32 ah1, al1;\n\t"
      ".reg .b64 bh1, bl1;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %7, 0;\n\t"
      "setp.ne.b32 pm, %8, 0;\n\t"
      "setp.eq.b32 pt, %6, %6;\n\t"
      "add.u32 ah1, %2, 8;\n\t"
      "add.u32 al1, %3, 8;\n\t"
      "add.u64 bh1, %4, 2;\n\t"
      "add.u64 bl1, %5, 2;\n\t"
#if AX2D_TC_TERMS == 4
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [%3], %5, %6, pf;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [%3], %4, %6, pt;\n\t"
#else
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [%3], %4, %6, pf;\n\t"
#endif
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [%2], %5, %6, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], %4, %6, pm;\n\t"
#if AX2D_TC_TERMS == 4
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [al1], bl1, %6, pt;\n\t"
#endif
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [al1], bh1, %6, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [ah1], bl1, %6, pt;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [ah1], bh1, %6, pt;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
      "}\n" ::"r"(t_main),
      "r"(t_small), "r"(a_hi), "r"(a_lo), "l"(db_hi), "l"(db_lo), "r"(idesc), "r"(acc_small_first), "r"(acc_main_first),
      "r"(smem_u32(free_bar))
      : "memory");
}
// The same for 32-wide k-blocks: four k-steps (A advances 8 TMEM columns, B 32 bytes = 2 descriptor units per step).
__device__ __forceinline__ void umma_kblock_ts32_w(uint32_t t_main, uint32_t t_small, uint32_t a_hi, uint32_t a_lo, uint64_t db_hi,
                                                   uint64_t db_lo, uint32_t idesc, uint32_t acc_small_first, uint32_t acc_main_first,
                                                   uint64_t* free_bar) {
#if AX2D_TC_TERMS == 4
#define AX2D_TS_STEP(AH, AL, BH, BL, PF, PM)                                              \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [" AL "], " BL ", %6, " PF ";\n\t"        \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [" AL "], " BH ", %6, pt;\n\t"            \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [" AH "], " BL ", %6, pt;\n\t"            \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [" AH "], " BH ", %6, " PM ";\n\t"
#else
#define AX2D_TS_STEP(AH, AL, BH, BL, PF, PM)                                              \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [" AL "], " BH ", %6, " PF ";\n\t"        \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%1], [" AH "], " BL ", %6, pt;\n\t"            \
  "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [" AH "], " BH ", %6, " PM ";\n\t"
#endif
  asm volatile(
      "{\n\t"
      ".reg .pred e, pf, pm, pt;\n\t"
      ".reg .b32 ah1, al1, ah2, al2, ah3, al3;\n\t"
      ".reg .b64 bh1, bl1, bh2, bl2, bh3, bl3;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 pf, %7, 0;\n\t"
      "setp.ne.b32 pm, %8, 0;\n\t"
      "setp.eq.b32 pt, %6, %6;\n\t"
      "add.u32 ah1, %2, 8;\n\t add.u32 al1, %3, 8;\n\t add.u64 bh1, %4, 2;\n\t add.u64 bl1, %5, 2;\n\t"
      "add.u32 ah2, %2, 16;\n\t add.u32 al2, %3, 16;\n\t add.u64 bh2, %4, 4;\n\t add.u64 bl2, %5, 4;\n\t"
      "add.u32 ah3, %2, 24;\n\t add.u32 al3, %3, 24;\n\t add.u64 bh3, %4, 6;\n\t add.u64 bl3, %5, 6;\n\t"
      AX2D_TS_STEP("%2", "%3", "%4", "%5", "pf", "pm")
      AX2D_TS_STEP("ah1", "al1", "bh1", "bl1", "pt", "pt")
      AX2D_TS_STEP("ah2", "al2", "bh2", "bl2", "pt", "pt")
      AX2D_TS_STEP("ah3", "al3", "bh3", "bl3", "pt", "pt")
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
      "}\n" ::"r"(t_main),
      "r"(t_small), "r"(a_hi), "r"(a_lo), "l"(db_hi), "l"(db_lo), "r"(idesc), "r"(acc_small_first), "r"(acc_main_first),
      "r"(smem_u32(free_bar))
      : "memory");
#undef AX2D_TS_STEP
}
// expect_tx + the three tensor-map loads of one k-block (A raw, B hi, B lo), one election
__device__ __forceinline__ void tma_kblock_w(uint64_t* bar, uint32_t bytes, void* dst_a, const CUtensorMap* map_a, int ka, int m0,
                                             void* dst_bh, const CUtensorMap* map_bh, void* dst_bl, const CUtensorMap* map_bl,
                                             int kbcol, int n0) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%2], [%3, {%4, %5}], [%0];\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%6], [%7, {%10, %11}], [%0];\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%8], [%9, {%10, %11}], [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(bytes), "r"(smem_u32(dst_a)), "l"(reinterpret_cast<uint64_t>(map_a)), "r"(ka), "r"(m0), "r"(smem_u32(dst_bh)),
      "l"(reinterpret_cast<uint64_t>(map_bh)), "r"(smem_u32(dst_bl)), "l"(reinterpret_cast<uint64_t>(map_bl)), "r"(kbcol), "r"(n0)
      : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_w(void* dst, const CUtensorMap* map, uint64_t* bar, int c_inner, int c_outer) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_w(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n\t"
      "}\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_w(uint64_t* bar, uint32_t bytes) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
      "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
      "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major operand, 64-byte swizzle (cute::UMMA::SmemDescriptor; canonical layout
// Swizzle<2,4,3> o ((8,n),(4,2)):((16,SBO),(1,4)) in fp32 elements: rows of 64 bytes, 8-row groups of 512 bytes):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major) |
//   [32,46) stride byte offset >> 4 = 512 B between 8-row groups | [46,48) version = 1 (sm_100) | [61,64) layout = 4
__device__ __forceinline__ uint64_t smem_desc_k_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;
  return d;
}
// The same for the 128-byte swizzle (rows of 128 bytes, 8-row groups of 1024 bytes, layout type 2).
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128, N = bn.
__device__ __forceinline__ uint32_t idesc_tf32(int bn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) | (static_cast<uint32_t>(TC_BM >> 4) << 24);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 row-major matrix [outer rows, inner cols] with leading dimension ld; box = [box_outer x 16], 64 B swizzle,
// out-of-bounds elements read as zero.
static inline int make_map(CUtensorMap* map, const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_outer,
                           int box_inner = TC_BK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_64B,
                           CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32, int elem_bytes = 4) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("ax2d_gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
    return AX2D_ERR_UNSUPPORTED;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * static_cast<cuuint64_t>(elem_bytes)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("ax2d_gemm_tc: cuTensorMapEncodeTiled failed (%d) for a [%lld x %lld] matrix, ld %lld, box %d", (int)r,
              (long long)outer, (long long)inner, (long long)ld, box_outer);
    return AX2D_ERR_ARG;
  }
  return AX2D_OK;
}

// The same fp32 row-major matrix seen as [column chunks of 32][rows][32 columns]: ONE box {32, rows, chunks} lands `chunks`
// consecutive [rows x 32] blocks in shared memory, each in the canonical 128-byte-swizzled MN-major layout -- one TMA
// instruction per operand and k-block instead of one per 32-column chunk.
static inline int make_map_chunks(CUtensorMap* map, const void* ptr, int64_t cols, int64_t rows, int64_t ld, int box_rows,
                                  int box_chunks) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("ax2d_gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
    return AX2D_ERR_UNSUPPORTED;
  }
  cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(cols / 32)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 4u, 128u};
  cuuint32_t box[3] = {32, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(box_chunks)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("ax2d_gemm_tc: cuTensorMapEncodeTiled (3-D chunk view) failed (%d) for a [%lld x %lld] matrix, ld %lld", (int)r,
              (long long)rows, (long long)cols, (long long)ld);
    return AX2D_ERR_ARG;
  }
  return AX2D_OK;
}

}  // namespace ax2d
