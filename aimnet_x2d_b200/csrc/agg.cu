// a1 -- CSR gather-reduce for ShellConvolutionLayer.message_passing (models/layers.py:133-167).
//
// out[r,:] = addend[r,:] + sum_{k in row r} x[col[k],:]    (sequential CSR order, no atomics)
//
// Tiled kernel (shipped collation, no edge leaves its molecule): one CTA per tile of whole molecules.
//   * one elected thread issues a 1-D bulk async copy (TMA engine, cp.async.bulk) of the tile's rows
//     x[r0:r1,:] into shared memory; completion is signalled on an mbarrier;
//   * meanwhile every 8-lane group fetches rowptr and the first 8 column indices of its row;
//   * 8 lanes own one output row: lane j accumulates float4 columns j, j+8, ... in registers, reading
//     neighbour rows from shared memory (a quarter-warp reads 128 contiguous bytes -> conflict free);
//   * 128-bit coalesced streaming stores.
// HBM traffic = x once + out once + indices: the algorithmic minimum of SURVEY.md section 8d.
// The global-gather kernel is the general path (hop-offset targets, edges crossing tiles).
#include "common.cuh"

namespace ax2d {

template <int V, bool TILED>
__global__ void __launch_bounds__(256) agg_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ out,
                                                  int64_t ldo, const int32_t* __restrict__ rowptr,
                                                  const int32_t* __restrict__ col, const float* __restrict__ addend,
                                                  int64_t ld_addend, const int32_t* __restrict__ tile_ptr,
                                                  int64_t n_out_rows, int rows_per_cta) {
  pdl_enter();
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  constexpr int W4 = 8 * V;                 // float4 per row
  const int lane8 = threadIdx.x & 7;
  const int group = threadIdx.x >> 3;
  const int n_groups = blockDim.x >> 3;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);   // the 8 lanes of this group inside the warp

  int64_t r0, r1;
  const float4* src;                         // where neighbour rows are read from
  int64_t src_ld4;
  int64_t col_base;
  if (TILED) {
    r0 = tile_ptr[blockIdx.x];
    r1 = tile_ptr[blockIdx.x + 1];
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t bytes = static_cast<uint32_t>((r1 - r0) * W4 * 16);
      mbar_expect_tx(&bar, bytes);
      // split the copy so several TMA requests are in flight
      const uint32_t chunk = 16384;
      for (uint32_t off = 0; off < bytes; off += chunk) {
        const uint32_t n = bytes - off < chunk ? bytes - off : chunk;
        bulk_g2s(smem_raw + off, reinterpret_cast<const unsigned char*>(x + r0 * ldx) + off, n, &bar);
      }
    }
    src = reinterpret_cast<const float4*>(smem_raw);
    src_ld4 = W4;
    col_base = r0;
  } else {
    r0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
    r1 = r0 + rows_per_cta < n_out_rows ? r0 + rows_per_cta : n_out_rows;
    src = reinterpret_cast<const float4*>(x);
    src_ld4 = ldx >> 2;
    col_base = 0;
  }

  bool waited = !TILED;
  for (int64_t r = r0 + group; r < r1; r += n_groups) {
    const int beg = __ldg(rowptr + r);
    const int end = __ldg(rowptr + r + 1);
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = beg;
    int my_c = (k + lane8 < end) ? __ldg(col + k + lane8) : 0;   // prefetch before waiting for the tile
    if (!waited) {
      mbar_wait(&bar, 0);
      waited = true;
    }
    while (k < end) {
      const int cnt = end - k < 8 ? end - k : 8;
      const int nxt = (k + 8 + lane8 < end) ? __ldg(col + k + 8 + lane8) : 0;
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const int c = __shfl_sync(gmask, my_c, j, 8);
        const float4* row = src + (static_cast<int64_t>(c) - col_base) * src_ld4 + lane8;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float4 t = TILED ? row[8 * v] : __ldg(row + 8 * v);
          acc[v].x += t.x;
          acc[v].y += t.y;
          acc[v].z += t.z;
          acc[v].w += t.w;
        }
      }
      my_c = nxt;
      k += 8;
    }
    float4* o = reinterpret_cast<float4*>(out + r * ldo) + lane8;
    if (addend != nullptr) {
      const float4* a = reinterpret_cast<const float4*>(addend + r * ld_addend) + lane8;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 t = __ldg(a + 8 * v);
        acc[v].x += t.x;
        acc[v].y += t.y;
        acc[v].z += t.z;
        acc[v].w += t.w;
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) st_na_f4(o + 8 * v, acc[v]);
  }
  if (TILED && !waited) mbar_wait(&bar, 0);   // never exit with the bulk copy still in flight
}

// Persistent, pipelined tiled kernel (the shipped collation: no edge leaves its molecule, so a tile of whole
// molecules is self-contained).  The grid is sized to the machine (two CTAs per SM when the stage ring fits); every CTA
// walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... through a ring of AGG_STAGES shared-memory stages:
//   * a dedicated producer warp reads the per-tile descriptors {row0, row1, edge0, edge1} (32 tiles per coalesced
//     load) and issues, up to AGG_STAGES tiles ahead, three bulk async copies (TMA engine) per tile: the tile's rows of
//     x, its window of rowptr and its window of col (16-byte aligned super-sets) -- so neither feature nor index
//     latency is ever exposed to the consumers (three dependent DRAM round trips per tile otherwise);
//   * eight consumer warps: 8 lanes own one output row, each lane accumulates float4 columns lane, lane + 8, ... in
//     registers from shared memory (a quarter-warp reads 128 contiguous bytes: conflict-free), strictly in CSR order
//     (bit-exact against the reference CPU scatter_add), 128-bit streaming stores; no atomics;
//   * full / empty mbarriers per stage (complete_tx for the copies, one arrive per consumer warp to release).
constexpr int AGG_STAGES = 4;
constexpr int AGG_CONSUMERS = 256;
__device__ unsigned long long* g_agg_dbg = nullptr;     // development aid (ax2d_debug_agg_timing)
__device__ __forceinline__ unsigned long long agg_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// (the register tests come FIRST: with the pointer test first every thread of every CTA loaded the __device__ variable
// from L2 -- an exposed ~700-cycle round trip twice per tile on the critical path of every consumer warp, which is what
// the first version of these kernels spent a third of its time on)
#define AGG_STAMP(slot)                                                                  \
  do {                                                                                   \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                           \
      unsigned long long* dbg_ = g_agg_dbg;                                              \
      if (dbg_ != nullptr) dbg_[slot] = agg_gtime();                                     \
    }                                                                                    \
  } while (0)
// Tiles are dealt to the persistent CTAs in SERPENTINE order: round k goes to the CTAs in ascending order when k is even and
// in descending order when k is odd.  The tile plan lists the tiles by decreasing edge count (collate.refresh_tile_info), so
// every CTA gets one tile of each size class and the per-CTA (and per-SM) work is balanced: with plain round-robin over
// tiles in memory order the CTA lifetimes of one C2 launch ranged from 9.7 to 17.5 us (per-CTA %globaltimer stamps,
// tools/agg_gen3.py; the end time of an SM is 4.5 us + 4.4 ns per edge it was dealt).
__device__ __forceinline__ int agg_tile_of(int k, int first, int stride) {
  return k * stride + ((k & 1) ? stride - 1 - first : first);
}
__device__ __forceinline__ int agg_my_tiles(int n_tiles, int first, int stride) {
  const int full = n_tiles / stride, rem = n_tiles - full * stride;
  const int pos = (full & 1) ? stride - 1 - first : first;
  return full + (pos < rem ? 1 : 0);
}
template <int V>
__global__ void __launch_bounds__(AGG_CONSUMERS + 32, 2) agg_tiles_kernel(
    const float* __restrict__ x, float* __restrict__ out, int64_t ldo, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ col, const float* __restrict__ addend, int64_t ld_addend,
    const int4* __restrict__ tile_info, int n_tiles, int stages, uint32_t x_bytes, uint32_t rp_bytes, uint32_t col_bytes) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[AGG_STAGES], empty[AGG_STAGES];
  __shared__ int4 s_info[AGG_STAGES];
  constexpr int W4 = 8 * V;
  const uint32_t stage_bytes = x_bytes + rp_bytes + col_bytes;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_my = agg_my_tiles(n_tiles, first, stride);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  AGG_STAMP(0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], AGG_CONSUMERS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_enter();      // the barrier set-up above overlaps the previous kernel's tail; no global access before this
  AGG_STAMP(1);

  if (warp == AGG_CONSUMERS / 32) {
    // ================================================================= producer warp
    for (int base = 0; base < n_my; base += 32) {
      int4 mine = make_int4(0, 0, 0, 0);
      if (base + lane < n_my) mine = __ldg(tile_info + agg_tile_of(base + lane, first, stride));
      const int cnt = n_my - base < 32 ? n_my - base : 32;
      for (int j = 0; j < cnt; ++j) {
        const int it = base + j;
        int4 inf;
        inf.x = __shfl_sync(0xffffffffu, mine.x, j);
        inf.y = __shfl_sync(0xffffffffu, mine.y, j);
        inf.z = __shfl_sync(0xffffffffu, mine.z, j);
        inf.w = __shfl_sync(0xffffffffu, mine.w, j);
        if (lane == 0) {
          const int s = it % stages;
          mbar_wait(&empty[s], (static_cast<uint32_t>(it / stages) & 1u) ^ 1u);
          unsigned char* dst = smem_raw + static_cast<size_t>(s) * stage_bytes;
          const uint32_t xb = static_cast<uint32_t>(inf.y - inf.x) * W4 * 16;
          const int rs = inf.x & ~3, es = inf.z & ~3;
          const uint32_t rb = static_cast<uint32_t>((inf.y + 1 - rs + 3) & ~3) * 4;
          const uint32_t cb = static_cast<uint32_t>((inf.w - es + 3) & ~3) * 4;
          s_info[s] = inf;                                  // visible to the consumers through the mbarrier
          mbar_expect_tx(&full[s], xb + rb + cb);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(x + static_cast<int64_t>(inf.x) * (W4 * 4));
          for (uint32_t off = 0; off < xb; off += 16384) {
            const uint32_t n = xb - off < 16384u ? xb - off : 16384u;
            bulk_g2s(dst + off, src + off, n, &full[s]);
          }
          bulk_g2s(dst + x_bytes, rowptr + rs, rb, &full[s]);
          if (cb > 0) bulk_g2s(dst + x_bytes + rp_bytes, col + es, cb, &full[s]);
        }
      }
    }
    return;
  }

  // =================================================================== consumers
  // 8 lanes own one output row (a warp works on 4 rows, the CTA on 32 = a whole QM9-sized tile at once); lane j of a
  // group accumulates the float4 columns j, j + 8, ... so a 128-bit shared-memory load of a group is one full 128-byte
  // wavefront without bank conflicts.  Accumulation is strictly in CSR order per row (bit-exact), 128-bit streaming
  // stores.  (Measured alternatives -- one warp per row with 32- or 128-bit loads, warp-uniform predicated groups --
  // were 20-30 % slower: the kernel is bound by shared-memory wavefronts, E * width * 4 / 128 of them.)
  constexpr int n_groups = AGG_CONSUMERS / 8;
  const int lane8 = lane & 7;
  const int group = threadIdx.x >> 3;
  const unsigned gmask = 0xffu << (threadIdx.x & 24);
  for (int it = 0; it < n_my; ++it) {
    const int s = it % stages;
    mbar_wait(&full[s], static_cast<uint32_t>(it / stages) & 1u);
    if (it < 6) AGG_STAMP(2 + 2 * it);
    const int4 inf = s_info[s];
    const unsigned char* st = smem_raw + static_cast<size_t>(s) * stage_bytes;
    const float4* xs = reinterpret_cast<const float4*>(st) + lane8;
    const int32_t* rp = reinterpret_cast<const int32_t*>(st + x_bytes) - (inf.x & ~3);             // indexed by global row
    const int32_t* cs = reinterpret_cast<const int32_t*>(st + x_bytes + rp_bytes) - (inf.z & ~3);  // indexed by global edge
    for (int r = inf.x + group; r < inf.y; r += n_groups) {
      const int beg = rp[r], end = rp[r + 1];
      float4 add4[V];
      if (addend != nullptr) {
        const float4* a = reinterpret_cast<const float4*>(addend + static_cast<int64_t>(r) * ld_addend) + lane8;
#pragma unroll
        for (int v = 0; v < V; ++v) add4[v] = ld_nc_f4(a + 8 * v);
      }
      float4 acc[V];
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = beg; k < end; k += 8) {
        const int my_c = (k + lane8 < end) ? (cs[k + lane8] - inf.x) * W4 : 0;
        const int cnt = end - k < 8 ? end - k : 8;
        int j = 0;
        // Two neighbour rows are loaded into distinct registers before either is added: the compiler otherwise
        // recycles one float4 per load and serialises every 128-bit load behind the previous four adds (measured:
        // ~1 us per row).  The additions stay in CSR order.
        for (; j + 2 <= cnt; j += 2) {
          const float4* ra = xs + __shfl_sync(gmask, my_c, j, 8);
          const float4* rb = xs + __shfl_sync(gmask, my_c, j + 1, 8);
          float4 ta[V], tb[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            ta[v] = ra[8 * v];
            tb[v] = rb[8 * v];
          }
#pragma unroll
          for (int v = 0; v < V; ++v) {
            acc[v].x = (acc[v].x + ta[v].x) + tb[v].x;
            acc[v].y = (acc[v].y + ta[v].y) + tb[v].y;
            acc[v].z = (acc[v].z + ta[v].z) + tb[v].z;
            acc[v].w = (acc[v].w + ta[v].w) + tb[v].w;
          }
        }
        if (j < cnt) {
          const float4* ra = xs + __shfl_sync(gmask, my_c, j, 8);
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const float4 tv = ra[8 * v];
            acc[v].x += tv.x; acc[v].y += tv.y; acc[v].z += tv.z; acc[v].w += tv.w;
          }
        }
      }
      if (addend != nullptr) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          acc[v].x += add4[v].x; acc[v].y += add4[v].y; acc[v].z += add4[v].z; acc[v].w += add4[v].w;
        }
      }
      float4* o = reinterpret_cast<float4*>(out + static_cast<int64_t>(r) * ldo) + lane8;
#pragma unroll
      for (int v = 0; v < V; ++v) st_na_f4(o + 8 * v, acc[v]);
    }
    __syncwarp();
    if (it < 6) AGG_STAMP(3 + 2 * it);
    if (lane == 0) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
  }
  AGG_STAMP(15);
}

__device__ __forceinline__ uint32_t bf2_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// ------------------------------------------------------------------------------------------------ warp-per-row consumers
// Second-generation persistent tile kernel (fp32 and bf16 rows).  Same producer / stage ring as agg_tiles_kernel; the
// consumers differ: ONE WARP owns an output row.  Lane l accumulates the 32-bit words l, l + 32, ... of the row (fp32: one
// column per word, bf16: two), so every shared-memory load of a neighbour row is a conflict-free 128-byte wavefront and --
// the point -- every loop bound (the row's degree) is WARP-UNIFORM: the 8-lanes-per-row mapping above lets the four rows of
// a warp diverge on their degrees, and the hardware then issues each row's instructions separately with 8 active lanes
// (measured: ~33 issue slots per (row, edge); per-CTA %globaltimer stamps showed 2.3 us per 21-row tile, all of it
// instruction issue).  Here one instruction serves a whole row: ~13 issue slots per (row, edge).  EDGES neighbour rows are
// loaded before the first of them is added, the additions stay strictly in CSR order (bit-exact).
constexpr int AGG2_WARPS = 8;             // consumer warps per CTA (+ 1 producer warp) when three CTAs share an SM
constexpr int AGG2_MAX_WARPS = 24;        // ... and when large tiles (drug-like molecules) leave room for one CTA only
constexpr int AGG2_EDGES = 4;             // neighbour rows in flight per warp
template <int V, bool BF16, int NW>
__global__ void __launch_bounds__(32 * (NW + 1), NW == AGG2_WARPS ? 3 : 1) agg_rows_kernel(
    const uint32_t* __restrict__ x, uint32_t* __restrict__ out, int64_t ldo_w, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ col, const uint32_t* __restrict__ addend, int64_t lda_w, const int4* __restrict__ tile_info,
    int n_tiles, int stages, int words, uint32_t x_bytes, uint32_t rp_bytes, uint32_t col_bytes, int dbg_flags) {
  // words: 32-bit words per row (fp32: width, bf16: width / 2); V = ceil(words / 32) words per lane
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[AGG_STAGES], empty[AGG_STAGES];
  __shared__ int4 s_info[AGG_STAGES];
  const uint32_t stage_bytes = x_bytes + rp_bytes + col_bytes;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_my = agg_my_tiles(n_tiles, first, stride);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int n_cw = NW;                                       // consumer warps; the last warp is the producer

  AGG_STAMP(0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], n_cw);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_enter();      // the barrier set-up above overlaps the previous kernel's tail; no global access before this

  if (warp == n_cw) {
    // ================================================================= producer warp
    const uint32_t row_bytes = static_cast<uint32_t>(words) * 4u;
    for (int base = 0; base < n_my; base += 32) {
      int4 mine = make_int4(0, 0, 0, 0);
      if (base + lane < n_my) mine = __ldg(tile_info + agg_tile_of(base + lane, first, stride));
      const int cnt = n_my - base < 32 ? n_my - base : 32;
      for (int j = 0; j < cnt; ++j) {
        const int it = base + j;
        int4 inf;
        inf.x = __shfl_sync(0xffffffffu, mine.x, j);
        inf.y = __shfl_sync(0xffffffffu, mine.y, j);
        inf.z = __shfl_sync(0xffffffffu, mine.z, j);
        inf.w = __shfl_sync(0xffffffffu, mine.w, j);
        if (lane == 0) {
          const int s = it % stages;
          mbar_wait(&empty[s], (static_cast<uint32_t>(it / stages) & 1u) ^ 1u);
          unsigned char* dst = smem_raw + static_cast<size_t>(s) * stage_bytes;
          const uint32_t xb = static_cast<uint32_t>(inf.y - inf.x) * row_bytes;
          const int rs = inf.x & ~3, es = inf.z & ~3;
          const uint32_t rb = static_cast<uint32_t>((inf.y + 1 - rs + 3) & ~3) * 4;
          const uint32_t cb = static_cast<uint32_t>((inf.w - es + 3) & ~3) * 4;
          s_info[s] = inf;
          mbar_expect_tx(&full[s], xb + rb + cb);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(x) + static_cast<int64_t>(inf.x) * row_bytes;
          for (uint32_t off = 0; off < xb; off += 16384) {
            const uint32_t n = xb - off < 16384u ? xb - off : 16384u;
            bulk_g2s(dst + off, src + off, n, &full[s]);
          }
          bulk_g2s(dst + x_bytes, rowptr + rs, rb, &full[s]);
          if (cb > 0) bulk_g2s(dst + x_bytes + rp_bytes, col + es, cb, &full[s]);
        }
      }
    }
    return;
  }

  // =================================================================== consumers: one warp per row
  bool on[V];                                   // word lane + 32 v exists (only the last slot can be partial)
#pragma unroll
  for (int v = 0; v < V; ++v) on[v] = lane + 32 * v < words;
  constexpr int A = BF16 ? 2 : 1;               // fp32 accumulators per word
  AGG_STAMP(1);
  for (int it = 0; it < n_my; ++it) {
    const int s = it % stages;
    mbar_wait(&full[s], static_cast<uint32_t>(it / stages) & 1u);
    if (it < 6) AGG_STAMP(2 + 2 * it);
    const int4 inf = s_info[s];
    const unsigned char* st = smem_raw + static_cast<size_t>(s) * stage_bytes;
    const uint32_t* xs = reinterpret_cast<const uint32_t*>(st) + lane;
    const int32_t* rp = reinterpret_cast<const int32_t*>(st + x_bytes) - (inf.x & ~3);             // indexed by global row
    const int32_t* cs = reinterpret_cast<const int32_t*>(st + x_bytes + rp_bytes) - (inf.z & ~3);  // indexed by global edge
    for (int r = inf.x + warp; r < inf.y; r += n_cw) {
      const int beg = rp[r], end = rp[r + 1];
      uint32_t addw[V];
      if (addend != nullptr) {
        const uint32_t* a = addend + static_cast<int64_t>(r) * lda_w + lane;
#pragma unroll
        for (int v = 0; v < V; ++v) addw[v] = on[v] ? __ldg(a + 32 * v) : 0u;
      }
      float acc[V][A];
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int e = 0; e < A; ++e) acc[v][e] = 0.f;
      auto add_word = [&](int v, uint32_t w) {
        if constexpr (BF16) {
          acc[v][0] += __uint_as_float(w << 16);
          acc[v][1] += __uint_as_float(w & 0xFFFF0000u);
        } else {
          acc[v][0] += __uint_as_float(w);
        }
      };
      for (int k = beg; k < ((dbg_flags & 1) ? beg : end); k += 32) {
        const int cnt = end - k < 32 ? end - k : 32;
        const int my_off = lane < cnt ? (cs[k + lane] - inf.x) * words : 0;       // one coalesced index load per 32 edges
        int j = 0;
        for (; j + AGG2_EDGES <= cnt; j += AGG2_EDGES) {
          uint32_t t[AGG2_EDGES][V];
#pragma unroll
          for (int e = 0; e < AGG2_EDGES; ++e) {
            const uint32_t* row = xs + __shfl_sync(0xffffffffu, my_off, j + e);
#pragma unroll
            for (int v = 0; v < V; ++v) t[e][v] = (v + 1 < V || on[v]) ? row[32 * v] : 0u;
          }
#pragma unroll
          for (int e = 0; e < AGG2_EDGES; ++e)
#pragma unroll
            for (int v = 0; v < V; ++v) add_word(v, t[e][v]);
        }
        for (; j < cnt; ++j) {
          const uint32_t* row = xs + __shfl_sync(0xffffffffu, my_off, j);
          uint32_t t[V];
#pragma unroll
          for (int v = 0; v < V; ++v) t[v] = (v + 1 < V || on[v]) ? row[32 * v] : 0u;
#pragma unroll
          for (int v = 0; v < V; ++v) add_word(v, t[v]);
        }
      }
      if (addend != nullptr) {
#pragma unroll
        for (int v = 0; v < V; ++v) add_word(v, addw[v]);
      }
      uint32_t* o = out + static_cast<int64_t>(r) * ldo_w + lane;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        uint32_t w;
        if constexpr (BF16) w = bf2_pack(acc[v][0], acc[v][1]);
        else w = __float_as_uint(acc[v][0]);
        if ((v + 1 < V || on[v]) && !(dbg_flags & 2)) o[32 * v] = w;
      }
    }
    __syncwarp();
    if (it < 6) AGG_STAMP(3 + 2 * it);
    if (lane == 0) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
  }
  AGG_STAMP(15);
}

static int g_agg_debug_ctas = 0, g_agg_debug_stages = 0, g_agg_use_rows = 1, g_agg_debug_flags = 0;
template <int V, bool BF16>
static int launch_agg_rows(const void* x, void* out, int64_t ldo, const int32_t* rowptr, const int32_t* col, const void* addend,
                           int64_t ld_addend, const int32_t* tile_info, int64_t n_tiles, int max_tile_rows, int max_tile_edges,
                           int width, cudaStream_t st) {
  const int words = BF16 ? width / 2 : width;
  const size_t xb = (static_cast<size_t>(max_tile_rows) * words * 4 + 15) / 16 * 16;
  const size_t rb = (static_cast<size_t>(max_tile_rows) + 8) * 4 / 16 * 16 + 16;
  const size_t cb = (static_cast<size_t>(max_tile_edges) + 8) * 4 / 16 * 16 + 16;
  const size_t stage = xb + rb + cb;
  if (stage > 220 * 1024) {
    set_error("ax2d_agg: tile of %d rows / %d edges needs %zu bytes of shared memory", max_tile_rows, max_tile_edges, stage);
    return AX2D_ERR_UNSUPPORTED;
  }
  // as many CTAs per SM as give each a ring of >= 3 stages (at most 3 CTAs: 27 warps), else fewer CTAs with what fits
  int ctas_per_sm = 3;
  while (ctas_per_sm > 1 && (220 * 1024 / ctas_per_sm) / stage < 3) --ctas_per_sm;
  if (g_agg_debug_ctas > 0) ctas_per_sm = g_agg_debug_ctas;
  int stages = static_cast<int>((220 * 1024 / ctas_per_sm) / stage);
  stages = stages > AGG_STAGES ? AGG_STAGES : (stages < 1 ? 1 : stages);
  if (g_agg_debug_stages > 0 && g_agg_debug_stages <= stages) stages = g_agg_debug_stages;
  const size_t smem = stage * stages;
  // 8 consumer warps per CTA when two or three CTAs share an SM; large tiles (drug-like molecules) that leave room for one
  // CTA only get 24, so that an SM always has 16-24 rows in flight
  const bool wide = ctas_per_sm == 1;
  static size_t configured[2] = {0, 0};
  if (smem > 48 * 1024 && smem > configured[wide]) {
    if (wide) cudaFuncSetAttribute(agg_rows_kernel<V, BF16, AGG2_MAX_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    else cudaFuncSetAttribute(agg_rows_kernel<V, BF16, AGG2_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured[wide] = smem;
  }
  int64_t grid = static_cast<int64_t>(kNumSMs) * ctas_per_sm;
  grid = grid > n_tiles ? n_tiles : grid;
  const int esz = BF16 ? 2 : 4;
  if (wide)
    launch_k(agg_rows_kernel<V, BF16, AGG2_MAX_WARPS>, dim3(static_cast<unsigned>(grid)), dim3(32 * (AGG2_MAX_WARPS + 1)), smem, st, 
        static_cast<const uint32_t*>(x), static_cast<uint32_t*>(out), ldo * esz / 4, rowptr, col, static_cast<const uint32_t*>(addend),
        ld_addend * esz / 4, reinterpret_cast<const int4*>(tile_info), static_cast<int>(n_tiles), stages, words,
        static_cast<uint32_t>(xb), static_cast<uint32_t>(rb), static_cast<uint32_t>(cb), g_agg_debug_flags);
  else
    launch_k(agg_rows_kernel<V, BF16, AGG2_WARPS>, dim3(static_cast<unsigned>(grid)), dim3(32 * (AGG2_WARPS + 1)), smem, st, 
        static_cast<const uint32_t*>(x), static_cast<uint32_t*>(out), ldo * esz / 4, rowptr, col, static_cast<const uint32_t*>(addend),
        ld_addend * esz / 4, reinterpret_cast<const int4*>(tile_info), static_cast<int>(n_tiles), stages, words,
        static_cast<uint32_t>(xb), static_cast<uint32_t>(rb), static_cast<uint32_t>(cb), g_agg_debug_flags);
  return launch_status("ax2d_agg");
}

// ------------------------------------------------------------------------------------------------ packed-add consumers
// Third-generation per-edge kernel (bit-exact, CSR order).  The warp-per-row kernel above is bound by instruction issue
// (SASS: ~14 issue slots per (row, edge) + ~65 per row, 210 per QM9 row; fp32 and bf16 run at the same speed although bf16
// moves half the bytes).  Blackwell halves the add count:
//   * fp32: a lane owns 64-bit units (two adjacent columns) lane, lane + 32, ...: one LDS.64 + ONE packed add.rn.f32x2
//     (SASS FADD2) per unit -- 3 + 3 instead of 5 + 5 instructions per neighbour row of 160 columns, same IEEE sums;
//   * bf16: a lane owns 32-bit words (two columns): the mixed-precision add.rn.f32.bf16 (SASS FHADD.BF16 with a free
//     .H0 / .H1 operand selector) adds a bf16 half straight into an fp32 accumulator -- no unpack shifts / masks;
//   * the tail of a row's edge list runs as ONE predicated batch (no per-edge serial remainder loop);
//   * rows go round-robin over the consumer warps ACROSS tiles (a 21-row QM9 tile on 8 warps otherwise leaves three warps
//     idle for a third of every tile).
template <bool BF16> struct AggAcc;
template <> struct AggAcc<false> {
  using unit = unsigned long long;
  unsigned long long v;
  __device__ __forceinline__ void zero() { v = 0ull; }
  __device__ __forceinline__ void add(unit t) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(t)); }
  __device__ __forceinline__ unit result() const { return v; }
  static __device__ __forceinline__ unit ldg(const unit* p) {
    unit r;
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ void stg(unit* p, unit w) {
    asm volatile("st.global.L1::no_allocate.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
  }
};
template <> struct AggAcc<true> {
  using unit = uint32_t;
  // (measured against widening with shift / mask and one packed FADD2 per word: 15.3 / 16.6 us here, 18.1 / 20.2 us there
  // on C2 shapes, 83 / 86 against 103 / 110 us on drug-like tiles -- the integer instructions cost more than they save)
  float lo, hi;
  __device__ __forceinline__ void zero() { lo = 0.f; hi = 0.f; }
  __device__ __forceinline__ void add(unit w) {
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tadd.rn.f32.bf16 %0, l, %0;\n\tadd.rn.f32.bf16 %1, h, %1;\n\t}"
        : "+f"(lo), "+f"(hi)
        : "r"(w));
  }
  __device__ __forceinline__ unit result() const { return bf2_pack(lo, hi); }
  static __device__ __forceinline__ unit ldg(const unit* p) {
    unit r;
    asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
  }
  static __device__ __forceinline__ void stg(unit* p, unit w) {
    asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" ::"l"(p), "r"(w) : "memory");
  }
};

template <int U, bool BF16, int NW>
__global__ void __launch_bounds__(32 * (NW + 1), NW == AGG2_WARPS ? 3 : 1) agg_rows3_kernel(
    const unsigned char* __restrict__ x, typename AggAcc<BF16>::unit* __restrict__ out, int64_t ldo_u,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const typename AggAcc<BF16>::unit* __restrict__ addend,
    int64_t lda_u, const int4* __restrict__ tile_info, int n_tiles, int stages, int units, uint32_t x_bytes, uint32_t rp_bytes,
    uint32_t col_bytes, int dbg_flags) {
  // units: units per row (fp32: width / 2 64-bit units, bf16: width / 2 32-bit words); U = ceil(units / 32) units per lane
  using Acc = AggAcc<BF16>;
  using unit = typename Acc::unit;
  constexpr int EB = U <= 3 ? 4 : (U <= 6 ? 2 : 1);      // neighbour rows in flight per warp
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[AGG_STAGES], empty[AGG_STAGES];
  __shared__ int4 s_info[AGG_STAGES];
  const uint32_t stage_bytes = x_bytes + rp_bytes + col_bytes;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_my = agg_my_tiles(n_tiles, first, stride);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  AGG_STAMP(0);
  if ((dbg_flags & 4) && threadIdx.x == 0) g_agg_dbg[16 + 2 * blockIdx.x] = agg_gtime();      // development: per-CTA start
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NW);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_enter();      // the barrier set-up above overlaps the previous kernel's tail; no global access before this

  if (warp == NW) {
    // ================================================================= producer warp (as in agg_rows_kernel)
    const uint32_t row_bytes = static_cast<uint32_t>(units) * sizeof(unit);
    for (int base = 0; base < n_my; base += 32) {
      int4 mine = make_int4(0, 0, 0, 0);
      if (base + lane < n_my) mine = __ldg(tile_info + agg_tile_of(base + lane, first, stride));
      const int cnt = n_my - base < 32 ? n_my - base : 32;
      for (int j = 0; j < cnt; ++j) {
        const int it = base + j;
        int4 inf;
        inf.x = __shfl_sync(0xffffffffu, mine.x, j);
        inf.y = __shfl_sync(0xffffffffu, mine.y, j);
        inf.z = __shfl_sync(0xffffffffu, mine.z, j);
        inf.w = __shfl_sync(0xffffffffu, mine.w, j);
        if (lane == 0) {
          const int s = it % stages;
          mbar_wait(&empty[s], (static_cast<uint32_t>(it / stages) & 1u) ^ 1u);
          unsigned char* dst = smem_raw + static_cast<size_t>(s) * stage_bytes;
          const uint32_t xb = static_cast<uint32_t>(inf.y - inf.x) * row_bytes;
          const int rs = inf.x & ~3, es = inf.z & ~3;
          const uint32_t rb = static_cast<uint32_t>((inf.y + 1 - rs + 3) & ~3) * 4;
          const uint32_t cb = static_cast<uint32_t>((inf.w - es + 3) & ~3) * 4;
          s_info[s] = inf;
          mbar_expect_tx(&full[s], xb + rb + cb);
          const unsigned char* src = x + static_cast<int64_t>(inf.x) * row_bytes;
          for (uint32_t off = 0; off < xb; off += 16384) {
            const uint32_t n = xb - off < 16384u ? xb - off : 16384u;
            bulk_g2s(dst + off, src + off, n, &full[s]);
          }
          bulk_g2s(dst + x_bytes, rowptr + rs, rb, &full[s]);
          if (cb > 0) bulk_g2s(dst + x_bytes + rp_bytes, col + es, cb, &full[s]);
        }
      }
    }
    return;
  }

  // =================================================================== consumers: one warp per row
  const bool on_last = lane + 32 * (U - 1) < units;       // only the last unit slot of a lane can fall off the row
  AGG_STAMP(1);
  int rot = 0;                                            // rows handed out so far, modulo NW (identical in every warp)
  for (int it = 0; it < n_my; ++it) {
    const int s = it % stages;
    mbar_wait(&full[s], static_cast<uint32_t>(it / stages) & 1u);
    if (it < 6) AGG_STAMP(2 + 2 * it);
    const int4 inf = s_info[s];
    const unsigned char* st = smem_raw + static_cast<size_t>(s) * stage_bytes;
    const unit* xs = reinterpret_cast<const unit*>(st) + lane;
    const int32_t* rp = reinterpret_cast<const int32_t*>(st + x_bytes) - (inf.x & ~3);             // indexed by global row
    const int32_t* cs = reinterpret_cast<const int32_t*>(st + x_bytes + rp_bytes) - (inf.z & ~3);  // indexed by global edge
    int mine = warp - rot;
    if (mine < 0) mine += NW;
    rot = (rot + (inf.y - inf.x)) % NW;
    for (int r = inf.x + mine; r < inf.y; r += NW) {
      const int beg = rp[r], end = rp[r + 1];
      unit addw[U];
      if (addend != nullptr) {
        const unit* a = addend + static_cast<int64_t>(r) * lda_u + lane;
#pragma unroll
        for (int u = 0; u < U; ++u) addw[u] = (u + 1 < U || on_last) ? Acc::ldg(a + 32 * u) : unit(0);
      }
      Acc acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u].zero();
      for (int k = beg; k < ((dbg_flags & 1) ? beg : end); k += 32) {
        const int cnt = end - k < 32 ? end - k : 32;
        const int my_off = lane < cnt ? (cs[k + lane] - inf.x) * units : 0;       // one coalesced index load per 32 edges
        int j = 0;
        for (; j + EB <= cnt; j += EB) {
          unit t[EB][U];
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            const unit* row = xs + __shfl_sync(0xffffffffu, my_off, j + e);
#pragma unroll
            for (int u = 0; u < U; ++u) t[e][u] = (u + 1 < U || on_last) ? row[32 * u] : unit(0);
          }
#pragma unroll
          for (int e = 0; e < EB; ++e)
#pragma unroll
            for (int u = 0; u < U; ++u) acc[u].add(t[e][u]);
        }
        if (EB > 1 && j < cnt) {
          // the remainder as one batch: warp-uniform predicates, the loads of all remaining rows before the first add
          unit t[EB][U];
#pragma unroll
          for (int e = 0; e + 1 < EB; ++e) {
            const unit* row = xs + __shfl_sync(0xffffffffu, my_off, (j + e) & 31);
            if (j + e < cnt) {
#pragma unroll
              for (int u = 0; u < U; ++u) t[e][u] = (u + 1 < U || on_last) ? row[32 * u] : unit(0);
            }
          }
#pragma unroll
          for (int e = 0; e + 1 < EB; ++e)
            if (j + e < cnt) {
#pragma unroll
              for (int u = 0; u < U; ++u) acc[u].add(t[e][u]);
            }
        }
      }
      if (addend != nullptr) {
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u].add(addw[u]);
      }
      unit* o = out + static_cast<int64_t>(r) * ldo_u + lane;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if ((u + 1 < U || on_last) && !(dbg_flags & 2)) Acc::stg(o + 32 * u, acc[u].result());
    }
    __syncwarp();
    if (it < 6) AGG_STAMP(3 + 2 * it);
    if (lane == 0) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
  }
  AGG_STAMP(15);
  if ((dbg_flags & 4) && threadIdx.x == 0) g_agg_dbg[17 + 2 * blockIdx.x] = agg_gtime();      // ... and end (consumer warp 0)
}

template <int U, bool BF16>
static int launch_agg_rows3(const void* x, void* out, int64_t ldo, const int32_t* rowptr, const int32_t* col, const void* addend,
                            int64_t ld_addend, const int32_t* tile_info, int64_t n_tiles, int max_tile_rows, int max_tile_edges,
                            int width, cudaStream_t st) {
  using unit = typename AggAcc<BF16>::unit;
  const int esz = BF16 ? 2 : 4;
  const int units = width / 2;                                   // two columns per unit in both types
  const size_t xb = (static_cast<size_t>(max_tile_rows) * width * esz + 15) / 16 * 16;
  const size_t rb = (static_cast<size_t>(max_tile_rows) + 8) * 4 / 16 * 16 + 16;
  const size_t cb = (static_cast<size_t>(max_tile_edges) + 8) * 4 / 16 * 16 + 16;
  const size_t stage = xb + rb + cb;
  if (stage > 220 * 1024) {
    set_error("ax2d_agg: tile of %d rows / %d edges needs %zu bytes of shared memory", max_tile_rows, max_tile_edges, stage);
    return AX2D_ERR_UNSUPPORTED;
  }
  int ctas_per_sm = 3;
  while (ctas_per_sm > 1 && (220 * 1024 / ctas_per_sm) / stage < 3) --ctas_per_sm;
  if (g_agg_debug_ctas > 0 && g_agg_debug_ctas <= 3) ctas_per_sm = g_agg_debug_ctas;
  int stages = static_cast<int>((220 * 1024 / ctas_per_sm) / stage);
  stages = stages > AGG_STAGES ? AGG_STAGES : (stages < 1 ? 1 : stages);
  if (g_agg_debug_stages > 0 && g_agg_debug_stages <= stages) stages = g_agg_debug_stages;
  const size_t smem = stage * stages;
  const bool wide = ctas_per_sm == 1;
  static size_t configured[2] = {0, 0};
  if (smem > 48 * 1024 && smem > configured[wide]) {
    if (wide) cudaFuncSetAttribute(agg_rows3_kernel<U, BF16, AGG2_MAX_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    else cudaFuncSetAttribute(agg_rows3_kernel<U, BF16, AGG2_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured[wide] = smem;
  }
  int64_t grid = static_cast<int64_t>(kNumSMs) * ctas_per_sm;
  grid = grid > n_tiles ? n_tiles : grid;
  if (wide)
    launch_k(agg_rows3_kernel<U, BF16, AGG2_MAX_WARPS>, dim3(static_cast<unsigned>(grid)), dim3(32 * (AGG2_MAX_WARPS + 1)), smem, st, 
        static_cast<const unsigned char*>(x), static_cast<unit*>(out), ldo / 2, rowptr, col, static_cast<const unit*>(addend),
        ld_addend / 2, reinterpret_cast<const int4*>(tile_info), static_cast<int>(n_tiles), stages, units,
        static_cast<uint32_t>(xb), static_cast<uint32_t>(rb), static_cast<uint32_t>(cb), g_agg_debug_flags);
  else
    launch_k(agg_rows3_kernel<U, BF16, AGG2_WARPS>, dim3(static_cast<unsigned>(grid)), dim3(32 * (AGG2_WARPS + 1)), smem, st, 
        static_cast<const unsigned char*>(x), static_cast<unit*>(out), ldo / 2, rowptr, col, static_cast<const unit*>(addend),
        ld_addend / 2, reinterpret_cast<const int4*>(tile_info), static_cast<int>(n_tiles), stages, units,
        static_cast<uint32_t>(xb), static_cast<uint32_t>(rb), static_cast<uint32_t>(cb), g_agg_debug_flags);
  return launch_status("ax2d_agg");
}

// ------------------------------------------------------------------------------------------------ tensor-core tiles
// Third formulation of the tile kernel: the gather-reduce of a whole-molecule tile IS a small dense product
//     out_tile[R, W] = A_tile[R, R] * x_tile[R, W],        A[r, c] = 1 iff the tile holds the edge (target r, source c),
// because no edge leaves its tile.  The per-edge kernels above read one 640-byte row of x from shared memory and issue
// ~13-33 instructions PER EDGE (E = 10 N): shared-memory wavefronts and instruction issue bound them at ~19 us per launch
// whatever the data type (tools/agg_dbg.py: 8.6 us without the edge loop).  Here every staged row is read once per 16-row
// output block by ldmatrix, and the sums are done by the tensor cores (mma.sync m16n8k16, bf16 operands, fp32 accumulation):
//   * the consumers write the tile's 0/1 adjacency as a dense bf16 matrix (one store per edge) ...
//   * ... and re-pack the staged rows into padded (ldmatrix conflict-free) bf16 planes: ONE plane for bf16 features, THREE
//     for fp32 features -- x = b0 + b1 + b2 with bf16 terms is EXACT (8 + 8 + 8 significand bits), and 0/1 times a bf16
//     term is exact, so the only difference to the sequential sum is the ORDER of the fp32 additions (~1e-7 relative; the
//     per-edge kernels stay available for bit-exact results: ax2d_agg_config);
//   * warp w owns (16-row block, 16-column pair) items; per item and 16-neighbour step: one ldmatrix.x4 of A, one
//     ldmatrix.x4.trans of x per plane, two MMAs per plane.
// Requires a duplicate-free edge list (a dense 0/1 matrix cannot count) -- checked at collation (GraphIndex.unique_edges).
constexpr int AGG3_THREADS = 256;              // 8 consumer warps (+ 1 producer warp)
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void agg3_sync() { asm volatile("bar.sync 1, %0;" ::"n"(AGG3_THREADS) : "memory"); }

template <bool BF16>
__global__ void __launch_bounds__(AGG3_THREADS + 32) agg_mma_kernel(
    const uint32_t* __restrict__ x, void* __restrict__ out, int64_t ldo, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ col, const void* __restrict__ addend, int64_t lda, const int4* __restrict__ tile_info,
    int n_tiles, int stages, int width, int rpad, uint32_t x_bytes, uint32_t rp_bytes, uint32_t col_bytes) {
  // width: columns (multiple of 16); rpad: tile-row capacity rounded up to 16 (rows and neighbours of the dense product)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[AGG_STAGES], empty[AGG_STAGES];
  __shared__ int4 s_info[AGG_STAGES];
  constexpr int P = BF16 ? 1 : 3;                         // bf16 planes of x
  const uint32_t stage_bytes = x_bytes + rp_bytes + col_bytes;
  const int SA = rpad + 8;                                // row pitch of the adjacency, bf16 elements (16-byte skew per row)
  const int SX = width + 8;                               // row pitch of the x planes
  unsigned char* work = smem_raw + static_cast<size_t>(stages) * stage_bytes;
  uint16_t* As = reinterpret_cast<uint16_t*>(work);                                      // [rpad][SA]
  uint16_t* Xp = As + static_cast<size_t>(rpad) * SA;                                     // [P][rpad][SX]
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_my = agg_my_tiles(n_tiles, first, stride);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  AGG_STAMP(0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], AGG3_THREADS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_enter();      // the barrier set-up above overlaps the previous kernel's tail; no global access before this

  if (warp == AGG3_THREADS / 32) {
    // ================================================================= producer warp (as in agg_rows_kernel)
    const uint32_t row_bytes = static_cast<uint32_t>(width) * (BF16 ? 2u : 4u);
    for (int base = 0; base < n_my; base += 32) {
      int4 mine = make_int4(0, 0, 0, 0);
      if (base + lane < n_my) mine = __ldg(tile_info + agg_tile_of(base + lane, first, stride));
      const int cnt = n_my - base < 32 ? n_my - base : 32;
      for (int j = 0; j < cnt; ++j) {
        const int it = base + j;
        int4 inf;
        inf.x = __shfl_sync(0xffffffffu, mine.x, j);
        inf.y = __shfl_sync(0xffffffffu, mine.y, j);
        inf.z = __shfl_sync(0xffffffffu, mine.z, j);
        inf.w = __shfl_sync(0xffffffffu, mine.w, j);
        if (lane == 0) {
          const int s = it % stages;
          mbar_wait(&empty[s], (static_cast<uint32_t>(it / stages) & 1u) ^ 1u);
          unsigned char* dst = smem_raw + static_cast<size_t>(s) * stage_bytes;
          const uint32_t xb = static_cast<uint32_t>(inf.y - inf.x) * row_bytes;
          const int rs = inf.x & ~3, es = inf.z & ~3;
          const uint32_t rb = static_cast<uint32_t>((inf.y + 1 - rs + 3) & ~3) * 4;
          const uint32_t cb = static_cast<uint32_t>((inf.w - es + 3) & ~3) * 4;
          s_info[s] = inf;
          mbar_expect_tx(&full[s], xb + rb + cb);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(x) + static_cast<int64_t>(inf.x) * row_bytes;
          for (uint32_t off = 0; off < xb; off += 16384) {
            const uint32_t n = xb - off < 16384u ? xb - off : 16384u;
            bulk_g2s(dst + off, src + off, n, &full[s]);
          }
          bulk_g2s(dst + x_bytes, rowptr + rs, rb, &full[s]);
          if (cb > 0) bulk_g2s(dst + x_bytes + rp_bytes, col + es, cb, &full[s]);
        }
      }
    }
    return;
  }

  // =================================================================== consumers
  const int tid = threadIdx.x;
  const int lane8 = lane & 7, group = tid >> 3;          // 8 lanes per row when the adjacency is written
  const int n_pairs = width >> 4;                        // 16-column pairs of n-blocks
  const int m_blocks = rpad >> 4, k_steps = rpad >> 4;
  const int n_items = m_blocks * n_pairs;
  const int units = width >> 3;                          // 8-element units per row
  for (int it = 0; it < n_my; ++it) {
    const int s = it % stages;
    mbar_wait(&full[s], static_cast<uint32_t>(it / stages) & 1u);
    if (it < 3) AGG_STAMP(1 + 4 * it);
    const int4 inf = s_info[s];
    const int rows = inf.y - inf.x;
    const unsigned char* st = smem_raw + static_cast<size_t>(s) * stage_bytes;
    const int32_t* rp = reinterpret_cast<const int32_t*>(st + x_bytes) - (inf.x & ~3);
    const int32_t* cs = reinterpret_cast<const int32_t*>(st + x_bytes + rp_bytes) - (inf.z & ~3);
    // ---- (1) zero the adjacency, re-pack x into padded bf16 planes (rows beyond the tile: zeros -- 0 * NaN would poison)
    for (int i = tid; i < rpad * (SA >> 3); i += AGG3_THREADS) reinterpret_cast<uint4*>(As)[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < rpad * units; i += AGG3_THREADS) {
      const int r = i / units, u = i - r * units;
      if constexpr (BF16) {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (r < rows) v = reinterpret_cast<const uint4*>(st)[r * units + u];
        *reinterpret_cast<uint4*>(Xp + r * SX + u * 8) = v;
      } else {
        float f[8];
        if (r < rows) {
          const float4 a = reinterpret_cast<const float4*>(st)[(r * units + u) * 2], b = reinterpret_cast<const float4*>(st)[(r * units + u) * 2 + 1];
          f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
        uint32_t w[3][4];
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          // x = b0 + b1 + b2 exactly: every remainder is computed exactly in fp32 and has 8 significand bits fewer
          float r0[2], r1[2];
          uint32_t p0 = bf2_pack(f[e], f[e + 1]);
          r0[0] = f[e] - __uint_as_float(p0 << 16);
          r0[1] = f[e + 1] - __uint_as_float(p0 & 0xFFFF0000u);
          uint32_t p1 = bf2_pack(r0[0], r0[1]);
          r1[0] = r0[0] - __uint_as_float(p1 << 16);
          r1[1] = r0[1] - __uint_as_float(p1 & 0xFFFF0000u);
          w[0][e >> 1] = p0;
          w[1][e >> 1] = p1;
          w[2][e >> 1] = bf2_pack(r1[0], r1[1]);
        }
#pragma unroll
        for (int p = 0; p < 3; ++p)
          *reinterpret_cast<uint4*>(Xp + (p * rpad + r) * SX + u * 8) = make_uint4(w[p][0], w[p][1], w[p][2], w[p][3]);
      }
    }
    agg3_sync();
    if (it < 3) AGG_STAMP(2 + 4 * it);
    // ---- (2) the tile's edges -> ones (8 lanes per row walk the row's edge list)
    for (int r = group; r < rows; r += AGG3_THREADS / 8) {
      const int beg = rp[inf.x + r], end = rp[inf.x + r + 1];
      for (int k = beg + lane8; k < end; k += 8) As[r * SA + (cs[k] - inf.x)] = 0x3F80;
    }
    agg3_sync();
    if (it < 3) AGG_STAMP(3 + 4 * it);
    // the staged tile (raw rows, rowptr and col windows) is consumed: hand the stage back while the products run
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    // ---- (3) products: item = (16-row block mb, 16-column pair np)
    const uint32_t as_base = smem_u32(As), xp_base = smem_u32(Xp);
    const int g = lane >> 2, t = lane & 3;
    for (int item = warp; item < n_items; item += AGG3_THREADS / 32) {
      const int mb = item % m_blocks, np = item / m_blocks;
      if (mb * 16 >= rows) continue;
      float acc_hi[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float acc_lo[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      for (int ks = 0; ks < k_steps; ++ks) {
        if (ks * 16 >= rows) break;                              // neighbours live in [0, rows)
        uint32_t a0, a1, a2, a3;
        // ldmatrix.x4 row addresses: matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15) of the 16 x 16 block
        ldsm_x4(as_base + 2u * static_cast<uint32_t>((mb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * SA + ks * 16 + (lane >> 4) * 8),
                a0, a1, a2, a3);
#pragma unroll
        for (int p = 0; p < P; ++p) {
          uint32_t b0, b1, b2, b3;
          // .trans: stored rows = neighbours k, 8 contiguous columns each; matrices (k 0-7 | 8-15) x (n-block 0 | 1)
          ldsm_x4_t(xp_base + 2u * static_cast<uint32_t>((p * rpad + ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * SX + np * 16 + (lane >> 4) * 8),
                    b0, b1, b2, b3);
          float (*acc)[4] = (BF16 || p == 0) ? acc_hi : acc_lo;
          mma_bf16(acc[0], a0, a1, a2, a3, b0, b1);
          mma_bf16(acc[1], a0, a1, a2, a3, b2, b3);
        }
      }
      // ---- (4) out[row, col .. col + 1] (+ addend): rows g and g + 8 of the block, columns np * 16 + nb * 8 + 2 t
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = mb * 16 + g + 8 * h;
          if (r >= rows) continue;
          const int64_t grow = static_cast<int64_t>(inf.x) + r;
          const int c = np * 16 + nb * 8 + 2 * t;
          float v0 = acc_hi[nb][2 * h] + acc_lo[nb][2 * h], v1 = acc_hi[nb][2 * h + 1] + acc_lo[nb][2 * h + 1];
          if constexpr (BF16) {
            if (addend != nullptr) {
              const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint16_t*>(addend) + grow * lda + c));
              v0 += __uint_as_float(a << 16);
              v1 += __uint_as_float(a & 0xFFFF0000u);
            }
            *reinterpret_cast<uint32_t*>(static_cast<uint16_t*>(out) + grow * ldo + c) = bf2_pack(v0, v1);
          } else {
            if (addend != nullptr) {
              const float2 a = __ldg(reinterpret_cast<const float2*>(static_cast<const float*>(addend) + grow * lda + c));
              v0 += a.x;
              v1 += a.y;
            }
            *reinterpret_cast<float2*>(static_cast<float*>(out) + grow * ldo + c) = make_float2(v0, v1);
          }
        }
    }
    if (it < 3) AGG_STAMP(4 + 4 * it);
    agg3_sync();              // As / Xp are rewritten by the next tile
  }
  AGG_STAMP(15);
}

template <bool BF16>
static int launch_agg_mma(const void* x, void* out, int64_t ldo, const int32_t* rowptr, const int32_t* col, const void* addend,
                          int64_t ld_addend, const int32_t* tile_info, int64_t n_tiles, int max_tile_rows, int max_tile_edges,
                          int width, cudaStream_t st, bool* handled) {
  *handled = false;
  const int rpad = (max_tile_rows + 15) / 16 * 16;
  const size_t esz = BF16 ? 2 : 4;
  const size_t xb = (static_cast<size_t>(max_tile_rows) * width * esz + 15) / 16 * 16;
  const size_t rb = (static_cast<size_t>(max_tile_rows) + 8) * 4 / 16 * 16 + 16;
  const size_t cb = (static_cast<size_t>(max_tile_edges) + 8) * 4 / 16 * 16 + 16;
  const size_t stage = xb + rb + cb;
  const size_t work = (static_cast<size_t>(rpad) * (rpad + 8) + static_cast<size_t>(BF16 ? 1 : 3) * rpad * (width + 8)) * 2 + 16;
  if (width % 16 != 0 || rpad > 128) return AX2D_OK;               // not applicable: the caller takes the per-edge kernel
  int ctas_per_sm = 4;
  while (ctas_per_sm > 1 && (220 * 1024 / ctas_per_sm) < static_cast<int64_t>(work + 2 * stage)) --ctas_per_sm;
  if (g_agg_debug_ctas > 0) ctas_per_sm = g_agg_debug_ctas;
  const int64_t room = 220 * 1024 / ctas_per_sm - static_cast<int64_t>(work);
  if (room < static_cast<int64_t>(stage)) return AX2D_OK;
  int stages = static_cast<int>(room / stage);
  stages = stages > AGG_STAGES ? AGG_STAGES : stages;
  if (g_agg_debug_stages > 0 && g_agg_debug_stages <= stages) stages = g_agg_debug_stages;
  const size_t smem = stage * stages + work;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(agg_mma_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  int64_t grid = static_cast<int64_t>(kNumSMs) * ctas_per_sm;
  grid = grid > n_tiles ? n_tiles : grid;
  launch_k(agg_mma_kernel<BF16>, dim3(static_cast<unsigned>(grid)), dim3(AGG3_THREADS + 32), smem, st, 
      static_cast<const uint32_t*>(x), out, ldo, rowptr, col, addend, ld_addend, reinterpret_cast<const int4*>(tile_info),
      static_cast<int>(n_tiles), stages, width, rpad, static_cast<uint32_t>(xb), static_cast<uint32_t>(rb), static_cast<uint32_t>(cb));
  *handled = true;
  return launch_status("ax2d_agg");
}

// ------------------------------------------------------------------------------------------------ bf16 features
// The same persistent tile pipeline for bf16 rows (BASELINE configs[3]): the staged tile is half the bytes, FOUR lanes own
// one output row (lane j accumulates the 16-byte units j, j + 4, ... = 8 bf16 each, in fp32 registers, CSR order), the
// result is rounded to bf16 once per row.  V = width / 32 units per lane.
__device__ __forceinline__ void bf8_add(float* acc, const uint4& u) {
  acc[0] += __uint_as_float(u.x << 16); acc[1] += __uint_as_float(u.x & 0xFFFF0000u);
  acc[2] += __uint_as_float(u.y << 16); acc[3] += __uint_as_float(u.y & 0xFFFF0000u);
  acc[4] += __uint_as_float(u.z << 16); acc[5] += __uint_as_float(u.z & 0xFFFF0000u);
  acc[6] += __uint_as_float(u.w << 16); acc[7] += __uint_as_float(u.w & 0xFFFF0000u);
}
__device__ __forceinline__ uint4 ld_nc_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_na_u4(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
template <int V>
__global__ void __launch_bounds__(AGG_CONSUMERS + 32, 2) agg_tiles_bf16_kernel(
    const uint16_t* __restrict__ x, uint16_t* __restrict__ out, int64_t ldo, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ col, const uint16_t* __restrict__ addend, int64_t ld_addend,
    const int4* __restrict__ tile_info, int n_tiles, int stages, uint32_t x_bytes, uint32_t rp_bytes, uint32_t col_bytes) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full[AGG_STAGES], empty[AGG_STAGES];
  __shared__ int4 s_info[AGG_STAGES];
  constexpr int U = 4 * V;                      // 16-byte units per row
  const uint32_t stage_bytes = x_bytes + rp_bytes + col_bytes;
  const int first = blockIdx.x, stride = gridDim.x;
  const int n_my = agg_my_tiles(n_tiles, first, stride);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], AGG_CONSUMERS / 32);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_enter();      // the barrier set-up above overlaps the previous kernel's tail; no global access before this

  if (warp == AGG_CONSUMERS / 32) {
    // ================================================================= producer warp (as in agg_tiles_kernel)
    for (int base = 0; base < n_my; base += 32) {
      int4 mine = make_int4(0, 0, 0, 0);
      if (base + lane < n_my) mine = __ldg(tile_info + agg_tile_of(base + lane, first, stride));
      const int cnt = n_my - base < 32 ? n_my - base : 32;
      for (int j = 0; j < cnt; ++j) {
        const int it = base + j;
        int4 inf;
        inf.x = __shfl_sync(0xffffffffu, mine.x, j);
        inf.y = __shfl_sync(0xffffffffu, mine.y, j);
        inf.z = __shfl_sync(0xffffffffu, mine.z, j);
        inf.w = __shfl_sync(0xffffffffu, mine.w, j);
        if (lane == 0) {
          const int s = it % stages;
          mbar_wait(&empty[s], (static_cast<uint32_t>(it / stages) & 1u) ^ 1u);
          unsigned char* dst = smem_raw + static_cast<size_t>(s) * stage_bytes;
          const uint32_t xb = static_cast<uint32_t>(inf.y - inf.x) * U * 16;
          const int rs = inf.x & ~3, es = inf.z & ~3;
          const uint32_t rb = static_cast<uint32_t>((inf.y + 1 - rs + 3) & ~3) * 4;
          const uint32_t cb = static_cast<uint32_t>((inf.w - es + 3) & ~3) * 4;
          s_info[s] = inf;
          mbar_expect_tx(&full[s], xb + rb + cb);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(x) + static_cast<int64_t>(inf.x) * (U * 16);
          for (uint32_t off = 0; off < xb; off += 16384) {
            const uint32_t n = xb - off < 16384u ? xb - off : 16384u;
            bulk_g2s(dst + off, src + off, n, &full[s]);
          }
          bulk_g2s(dst + x_bytes, rowptr + rs, rb, &full[s]);
          if (cb > 0) bulk_g2s(dst + x_bytes + rp_bytes, col + es, cb, &full[s]);
        }
      }
    }
    return;
  }

  // =================================================================== consumers: 4 lanes per row, 64 rows per CTA pass
  constexpr int n_groups = AGG_CONSUMERS / 4;
  const int lane4 = lane & 3;
  const int group = threadIdx.x >> 2;
  const unsigned gmask = 0xfu << (threadIdx.x & 28);
  for (int it = 0; it < n_my; ++it) {
    const int s = it % stages;
    mbar_wait(&full[s], static_cast<uint32_t>(it / stages) & 1u);
    const int4 inf = s_info[s];
    const unsigned char* st = smem_raw + static_cast<size_t>(s) * stage_bytes;
    const uint4* xs = reinterpret_cast<const uint4*>(st) + lane4;
    const int32_t* rp = reinterpret_cast<const int32_t*>(st + x_bytes) - (inf.x & ~3);
    const int32_t* cs = reinterpret_cast<const int32_t*>(st + x_bytes + rp_bytes) - (inf.z & ~3);
    for (int r = inf.x + group; r < inf.y; r += n_groups) {
      const int beg = rp[r], end = rp[r + 1];
      uint4 add4[V];
      if (addend != nullptr) {
        const uint4* a = reinterpret_cast<const uint4*>(addend + static_cast<int64_t>(r) * ld_addend) + lane4;
#pragma unroll
        for (int v = 0; v < V; ++v) add4[v] = ld_nc_u4(a + 4 * v);
      }
      float acc[V][8];
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[v][e] = 0.f;
      for (int k = beg; k < end; k += 4) {
        const int my_c = (k + lane4 < end) ? (cs[k + lane4] - inf.x) * U : 0;
        const int cnt = end - k < 4 ? end - k : 4;
        int j = 0;
        for (; j + 2 <= cnt; j += 2) {            // two neighbour rows in flight; additions stay in CSR order
          const uint4* ra = xs + __shfl_sync(gmask, my_c, j, 4);
          const uint4* rb = xs + __shfl_sync(gmask, my_c, j + 1, 4);
          uint4 ta[V], tb[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            ta[v] = ra[4 * v];
            tb[v] = rb[4 * v];
          }
#pragma unroll
          for (int v = 0; v < V; ++v) {
            bf8_add(acc[v], ta[v]);
            bf8_add(acc[v], tb[v]);
          }
        }
        if (j < cnt) {
          const uint4* ra = xs + __shfl_sync(gmask, my_c, j, 4);
#pragma unroll
          for (int v = 0; v < V; ++v) bf8_add(acc[v], ra[4 * v]);
        }
      }
      if (addend != nullptr) {
#pragma unroll
        for (int v = 0; v < V; ++v) bf8_add(acc[v], add4[v]);
      }
      uint4* o = reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * ldo) + lane4;
#pragma unroll
      for (int v = 0; v < V; ++v)
        st_na_u4(o + 4 * v, make_uint4(bf2_pack(acc[v][0], acc[v][1]), bf2_pack(acc[v][2], acc[v][3]),
                                        bf2_pack(acc[v][4], acc[v][5]), bf2_pack(acc[v][6], acc[v][7])));
    }
    __syncwarp();
    if (lane == 0) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
    }
  }
}

// general bf16 path (any CSR, width % 8 == 0): one warp per output row, lanes over the 16-byte units
__global__ void __launch_bounds__(256) agg_generic_bf16_kernel(const uint16_t* __restrict__ x, int64_t ldx, uint16_t* __restrict__ out,
                                                               int64_t ldo, const int32_t* __restrict__ rowptr,
                                                               const int32_t* __restrict__ col, const uint16_t* __restrict__ addend,
                                                               int64_t ld_addend, int64_t n_out_rows, int w8) {
  pdl_enter();
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * warps + (threadIdx.x >> 5);
  if (r >= n_out_rows) return;
  const int beg = __ldg(rowptr + r), end = __ldg(rowptr + r + 1);
  for (int c = lane; c < w8; c += 32) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = beg; k < end; ++k)
      bf8_add(acc, __ldg(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(__ldg(col + k)) * ldx) + c));
    if (addend != nullptr) bf8_add(acc, __ldg(reinterpret_cast<const uint4*>(addend + r * ld_addend) + c));
    reinterpret_cast<uint4*>(out + r * ldo)[c] = make_uint4(bf2_pack(acc[0], acc[1]), bf2_pack(acc[2], acc[3]),
                                                            bf2_pack(acc[4], acc[5]), bf2_pack(acc[6], acc[7]));
  }
}

template <int V>
static int launch_agg_bf16(const uint16_t* x, uint16_t* out, int64_t ldo, const int32_t* rowptr, const int32_t* col,
                           const uint16_t* addend, int64_t ld_addend, const int32_t* tile_info, int64_t n_tiles, int max_tile_rows,
                           int max_tile_edges, cudaStream_t st) {
  const size_t xb = static_cast<size_t>(max_tile_rows) * V * 4 * 16;
  const size_t rb = (static_cast<size_t>(max_tile_rows) + 8) * 4 / 16 * 16 + 16;
  const size_t cb = (static_cast<size_t>(max_tile_edges) + 8) * 4 / 16 * 16 + 16;
  const size_t stage = xb + rb + cb;
  if (stage > 220 * 1024) {
    set_error("ax2d_agg: tile of %d rows / %d edges needs %zu bytes of shared memory", max_tile_rows, max_tile_edges, stage);
    return AX2D_ERR_UNSUPPORTED;
  }
  int ctas_per_sm = 2;
  int stages = static_cast<int>((110 * 1024) / stage);
  if (stages < 2) {
    stages = static_cast<int>((220 * 1024) / stage);
    ctas_per_sm = 1;
  }
  stages = stages > AGG_STAGES ? AGG_STAGES : stages;
  const size_t smem = stage * stages;
  auto kern = agg_tiles_bf16_kernel<V>;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    configured = smem;
  }
  int64_t grid = static_cast<int64_t>(kNumSMs) * ctas_per_sm;
  grid = grid > n_tiles ? n_tiles : grid;
  launch_k(kern, dim3(static_cast<unsigned>(grid)), dim3(AGG_CONSUMERS + 32), smem, st, 
      x, out, ldo, rowptr, col, addend, ld_addend, reinterpret_cast<const int4*>(tile_info), static_cast<int>(n_tiles), stages,
      static_cast<uint32_t>(xb), static_cast<uint32_t>(rb), static_cast<uint32_t>(cb));
  return launch_status("ax2d_agg");
}

// generic width (multiple of 4, not of 32): one lane per float4 column, strided over the row
__global__ void __launch_bounds__(256) agg_generic_kernel(const float* __restrict__ x, int64_t ldx,
                                                          float* __restrict__ out, int64_t ldo,
                                                          const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ col,
                                                          const float* __restrict__ addend, int64_t ld_addend,
                                                          int64_t n_out_rows, int w4) {
  pdl_enter();
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blockIdx.x) * warps + (threadIdx.x >> 5);
  if (r >= n_out_rows) return;
  const int beg = __ldg(rowptr + r), end = __ldg(rowptr + r + 1);
  for (int c4 = lane; c4 < w4; c4 += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = beg; k < end; ++k) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(x + static_cast<int64_t>(__ldg(col + k)) * ldx) + c4);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if (addend != nullptr) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(addend + r * ld_addend) + c4);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    reinterpret_cast<float4*>(out + r * ldo)[c4] = acc;
  }
}

template <int V>
static int launch_agg(const float* x, int64_t ldx, float* out, int64_t ldo, int64_t n_out_rows,
                      const int32_t* rowptr, const int32_t* col, const float* addend, int64_t ld_addend,
                      const int32_t* tile_info, int64_t n_tiles, int max_tile_rows, int max_tile_edges, cudaStream_t st) {
  if (tile_info != nullptr) {
    const size_t xb = static_cast<size_t>(max_tile_rows) * V * 8 * 16;
    const size_t rb = (static_cast<size_t>(max_tile_rows) + 8) * 4 / 16 * 16 + 16;
    const size_t cb = (static_cast<size_t>(max_tile_edges) + 8) * 4 / 16 * 16 + 16;
    const size_t stage = xb + rb + cb;
    if (stage > 220 * 1024) {
      set_error("ax2d_agg: tile of %d rows / %d edges needs %zu bytes of shared memory", max_tile_rows, max_tile_edges, stage);
      return AX2D_ERR_UNSUPPORTED;
    }
    // ring of up to AGG_STAGES tiles per CTA; two CTAs per SM when the ring fits half of the shared memory
    int ctas_per_sm = 2;
    int stages = static_cast<int>((110 * 1024) / stage);
    if (stages < 2) {
      stages = static_cast<int>((220 * 1024) / stage);
      ctas_per_sm = 1;
    }
    stages = stages > AGG_STAGES ? AGG_STAGES : stages;
    const size_t smem = stage * stages;
    auto kern = agg_tiles_kernel<V>;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      configured = smem;
    }
    int64_t grid = static_cast<int64_t>(kNumSMs) * ctas_per_sm;
    grid = grid > n_tiles ? n_tiles : grid;
    launch_k(kern, dim3(static_cast<unsigned>(grid)), dim3(AGG_CONSUMERS + 32), smem, st, 
        x, out, ldo, rowptr, col, addend, ld_addend, reinterpret_cast<const int4*>(tile_info), static_cast<int>(n_tiles),
        stages, static_cast<uint32_t>(xb), static_cast<uint32_t>(rb), static_cast<uint32_t>(cb));
  } else {
    const int rows_per_cta = 32;
    const int64_t grid = (n_out_rows + rows_per_cta - 1) / rows_per_cta;
    launch_k(agg_kernel<V, false>, dim3(static_cast<unsigned>(grid)), dim3(256), 0, st, x, ldx, out, ldo, rowptr, col, addend,
                                                                       ld_addend, nullptr, n_out_rows, rows_per_cta);
  }
  return launch_status("ax2d_agg");
}

}  // namespace ax2d

// development aid (not part of include/ax2d.h): CTAs per SM / stages of the warp-per-row kernels (0 = automatic) and which
// per-edge kernel runs (all of them bit-identical): 0 = first generation (8 / 4 lanes per row) for both types, 1 (default) =
// first generation for fp32 (the fp32 kernels are all bound by the shared-memory pipe and run at the same speed on B200,
// measured) and the packed-add warp-per-row kernel for bf16 (1.15-1.25x faster than the second generation, 1.7x than the
// first), 2 = second generation (warp per row) for both, 3 = packed-add kernel for both
extern "C" void ax2d_debug_agg_config(int ctas_per_sm, int stages, int use_rows_kernel) {
  ax2d::g_agg_debug_ctas = ctas_per_sm;
  ax2d::g_agg_debug_stages = stages & 0xff;
  ax2d::g_agg_debug_flags = stages >> 8;          // development: bit 0 skip the edge loop, bit 1 skip the stores
  ax2d::g_agg_use_rows = use_rows_kernel;
}
// development aid (not part of include/ax2d.h): 16 x u64 device buffer receiving %globaltimer stamps of CTA 0
extern "C" void ax2d_debug_agg_timing(unsigned long long* buf) {
  cudaMemcpyToSymbol(ax2d::g_agg_dbg, &buf, sizeof(buf));
}

// 1 iff the tensor-core tile kernel handles tiles of this shape (width % 16 == 0, tile capacity <= 128 rows, the dense
// adjacency + the bf16 planes + two stages fit the shared memory)
extern "C" int ax2d_agg_tiles_mma_supported(int width, int max_tile_rows, int max_tile_edges, int dtype) {
  using namespace ax2d;
  if (width <= 0 || width % 16 != 0 || max_tile_rows <= 0 || (dtype != AX2D_F32 && dtype != AX2D_BF16)) return 0;
  const int rpad = (max_tile_rows + 15) / 16 * 16;
  if (rpad > 128) return 0;
  const size_t esz = dtype == AX2D_BF16 ? 2 : 4;
  const size_t stage = (static_cast<size_t>(max_tile_rows) * width * esz + 15) / 16 * 16 + (static_cast<size_t>(max_tile_rows) + 8) * 4 / 16 * 16 +
                       16 + (static_cast<size_t>(max_tile_edges) + 8) * 4 / 16 * 16 + 16;
  const size_t work = (static_cast<size_t>(rpad) * (rpad + 8) + static_cast<size_t>(dtype == AX2D_BF16 ? 1 : 3) * rpad * (width + 8)) * 2 + 16;
  return work + 2 * stage <= 220 * 1024 ? 1 : 0;
}

// a1 on the tensor cores: the same contract as the tiled mode of ax2d_agg (whole-molecule tiles, tile-local columns) PLUS a
// duplicate-free edge list; fp32 results equal the sequential sums up to the order of the fp32 additions (see agg_mma_kernel).
extern "C" int ax2d_agg_tiles_mma(const void* x, int64_t ldx, int64_t n_rows, void* out, int64_t ldo, const int32_t* rowptr,
                                  const int32_t* col, const void* addend, int64_t ld_addend, int width, const int32_t* tile_info,
                                  int64_t n_tiles, int max_tile_rows, int max_tile_edges, int dtype, ax2d_stream_t stream) {
  using namespace ax2d;
  AX2D_CHECK_ARG(ax2d_agg_tiles_mma_supported(width, max_tile_rows, max_tile_edges, dtype),
                 "ax2d_agg_tiles_mma: unsupported tile shape (width %d, %d rows, %d edges, dtype %d)", width, max_tile_rows,
                 max_tile_edges, dtype);
  const int al = dtype == AX2D_BF16 ? 8 : 4;
  AX2D_CHECK_ARG(ldx == width && ldo % al == 0 && ldo >= width && (addend == nullptr || (ld_addend % al == 0 && ld_addend >= width)),
                 "ax2d_agg_tiles_mma: x must be contiguous (ldx == width); bad leading dimensions");
  AX2D_CHECK_ARG(tile_info != nullptr && n_tiles > 0 && max_tile_edges >= 0, "ax2d_agg_tiles_mma: tile plan required");
  AX2D_CHECK_ALIGN(x);
  AX2D_CHECK_ALIGN(out);
  AX2D_CHECK_ALIGN(addend);
  AX2D_CHECK_ALIGN(tile_info);
  AX2D_CHECK_ALIGN(rowptr);
  AX2D_CHECK_ALIGN(col);
  if (n_rows <= 0) return AX2D_OK;
  bool handled = false;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = dtype == AX2D_BF16
               ? launch_agg_mma<true>(x, out, ldo, rowptr, col, addend, ld_addend, tile_info, n_tiles, max_tile_rows, max_tile_edges, width, st, &handled)
               : launch_agg_mma<false>(x, out, ldo, rowptr, col, addend, ld_addend, tile_info, n_tiles, max_tile_rows, max_tile_edges, width, st, &handled);
  if (rc == AX2D_OK && !handled) {
    set_error("ax2d_agg_tiles_mma: tile shape does not fit the shared memory");
    return AX2D_ERR_UNSUPPORTED;
  }
  return rc;
}

extern "C" int ax2d_agg(const void* x, int64_t ldx, int64_t n_src_rows, void* out, int64_t ldo, int64_t n_out_rows,
                        const int32_t* rowptr, const int32_t* col, const void* addend, int64_t ld_addend, int width,
                        const int32_t* tile_info, int64_t n_tiles, int max_tile_rows, int max_tile_edges, int dtype,
                        ax2d_stream_t stream) {
  using namespace ax2d;
  if (dtype == AX2D_BF16) {
    // bf16 rows (leading dimensions in bf16 elements): fp32 accumulation in CSR order, one rounding per output row
    AX2D_CHECK_ARG(width > 0 && width % 8 == 0, "ax2d_agg: bf16 width %d must be a positive multiple of 8", width);
    AX2D_CHECK_ARG(ldx % 8 == 0 && ldo % 8 == 0 && ldx >= width && ldo >= width, "ax2d_agg: bad leading dimensions");
    AX2D_CHECK_ARG(addend == nullptr || ld_addend % 8 == 0, "ax2d_agg: bad addend leading dimension");
    AX2D_CHECK_ALIGN(x);
    AX2D_CHECK_ALIGN(out);
    AX2D_CHECK_ALIGN(addend);
    if (n_out_rows <= 0) return AX2D_OK;
    cudaStream_t st16 = reinterpret_cast<cudaStream_t>(stream);
    const uint16_t* xh = static_cast<const uint16_t*>(x);
    uint16_t* oh = static_cast<uint16_t*>(out);
    const uint16_t* ah = static_cast<const uint16_t*>(addend);
    if (tile_info != nullptr && width % 32 == 0 && width <= 32 * 16) {
      AX2D_CHECK_ARG(n_out_rows == n_src_rows && ldx == width && n_tiles > 0 && max_tile_rows > 0 && max_tile_edges >= 0,
                     "ax2d_agg: tiled mode needs n_out_rows == n_src_rows, ldx == width, width %% 32 == 0");
      AX2D_CHECK_ALIGN(tile_info);
      AX2D_CHECK_ALIGN(rowptr);
      AX2D_CHECK_ALIGN(col);
#define AX2D_AGG16_CASE(V) \
  case V: return (g_agg_use_rows == 3 || g_agg_use_rows == 1) \
      ? launch_agg_rows3<(V + 1) / 2, true>(xh, oh, ldo, rowptr, col, ah, ld_addend, tile_info, n_tiles, max_tile_rows, max_tile_edges, width, st16) \
      : g_agg_use_rows \
      ? launch_agg_rows<(V + 1) / 2, true>(xh, oh, ldo, rowptr, col, ah, ld_addend, tile_info, n_tiles, max_tile_rows, max_tile_edges, width, st16) \
      : launch_agg_bf16<V>(xh, oh, ldo, rowptr, col, ah, ld_addend, tile_info, n_tiles, max_tile_rows, max_tile_edges, st16);
      switch (width / 32) {
        AX2D_AGG16_CASE(1) AX2D_AGG16_CASE(2) AX2D_AGG16_CASE(3) AX2D_AGG16_CASE(4) AX2D_AGG16_CASE(5) AX2D_AGG16_CASE(6)
        AX2D_AGG16_CASE(7) AX2D_AGG16_CASE(8) AX2D_AGG16_CASE(9) AX2D_AGG16_CASE(10) AX2D_AGG16_CASE(11) AX2D_AGG16_CASE(12)
        AX2D_AGG16_CASE(13) AX2D_AGG16_CASE(14) AX2D_AGG16_CASE(15) AX2D_AGG16_CASE(16)
      }
#undef AX2D_AGG16_CASE
    }
    const int warps = 8;
    launch_k(agg_generic_bf16_kernel, dim3(static_cast<unsigned>((n_out_rows + warps - 1) / warps)), dim3(warps * 32), 0, st16, 
        xh, ldx, oh, ldo, rowptr, col, ah, ld_addend, n_out_rows, width / 8);
    return launch_status("ax2d_agg");
  }
  if (dtype != AX2D_F32) {
    set_error("ax2d_agg: dtype %d not supported", dtype);
    return AX2D_ERR_DTYPE;
  }
  AX2D_CHECK_ARG(width > 0 && width % 4 == 0, "ax2d_agg: width %d must be a positive multiple of 4", width);
  AX2D_CHECK_ARG(ldx % 4 == 0 && ldo % 4 == 0 && ldx >= width && ldo >= width, "ax2d_agg: bad leading dimensions");
  AX2D_CHECK_ARG(addend == nullptr || ld_addend % 4 == 0, "ax2d_agg: bad addend leading dimension");
  AX2D_CHECK_ALIGN(x);
  AX2D_CHECK_ALIGN(out);
  AX2D_CHECK_ALIGN(addend);
  if (n_out_rows <= 0) return AX2D_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float* xf = static_cast<const float*>(x);
  float* of = static_cast<float*>(out);
  const float* af = static_cast<const float*>(addend);
  if (tile_info != nullptr) {
    AX2D_CHECK_ARG(n_out_rows == n_src_rows && ldx == width && width % 32 == 0 && n_tiles > 0 && max_tile_rows > 0 &&
                       max_tile_edges >= 0,
                   "ax2d_agg: tiled mode needs n_out_rows == n_src_rows, ldx == width, width %% 32 == 0");
    AX2D_CHECK_ALIGN(tile_info);
    AX2D_CHECK_ALIGN(rowptr);
    AX2D_CHECK_ALIGN(col);
  }
  if (width % 32 != 0 || width > 32 * 16) {
    const int warps = 8;
    launch_k(agg_generic_kernel, dim3(static_cast<unsigned>((n_out_rows + warps - 1) / warps)), dim3(warps * 32), 0, st, 
        xf, ldx, of, ldo, rowptr, col, af, ld_addend, n_out_rows, width / 4);
    return launch_status("ax2d_agg");
  }
  // fp32, default configuration: small tiles (QM9-sized molecules, three CTAs per SM) run at the same speed in every
  // generation (shared-memory pipe); LARGE tiles (drug-like molecules: one CTA with 24 consumer warps per SM) are 1.3x
  // faster in the packed-add kernel than in the first generation (100 vs 132 us on 2048 drug-like molecules x 4 hops)
  const bool big_tiles = tile_info != nullptr &&
      (220 * 1024 / 3) / (static_cast<size_t>(max_tile_rows) * width * 4 + static_cast<size_t>(max_tile_rows + max_tile_edges) * 4 + 64) < 3;
#define AX2D_AGG_CASE(V) \
  case V: return (tile_info != nullptr && (g_agg_use_rows == 3 || (g_agg_use_rows == 1 && big_tiles))) \
      ? launch_agg_rows3<(V + 1) / 2, false>(xf, of, ldo, rowptr, col, af, ld_addend, tile_info, n_tiles, max_tile_rows, max_tile_edges, width, st) \
      : (tile_info != nullptr && g_agg_use_rows == 2) \
      ? launch_agg_rows<V, false>(xf, of, ldo, rowptr, col, af, ld_addend, tile_info, n_tiles, max_tile_rows, max_tile_edges, width, st) \
      : launch_agg<V>(xf, ldx, of, ldo, n_out_rows, rowptr, col, af, ld_addend, tile_info, n_tiles, max_tile_rows, max_tile_edges, st);
  switch (width / 32) {
    AX2D_AGG_CASE(1) AX2D_AGG_CASE(2) AX2D_AGG_CASE(3) AX2D_AGG_CASE(4) AX2D_AGG_CASE(5) AX2D_AGG_CASE(6)
    AX2D_AGG_CASE(7) AX2D_AGG_CASE(8) AX2D_AGG_CASE(9) AX2D_AGG_CASE(10) AX2D_AGG_CASE(11) AX2D_AGG_CASE(12)
    AX2D_AGG_CASE(13) AX2D_AGG_CASE(14) AX2D_AGG_CASE(15) AX2D_AGG_CASE(16)
  }
#undef AX2D_AGG_CASE
  return AX2D_ERR_UNSUPPORTED;
}
