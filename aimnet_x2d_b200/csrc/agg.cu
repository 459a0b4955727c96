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

// generic width (multiple of 4, not of 32): one lane per float4 column, strided over the row
__global__ void __launch_bounds__(256) agg_generic_kernel(const float* __restrict__ x, int64_t ldx,
                                                          float* __restrict__ out, int64_t ldo,
                                                          const int32_t* __restrict__ rowptr,
                                                          const int32_t* __restrict__ col,
                                                          const float* __restrict__ addend, int64_t ld_addend,
                                                          int64_t n_out_rows, int w4) {
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
                      const int32_t* tile_ptr, int64_t n_tiles, int max_tile_rows, cudaStream_t st) {
  if (tile_ptr != nullptr) {
    const size_t smem = static_cast<size_t>(max_tile_rows) * V * 8 * 16;
    if (smem > 227 * 1024) {
      set_error("ax2d_agg: tile of %d rows needs %zu bytes of shared memory", max_tile_rows, smem);
      return AX2D_ERR_UNSUPPORTED;
    }
    auto kern = agg_kernel<V, true>;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    // one 8-lane group per row of the largest tile, rounded to whole warps, 64..256 threads
    int threads = ((max_tile_rows * 8 + 31) / 32) * 32;
    threads = threads < 64 ? 64 : (threads > 256 ? 256 : threads);
    kern<<<static_cast<unsigned>(n_tiles), threads, smem, st>>>(x, ldx, out, ldo, rowptr, col, addend, ld_addend,
                                                                 tile_ptr, n_out_rows, 0);
  } else {
    const int rows_per_cta = 32;
    const int64_t grid = (n_out_rows + rows_per_cta - 1) / rows_per_cta;
    agg_kernel<V, false><<<static_cast<unsigned>(grid), 256, 0, st>>>(x, ldx, out, ldo, rowptr, col, addend,
                                                                       ld_addend, nullptr, n_out_rows, rows_per_cta);
  }
  return launch_status("ax2d_agg");
}

}  // namespace ax2d

extern "C" int ax2d_agg(const void* x, int64_t ldx, int64_t n_src_rows, void* out, int64_t ldo, int64_t n_out_rows,
                        const int32_t* rowptr, const int32_t* col, const void* addend, int64_t ld_addend, int width,
                        const int32_t* tile_ptr, int64_t n_tiles, int max_tile_rows, int dtype,
                        ax2d_stream_t stream) {
  using namespace ax2d;
  if (dtype != AX2D_F32) {
    set_error("ax2d_agg: dtype %d not supported (f32 only in this build)", dtype);
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
  if (tile_ptr != nullptr) {
    AX2D_CHECK_ARG(n_out_rows == n_src_rows && ldx == width && width % 32 == 0 && n_tiles > 0 && max_tile_rows > 0,
                   "ax2d_agg: tiled mode needs n_out_rows == n_src_rows, ldx == width, width %% 32 == 0");
  }
  if (width % 32 != 0 || width > 32 * 16) {
    const int warps = 8;
    agg_generic_kernel<<<static_cast<unsigned>((n_out_rows + warps - 1) / warps), warps * 32, 0, st>>>(
        xf, ldx, of, ldo, rowptr, col, af, ld_addend, n_out_rows, width / 4);
    return launch_status("ax2d_agg");
  }
#define AX2D_AGG_CASE(V) \
  case V: return launch_agg<V>(xf, ldx, of, ldo, n_out_rows, rowptr, col, af, ld_addend, tile_ptr, n_tiles, max_tile_rows, st);
  switch (width / 32) {
    AX2D_AGG_CASE(1) AX2D_AGG_CASE(2) AX2D_AGG_CASE(3) AX2D_AGG_CASE(4) AX2D_AGG_CASE(5) AX2D_AGG_CASE(6)
    AX2D_AGG_CASE(7) AX2D_AGG_CASE(8) AX2D_AGG_CASE(9) AX2D_AGG_CASE(10) AX2D_AGG_CASE(11) AX2D_AGG_CASE(12)
    AX2D_AGG_CASE(13) AX2D_AGG_CASE(14) AX2D_AGG_CASE(15) AX2D_AGG_CASE(16)
  }
#undef AX2D_AGG_CASE
  return AX2D_ERR_UNSUPPORTED;
}
