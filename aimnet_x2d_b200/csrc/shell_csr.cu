// f-1 on the device: shell-edge BFS of a whole batch of molecules, emitting the forward and transposed CSR of the
// aggregation kernels directly from the bond lists (datasets/features.py:82-150 + the collation of molecular.py:426-438 +
// the stable sort of ax2d_host_csr_build in one step) -- bit-identical to ax2d_host_shell_csr.
//
// One warp per molecule.  The adjacency of a molecule is a bit matrix in shared memory (n rows of W 32-bit words, n <= 32 W
// <= 256; larger molecules take the host function); a lane owns target atoms u = lane, lane + 32, ... and runs the BFS
// from u on bitsets held in registers:
//   * count pass:  hop h + 1 = (OR of adj[v] over the hop-h set) & ~visited; the degree of u is the sum of the popcounts;
//   * a one-CTA scan turns the degrees into rowptr (shared by both CSRs: shell relations are symmetric);
//   * fill pass:   the reference's DISCOVERY order -- hop h + 1 walks hop h in order (read back from the col segment this
//     lane has just written) and appends the unvisited neighbours of each atom in ascending order -- goes to col; the
//     transposed row of u is the same hop set in ascending order (the bits of the hop's bitset).
#include "common.cuh"

namespace ax2d {

constexpr int SHELL_WARPS = 4;

template <int W>
__global__ void __launch_bounds__(32 * SHELL_WARPS) shell_bfs_kernel(const int32_t* __restrict__ atom_ptr,
                                                                     const int32_t* __restrict__ bond_ptr,
                                                                     const int32_t* __restrict__ bonds, int64_t B, int num_hops,
                                                                     int32_t* __restrict__ deg, const int32_t* __restrict__ rowptr,
                                                                     int32_t* __restrict__ col, int32_t* __restrict__ col_t,
                                                                     int* __restrict__ err) {
  __shared__ uint32_t s_adj[SHELL_WARPS][32 * W * W];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t g = static_cast<int64_t>(blockIdx.x) * SHELL_WARPS + warp;
  if (g >= B) return;
  const int off = atom_ptr[g];
  const int n = atom_ptr[g + 1] - off;
  if (n > 32 * W) {                      // the host sized W from the largest molecule: cannot happen with consistent inputs
    if (lane == 0) atomicExch(err, 1);
    return;
  }
  uint32_t* adj = s_adj[warp];
  for (int i = lane; i < n * W; i += 32) adj[i] = 0u;
  __syncwarp();
  for (int b = bond_ptr[g] + lane; b < bond_ptr[g + 1]; b += 32) {
    const int a0 = bonds[2 * b], a1 = bonds[2 * b + 1];
    if (a0 < 0 || a1 < 0 || a0 >= n || a1 >= n) {
      atomicExch(err, 2);
      continue;
    }
    if (a0 == a1) continue;
    atomicOr(&adj[a0 * W + (a1 >> 5)], 1u << (a1 & 31));
    atomicOr(&adj[a1 * W + (a0 >> 5)], 1u << (a0 & 31));
  }
  __syncwarp();
  const bool fill = col != nullptr;
  for (int u = lane; u < n; u += 32) {
    uint32_t visited[W], cur[W];
#pragma unroll
    for (int k = 0; k < W; ++k) {
      visited[k] = 0u;
      cur[k] = adj[u * W + k];
    }
    visited[u >> 5] |= 1u << (u & 31);
    int total = 0;
    int pos = fill ? rowptr[off + u] : 0;          // write cursor of this row
    int seg0 = pos;                                 // start of the previous hop's segment
    for (int h = 1; h <= num_hops; ++h) {
      int cnt = 0;
#pragma unroll
      for (int k = 0; k < W; ++k) cnt += __popc(cur[k]);
      if (cnt == 0) break;
      if (fill) {
        if (h == 1) {
          // hop 1: neighbours ascending = discovery order
#pragma unroll
          for (int k = 0; k < W; ++k) {
            uint32_t bits = cur[k];
            while (bits) {
              const int w = 32 * k + __ffs(bits) - 1;
              bits &= bits - 1u;
              col[pos++] = off + w;
            }
          }
        } else {
          // hop h: walk hop h - 1 in order, append the not yet visited neighbours of each atom in ascending order
          uint32_t left[W];
#pragma unroll
          for (int k = 0; k < W; ++k) left[k] = cur[k];
          const int seg1 = pos;
          for (int i = seg0; i < seg1; ++i) {
            const int v = col[i] - off;
#pragma unroll
            for (int k = 0; k < W; ++k) {
              uint32_t bits = adj[v * W + k] & left[k];
              left[k] &= ~bits;
              while (bits) {
                const int w = 32 * k + __ffs(bits) - 1;
                bits &= bits - 1u;
                col[pos++] = off + w;
              }
            }
          }
          seg0 = seg1;
        }
        // transposed row of u: the hop set with ascending targets, at the same positions
        int pt = pos - cnt;
#pragma unroll
        for (int k = 0; k < W; ++k) {
          uint32_t bits = cur[k];
          while (bits) {
            const int w = 32 * k + __ffs(bits) - 1;
            bits &= bits - 1u;
            col_t[pt++] = off + w;
          }
        }
      }
      total += cnt;
#pragma unroll
      for (int k = 0; k < W; ++k) visited[k] |= cur[k];
      if (h == num_hops) break;
      uint32_t nxt[W];
#pragma unroll
      for (int k = 0; k < W; ++k) nxt[k] = 0u;
#pragma unroll
      for (int kk = 0; kk < W; ++kk) {
        uint32_t bits = cur[kk];
        while (bits) {
          const int v = 32 * kk + __ffs(bits) - 1;
          bits &= bits - 1u;
#pragma unroll
          for (int k = 0; k < W; ++k) nxt[k] |= adj[v * W + k];
        }
      }
#pragma unroll
      for (int k = 0; k < W; ++k) cur[k] = nxt[k] & ~visited[k];
    }
    if (!fill) deg[off + u] = total;
  }
}

// rowptr[0] = 0, rowptr[i + 1] = deg[0] + ... + deg[i]: one CTA walks the array in chunks of 4096 with a running carry
// (featurisation-time work: ~0.2 ms per million atoms)
__global__ void __launch_bounds__(1024) shell_scan_kernel(const int32_t* __restrict__ deg, int64_t n, int32_t* __restrict__ rowptr,
                                                          int* __restrict__ err) {
  __shared__ int s_warp[32];
  __shared__ long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    s_carry = 0;
    rowptr[0] = 0;
  }
  __syncthreads();
  for (int64_t base = 0; base < n; base += 4096) {
    const int64_t i0 = base + 4 * static_cast<int64_t>(tid);
    int v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = i0 + k < n ? deg[i0 + k] : 0;
    const int mine = v[0] + v[1] + v[2] + v[3];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;                        // inclusive over the warps
    }
    __syncthreads();
    const long long carry = s_carry;
    long long run = carry + (warp > 0 ? s_warp[warp - 1] : 0) + (incl - mine);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      run += v[k];
      if (i0 + k < n) rowptr[i0 + k + 1] = static_cast<int32_t>(run);
    }
    __syncthreads();
    if (tid == 1023) {
      s_carry = run;
      if (run >= (1ll << 31)) atomicExch(err, 3);
    }
    __syncthreads();
  }
}

template <int W>
static int launch_shell(const int32_t* atom_ptr, const int32_t* bond_ptr, const int32_t* bonds, int64_t B, int num_hops,
                        int32_t* deg, const int32_t* rowptr, int32_t* col, int32_t* col_t, int* err, cudaStream_t st) {
  const int64_t grid = (B + SHELL_WARPS - 1) / SHELL_WARPS;
  shell_bfs_kernel<W><<<static_cast<unsigned>(grid), 32 * SHELL_WARPS, 0, st>>>(atom_ptr, bond_ptr, bonds, B, num_hops, deg, rowptr,
                                                                                 col, col_t, err);
  return launch_status("ax2d_shell_csr");
}

static int dispatch_shell(int max_atoms, const int32_t* atom_ptr, const int32_t* bond_ptr, const int32_t* bonds, int64_t B,
                          int num_hops, int32_t* deg, const int32_t* rowptr, int32_t* col, int32_t* col_t, int* err,
                          cudaStream_t st) {
  if (max_atoms <= 32) return launch_shell<1>(atom_ptr, bond_ptr, bonds, B, num_hops, deg, rowptr, col, col_t, err, st);
  if (max_atoms <= 64) return launch_shell<2>(atom_ptr, bond_ptr, bonds, B, num_hops, deg, rowptr, col, col_t, err, st);
  if (max_atoms <= 128) return launch_shell<4>(atom_ptr, bond_ptr, bonds, B, num_hops, deg, rowptr, col, col_t, err, st);
  return launch_shell<8>(atom_ptr, bond_ptr, bonds, B, num_hops, deg, rowptr, col, col_t, err, st);
}

}  // namespace ax2d

using namespace ax2d;

// Pass 1: degrees + rowptr.  err: one device int, zeroed by the caller, set non-zero on inconsistent inputs (1: a molecule
// larger than max_atoms, 2: bond index out of range, 3: more than 2^31 edges).  workspace: N int32.
extern "C" int ax2d_shell_csr_count(const int32_t* atom_ptr, const int32_t* bond_ptr, const int32_t* bonds, int64_t B, int64_t N,
                                    int max_atoms, int num_hops, int32_t* rowptr, int32_t* workspace, int32_t* err,
                                    ax2d_stream_t stream) {
  AX2D_CHECK_ARG(atom_ptr != nullptr && bond_ptr != nullptr && rowptr != nullptr && workspace != nullptr && err != nullptr,
                 "ax2d_shell_csr_count: null argument");
  AX2D_CHECK_ARG(B >= 0 && N >= 0 && num_hops >= 1 && max_atoms >= 0, "ax2d_shell_csr_count: bad sizes");
  AX2D_CHECK_ARG(max_atoms <= 256, "ax2d_shell_csr_count: molecules of more than 256 atoms (%d) take the host path "
                                   "(ax2d_host_shell_csr)", max_atoms);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (B > 0) {
    const int rc = dispatch_shell(max_atoms, atom_ptr, bond_ptr, bonds, B, num_hops, workspace, nullptr, nullptr, nullptr, err, st);
    if (rc != AX2D_OK) return rc;
  }
  shell_scan_kernel<<<1, 1024, 0, st>>>(workspace, N, rowptr, err);
  return launch_status("ax2d_shell_csr_count");
}

// Pass 2: col (forward CSR, discovery order) and col_t (transposed CSR, ascending targets) at the offsets of rowptr.
extern "C" int ax2d_shell_csr_fill(const int32_t* atom_ptr, const int32_t* bond_ptr, const int32_t* bonds, int64_t B,
                                   int max_atoms, int num_hops, const int32_t* rowptr, int32_t* col, int32_t* col_t, int32_t* err,
                                   ax2d_stream_t stream) {
  AX2D_CHECK_ARG(atom_ptr != nullptr && bond_ptr != nullptr && rowptr != nullptr && col != nullptr && col_t != nullptr &&
                     err != nullptr,
                 "ax2d_shell_csr_fill: null argument");
  AX2D_CHECK_ARG(B >= 0 && num_hops >= 1 && max_atoms >= 0 && max_atoms <= 256, "ax2d_shell_csr_fill: bad sizes");
  if (B == 0) return AX2D_OK;
  return dispatch_shell(max_atoms, atom_ptr, bond_ptr, bonds, B, num_hops, nullptr, rowptr, col, col_t, err,
                        reinterpret_cast<cudaStream_t>(stream));
}
