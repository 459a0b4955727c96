// Collation-time integer work on the HOST (a9 and the "next" row f-1 of SURVEY.md section 8):
// stable CSR emission, tile planning, shell-edge BFS.  Exact integers, single-threaded C++; these run
// in DataLoader workers exactly where the reference runs MyBatch.from_data_list (loaders.py:38-45).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "common.cuh"

namespace ax2d {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<uint64_t> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }
// Programmatic dependent launch of our kernels (common.cuh launch_k): -1 = not decided yet (env AX2D_PDL=1, default off:
// measured neutral on the captured C2 step, 3.057 vs 3.067 ms -- the kernels of the graph already run back to back).
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("AX2D_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
}  // namespace ax2d

extern "C" int ax2d_set_pdl(int on) {
  const int prev = ax2d::pdl_enabled() ? 1 : 0;
  ax2d::g_pdl.store(on != 0 ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

extern "C" int ax2d_abi_version(void) { return 1; }
extern "C" uint64_t ax2d_launch_count(void) { return ax2d::g_launches.load(std::memory_order_relaxed); }

extern "C" const char* ax2d_last_error(void) { return ax2d::g_err; }

extern "C" const char* ax2d_error_string(int code) {
  switch (code) {
    case AX2D_OK: return "ok";
    case AX2D_ERR_ARG: return "bad argument (shape / range)";
    case AX2D_ERR_ALIGN: return "misaligned pointer";
    case AX2D_ERR_DTYPE: return "unsupported dtype";
    case AX2D_ERR_LAUNCH: return "CUDA launch failure";
    case AX2D_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
  }
}

// Stable counting sort by key -> CSR.  (The reference CPU scatter_add accumulates each output row in
// original edge order, layers.py:158-163; a stable sort by target reproduces that order exactly.)
extern "C" int ax2d_host_csr_build(const int64_t* edges, int64_t E, int64_t stride_e, int64_t stride_c, int64_t N,
                                   int64_t R, int transpose, int32_t* rowptr, int32_t* col, int32_t* perm) {
  using namespace ax2d;
  AX2D_CHECK_ARG(E >= 0 && N >= 0 && R >= 0 && rowptr != nullptr, "ax2d_host_csr_build: bad sizes");
  AX2D_CHECK_ARG(E < (1ll << 31) && R < (1ll << 31), "ax2d_host_csr_build: int32 index overflow");
  AX2D_CHECK_ARG(!transpose || R == N, "ax2d_host_csr_build: transposed CSR has N rows");
  memset(rowptr, 0, sizeof(int32_t) * static_cast<size_t>(R + 1));
  if (E == 0) return AX2D_OK;
  AX2D_CHECK_ARG(N > 0, "ax2d_host_csr_build: edges but no atoms");
  for (int64_t e = 0; e < E; ++e) {
    const int64_t t = edges[e * stride_e];
    const int64_t s = edges[e * stride_e + stride_c];
    AX2D_CHECK_ARG(s >= 0 && t >= 0, "ax2d_host_csr_build: negative index at edge %lld", (long long)e);
    const int64_t key = transpose ? (s % N) : t;
    AX2D_CHECK_ARG(key < R, "ax2d_host_csr_build: edge %lld key %lld out of range [0,%lld)", (long long)e,
                   (long long)key, (long long)R);
    rowptr[key + 1]++;
  }
  for (int64_t r = 0; r < R; ++r) rowptr[r + 1] += rowptr[r];
  std::vector<int32_t> cursor(rowptr, rowptr + R);
  for (int64_t e = 0; e < E; ++e) {
    const int64_t t = edges[e * stride_e];
    const int64_t s = edges[e * stride_e + stride_c] % N;
    const int64_t key = transpose ? s : t;
    const int32_t pos = cursor[key]++;
    col[pos] = static_cast<int32_t>(transpose ? t : s);
    if (perm != nullptr) perm[pos] = static_cast<int32_t>(e);
  }
  return AX2D_OK;
}

// unique[0] = 1 iff no (row, column) pair occurs twice in the CSR (one stamp per column: O(E)).
extern "C" int ax2d_host_csr_rows_unique(const int32_t* rowptr, const int32_t* col, int64_t R, int64_t N, int32_t* unique) {
  using namespace ax2d;
  AX2D_CHECK_ARG(rowptr != nullptr && unique != nullptr && R >= 0 && N >= 0, "ax2d_host_csr_rows_unique: bad arguments");
  unique[0] = 1;
  std::vector<int32_t> stamp(static_cast<size_t>(N), -1);
  for (int64_t r = 0; r < R; ++r) {
    for (int32_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      const int32_t c = col[e];
      AX2D_CHECK_ARG(c >= 0 && c < N, "ax2d_host_csr_rows_unique: column %d out of range at entry %d", c, e);
      if (stamp[c] == static_cast<int32_t>(r)) {
        unique[0] = 0;
        return AX2D_OK;
      }
      stamp[c] = static_cast<int32_t>(r);
    }
  }
  return AX2D_OK;
}

extern "C" int ax2d_host_tile_plan(const int32_t* seg_ptr, int64_t B, int64_t cap, const int32_t* rowptr,
                                   const int32_t* col, int32_t* tile_ptr, int64_t* n_tiles, int64_t* max_rows,
                                   int32_t* tile_local) {
  using namespace ax2d;
  AX2D_CHECK_ARG(seg_ptr != nullptr && tile_ptr != nullptr && n_tiles != nullptr && max_rows != nullptr && cap > 0,
                 "ax2d_host_tile_plan: bad arguments");
  int64_t nt = 0, mx = 0;
  tile_ptr[0] = seg_ptr[0];
  int64_t start = seg_ptr[0];
  for (int64_t g = 0; g < B; ++g) {
    const int64_t end = seg_ptr[g + 1];
    AX2D_CHECK_ARG(end >= seg_ptr[g], "ax2d_host_tile_plan: seg_ptr not monotone at %lld", (long long)g);
    // close the current tile before molecule g if adding it would overflow the cap
    if (seg_ptr[g] > start && end - start > cap) {
      tile_ptr[++nt] = seg_ptr[g];
      if (seg_ptr[g] - start > mx) mx = seg_ptr[g] - start;
      start = seg_ptr[g];
    }
  }
  if (B > 0 && seg_ptr[B] > start) {
    tile_ptr[++nt] = seg_ptr[B];
    if (seg_ptr[B] - start > mx) mx = seg_ptr[B] - start;
  }
  *n_tiles = nt;
  *max_rows = mx;
  if (tile_local != nullptr) {
    int32_t ok = 1;
    if (rowptr != nullptr && col != nullptr) {
      for (int64_t t = 0; t < nt && ok; ++t) {
        const int32_t lo = tile_ptr[t], hi = tile_ptr[t + 1];
        for (int32_t k = rowptr[lo]; k < rowptr[hi]; ++k)
          if (col[k] < lo || col[k] >= hi) {
            ok = 0;
            break;
          }
      }
    }
    *tile_local = ok;
  }
  return AX2D_OK;
}

// BFS in edge space, identical discovery order to datasets/features.py:97-150.
extern "C" int64_t ax2d_host_shell_edges(int64_t B, const int64_t* atom_ptr, const int64_t* bond_ptr,
                                         const int32_t* bonds, int num_hops, int64_t* hop_counts, int64_t* edges_out,
                                         int64_t capacity) {
  using namespace ax2d;
  if (B < 0 || num_hops < 1 || atom_ptr == nullptr || bond_ptr == nullptr || hop_counts == nullptr) {
    set_error("ax2d_host_shell_edges: bad arguments");
    return AX2D_ERR_ARG;
  }
  int64_t total = 0;
  std::vector<std::vector<int32_t>> nbr;
  std::vector<uint8_t> seen;
  std::vector<int32_t> fr_u, fr_v, nx_u, nx_v;
  for (int64_t g = 0; g < B; ++g) {
    const int64_t n = atom_ptr[g + 1] - atom_ptr[g];
    const int64_t off = atom_ptr[g];
    nbr.assign(static_cast<size_t>(n), {});
    std::vector<uint8_t> adj(static_cast<size_t>(n * n), 0);
    for (int64_t b = bond_ptr[g]; b < bond_ptr[g + 1]; ++b) {
      const int32_t a0 = bonds[2 * b], a1 = bonds[2 * b + 1];
      if (a0 < 0 || a1 < 0 || a0 >= n || a1 >= n) {
        set_error("ax2d_host_shell_edges: bond %lld of molecule %lld out of range", (long long)b, (long long)g);
        return AX2D_ERR_ARG;
      }
      if (a0 == a1) continue;
      adj[a0 * n + a1] = adj[a1 * n + a0] = 1;
    }
    for (int64_t v = 0; v < n; ++v)
      for (int64_t w = 0; w < n; ++w)
        if (adj[v * n + w]) nbr[v].push_back(static_cast<int32_t>(w));   // ascending, no self loop
    seen.assign(static_cast<size_t>(n * n), 0);
    fr_u.clear();
    fr_v.clear();
    for (int32_t v = 0; v < n; ++v)
      for (int32_t w : nbr[v])
        if (!seen[v * n + w]) {
          seen[v * n + w] = 1;
          fr_u.push_back(v);
          fr_v.push_back(w);
        }
    for (int h = 0; h < num_hops; ++h) {
      const int64_t cnt = static_cast<int64_t>(fr_u.size());
      hop_counts[g * num_hops + h] = cnt;
      if (edges_out != nullptr) {
        if (total + cnt > capacity) {
          set_error("ax2d_host_shell_edges: capacity %lld too small", (long long)capacity);
          return AX2D_ERR_ARG;
        }
        for (int64_t i = 0; i < cnt; ++i) {
          edges_out[2 * (total + i)] = fr_u[i] + off;       // target = BFS origin u   (quirk Q2)
          edges_out[2 * (total + i) + 1] = fr_v[i] + off;   // src    = reached atom w
        }
      }
      total += cnt;
      if (h + 1 == num_hops) break;
      nx_u.clear();
      nx_v.clear();
      for (int64_t i = 0; i < cnt; ++i) {
        const int32_t u = fr_u[i], v = fr_v[i];
        for (int32_t w : nbr[v])
          if (w != u && !seen[u * n + w]) {
            seen[u * n + w] = 1;
            nx_u.push_back(u);
            nx_v.push_back(w);
          }
      }
      fr_u.swap(nx_u);
      fr_v.swap(nx_v);
    }
  }
  return total;
}

// f-1: the shell-edge BFS emitting the CSR the aggregation kernels consume DIRECTLY (no edge list, no sort).
// For a target row u the reference's stable edge order is hop-major and, inside a hop, the BFS discovery order of u's
// frontier (features.py:97-150); the transposed row of a source w is hop-major with ASCENDING targets (the frontier lists
// are grouped by ascending origin).  Shell relations are symmetric (distance(u, w) == distance(w, u)), so both CSRs share
// one rowptr.  Equal, bit for bit, to ax2d_host_csr_build over the edge list of ax2d_host_shell_edges.
extern "C" int64_t ax2d_host_shell_csr(int64_t B, const int64_t* atom_ptr, const int64_t* bond_ptr, const int32_t* bonds,
                                       int num_hops, int32_t* rowptr, int32_t* col, int32_t* col_t, int64_t capacity) {
  using namespace ax2d;
  if (B < 0 || num_hops < 1 || atom_ptr == nullptr || bond_ptr == nullptr || rowptr == nullptr) {
    set_error("ax2d_host_shell_csr: bad arguments");
    return AX2D_ERR_ARG;
  }
  const bool fill = col != nullptr;
  if (fill && col_t == nullptr) {
    set_error("ax2d_host_shell_csr: col and col_t go together");
    return AX2D_ERR_ARG;
  }
  int64_t total = 0;
  rowptr[atom_ptr[0]] = 0;
  std::vector<std::vector<int32_t>> nbr;
  std::vector<int32_t> dist, order, next_order, hop_end;
  for (int64_t g = 0; g < B; ++g) {
    const int64_t n = atom_ptr[g + 1] - atom_ptr[g];
    const int64_t off = atom_ptr[g];
    if (n >= (1ll << 30) || off + n >= (1ll << 31)) {
      set_error("ax2d_host_shell_csr: int32 index overflow");
      return AX2D_ERR_ARG;
    }
    nbr.assign(static_cast<size_t>(n), {});
    std::vector<uint8_t> adj(static_cast<size_t>(n * n), 0);
    for (int64_t b = bond_ptr[g]; b < bond_ptr[g + 1]; ++b) {
      const int32_t a0 = bonds[2 * b], a1 = bonds[2 * b + 1];
      if (a0 < 0 || a1 < 0 || a0 >= n || a1 >= n) {
        set_error("ax2d_host_shell_csr: bond %lld of molecule %lld out of range", (long long)b, (long long)g);
        return AX2D_ERR_ARG;
      }
      if (a0 == a1) continue;
      adj[a0 * n + a1] = adj[a1 * n + a0] = 1;
    }
    for (int64_t v = 0; v < n; ++v)
      for (int64_t w = 0; w < n; ++w)
        if (adj[v * n + w]) nbr[v].push_back(static_cast<int32_t>(w));
    dist.assign(static_cast<size_t>(n), 0);
    for (int64_t u = 0; u < n; ++u) {
      // BFS from u in the reference's discovery order: hop h + 1 walks hop h in order, neighbours ascending
      std::fill(dist.begin(), dist.end(), -1);
      dist[u] = 0;
      order.assign(nbr[u].begin(), nbr[u].end());
      for (int32_t w : order) dist[w] = 1;
      const int64_t row0 = total;
      for (int h = 1; h <= num_hops && !order.empty(); ++h) {
        if (fill) {
          if (total + static_cast<int64_t>(order.size()) > capacity) {
            set_error("ax2d_host_shell_csr: capacity %lld too small", (long long)capacity);
            return AX2D_ERR_ARG;
          }
          for (size_t i = 0; i < order.size(); ++i) col[total + i] = static_cast<int32_t>(order[i] + off);
          // transposed row of u: the same hop set with ascending targets
          size_t k = 0;
          for (int64_t w = 0; w < n; ++w)
            if (dist[w] == h) col_t[total + k++] = static_cast<int32_t>(w + off);
        }
        total += static_cast<int64_t>(order.size());
        if (h == num_hops) break;
        next_order.clear();
        for (int32_t v : order)
          for (int32_t w : nbr[v])
            if (dist[w] < 0) {
              dist[w] = h + 1;
              next_order.push_back(w);
            }
        order.swap(next_order);
      }
      (void)row0;
      rowptr[off + u + 1] = static_cast<int32_t>(total);
    }
  }
  if (total >= (1ll << 31)) {
    set_error("ax2d_host_shell_csr: int32 index overflow");
    return AX2D_ERR_ARG;
  }
  return total;
}
