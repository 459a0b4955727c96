// a6 embedding lookups (models/gnn.py:262-274), a10/a11 post-allreduce clip + Adam on the flat arena
// (training/trainer.py:164-165), a12 weighted L1 / MSE loss (models/losses.py:14-87).
#include "common.cuh"

namespace ax2d {

constexpr int kMaxTables = 8;
struct EmbedArgs {
  const float* table[kMaxTables];
  const int64_t* index[kMaxTables];
};

// block = (64 float4 columns, 4 rows); a CTA covers kEmbedRowsPerCta atoms (no 64-bit divisions per element)
constexpr int kEmbedRowsPerCta = 32;
__global__ void __launch_bounds__(256) embed_fwd_kernel(EmbedArgs a, int n_tables, int E4, int64_t N,
                                                        float* __restrict__ out, int64_t ldo) {
  pdl_enter();
  const int per_row = n_tables * E4;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * kEmbedRowsPerCta;
  for (int c = threadIdx.x; c < per_row; c += 64) {
    const int tb = c / E4, k = c - tb * E4;
    const int64_t* idx = a.index[tb];
    const float4* tab = reinterpret_cast<const float4*>(a.table[tb]);
#pragma unroll 4
    for (int r = threadIdx.y; r < kEmbedRowsPerCta; r += 4) {
      const int64_t n = n0 + r;
      if (n < N) reinterpret_cast<float4*>(out + n * ldo)[c] = __ldg(tab + __ldg(idx + n) * E4 + k);
    }
  }
}

// bf16 output (BASELINE configs[3]): the fp32 table rows are rounded on the way out; 8 columns (16 bytes) per thread
__global__ void __launch_bounds__(256) embed_fwd_bf16_kernel(EmbedArgs a, int n_tables, int E8, int64_t N,
                                                             uint16_t* __restrict__ out, int64_t ldo) {
  pdl_enter();
  const int per_row = n_tables * E8;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * kEmbedRowsPerCta;
  for (int c = threadIdx.x; c < per_row; c += 64) {
    const int tb = c / E8, k = c - tb * E8;
    const int64_t* idx = a.index[tb];
    const float4* tab = reinterpret_cast<const float4*>(a.table[tb]);
#pragma unroll 4
    for (int r = threadIdx.y; r < kEmbedRowsPerCta; r += 4) {
      const int64_t n = n0 + r;
      if (n < N) {
        const float4* src = tab + (__ldg(idx + n) * E8 + k) * 2;
        const float4 u = __ldg(src), v = __ldg(src + 1);
        uint4 o;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.x) : "f"(u.y), "f"(u.x));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.y) : "f"(u.w), "f"(u.z));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.z) : "f"(v.y), "f"(v.x));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(o.w) : "f"(v.w), "f"(v.z));
        reinterpret_cast<uint4*>(out + n * ldo)[c] = o;
      }
    }
  }
}

// Backward of all lookups in one pass: thread = gradient column (table t, component e), a CTA walks its slice of the
// atoms in order and accumulates g[n, column] into a shared-memory copy of the (small) tables at row index[t][n]
// -- the same thread owns a column for all atoms, so there are no conflicts and the order is the atom order.  Per-CTA
// partial tables go to the workspace and are summed in CTA order by embed_bwd_all_final_kernel (deterministic).
struct EmbedBwdArgs {
  const int64_t* index[kMaxTables];
  float* g_table[kMaxTables];
  int row_off[kMaxTables + 1];     // first row of every table in the stacked [total_rows, E] layout
};
__device__ __forceinline__ float embed_g(const float* p) { return __ldg(p); }
__device__ __forceinline__ float embed_g(const uint16_t* p) { return __uint_as_float(static_cast<uint32_t>(__ldg(p)) << 16); }
template <typename GT>
__global__ void __launch_bounds__(256) embed_bwd_all_kernel(EmbedBwdArgs a, const GT* __restrict__ g, int64_t ldg, int n_tables,
                                                            int E, int64_t N, int64_t rows_per_cta, float* __restrict__ ws) {
  pdl_enter();
  extern __shared__ float tab[];          // [total_rows][E]
  const int total = a.row_off[n_tables] * E;
  for (int i = threadIdx.x; i < total; i += blockDim.x) tab[i] = 0.f;
  __syncthreads();
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < N ? r0 + rows_per_cta : N;
  for (int col = threadIdx.x; col < n_tables * E; col += blockDim.x) {
    const int t = col / E, e = col - t * E;
    const int64_t* idx = a.index[t];
    float* base = tab + a.row_off[t] * E + e;
    int64_t r = r0;
    // loads first, then the (possibly same-address) updates in atom order (16 rows per batch instead of 4 measured
    // SLOWER, 49 -> 55 us on C2: the batches are not what the kernel waits for)
    for (; r + 4 <= r1; r += 4) {
      const int64_t i0 = __ldg(idx + r), i1 = __ldg(idx + r + 1), i2 = __ldg(idx + r + 2), i3 = __ldg(idx + r + 3);
      const float g0 = embed_g(g + r * ldg + col), g1 = embed_g(g + (r + 1) * ldg + col);
      const float g2 = embed_g(g + (r + 2) * ldg + col), g3 = embed_g(g + (r + 3) * ldg + col);
      base[i0 * E] += g0;
      base[i1 * E] += g1;
      base[i2 * E] += g2;
      base[i3 * E] += g3;
    }
    for (; r < r1; ++r) base[__ldg(idx + r) * E] += embed_g(g + r * ldg + col);
  }
  __syncthreads();
  float* dst = ws + static_cast<int64_t>(blockIdx.x) * total;
  for (int i = threadIdx.x; i < total; i += blockDim.x) dst[i] = tab[i];
}
// eight lanes per output element: lane z sums the partials z, z + 8, ..., a shuffle tree combines them (fixed order)
__global__ void __launch_bounds__(256) embed_bwd_all_final_kernel(EmbedBwdArgs a, const float* __restrict__ ws, int n_tables, int E,
                                                                  int n_part) {
  pdl_enter();
  const int total = a.row_off[n_tables] * E;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int z0 = threadIdx.x & 7;
  float s = 0.f;
  if (i < total) {
#pragma unroll 4
    for (int p = z0; p < n_part; p += 8) s += __ldg(ws + static_cast<int64_t>(p) * total + i);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (i >= total || z0 != 0) return;
  const int row = i / E;
  int t = 0;
  while (t + 1 < n_tables && row >= a.row_off[t + 1]) ++t;
  a.g_table[t][i - a.row_off[t] * E] = s;
}

// Two-level fixed-order segment sum over atoms sorted (stably) by table row.
// grid = (vocab, splits): CTA (v, s) sums its slice of row v's atoms.  256 threads = 16 float4 column lanes x 16 row
// lanes: row lane y takes atoms lo + y, lo + y + 16, ... (each a coalesced 16 x 16 B read), the 16 partial sums are
// combined through shared memory in row-lane order.  emb_dim <= 64 per pass (wider tables loop over column blocks).
__global__ void __launch_bounds__(256) embed_bwd_partial_kernel(const float* __restrict__ g_out, int64_t ldg,
                                                                int col0, int E4, const int32_t* __restrict__ order,
                                                                const int32_t* __restrict__ ptr, int splits,
                                                                float* __restrict__ partial) {
  pdl_enter();
  __shared__ float4 red[16][16];
  const int v = blockIdx.x, s = blockIdx.y;
  const int beg = ptr[v], end = ptr[v + 1];
  const int len = end - beg;
  const int chunk = (len + splits - 1) / splits;
  const int lo = beg + s * chunk;
  const int hi = lo + chunk < end ? lo + chunk : end;
  const int cx = threadIdx.x & 15, ry = threadIdx.x >> 4;
  for (int c0 = 0; c0 < E4; c0 += 16) {
    const int c = c0 + cx;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < E4) {
#pragma unroll 4
      for (int k = lo + ry; k < hi; k += 16) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(g_out + static_cast<int64_t>(__ldg(order + k)) * ldg + col0) + c);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
    }
    red[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && c < E4) {
      float4 t = red[0][cx];
#pragma unroll
      for (int y = 1; y < 16; ++y) {
        const float4 u = red[y][cx];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      reinterpret_cast<float4*>(partial + (static_cast<int64_t>(v) * splits + s) * E4 * 4)[c] = t;
    }
    __syncthreads();
  }
}
__global__ void embed_bwd_final_kernel(const float* __restrict__ partial, int splits, int E, int64_t vocab,
                                       float* __restrict__ g_table) {
  pdl_enter();
  const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= vocab * E) return;
  const int64_t v = t / E;
  const int c = static_cast<int>(t % E);
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partial[(v * splits + k) * E + c];
  g_table[t] = s;
}

// ----------------------------------------------------------------------------------------- optimiser
constexpr int kNormBlocks = 1024;
__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const float* __restrict__ g, int64_t n,
                                                             float* __restrict__ partial) {
  pdl_enter();
  __shared__ float red[8];
  float s = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[(n4 << 2) + threadIdx.x];
    s += v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) sqnorm_final_kernel(const float* __restrict__ partial, int n_part,
                                                           float* __restrict__ norm2) {
  pdl_enter();
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n_part; i += blockDim.x) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    norm2[0] = t;
  }
}

__global__ void step_inc_kernel(int64_t* step) {
  pdl_enter(); step[0] += 1; }

__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                        const float* __restrict__ norm2, float grad_scale,
                                                        float max_norm, float lr, float beta1, float beta2, float eps,
                                                        const int64_t* __restrict__ step) {
  pdl_enter();
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = static_cast<double>(step[0]);
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), t);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), t);
    s_step_size = static_cast<float>(static_cast<double>(lr) / bc1);
    s_bc2_sqrt = static_cast<float>(sqrt(bc2));
    float coef = grad_scale;
    if (max_norm > 0.f) {                                   // clip_grad_norm_: coef = clamp(max/(norm+1e-6), max=1)
      const float total = sqrtf(norm2[0]) * grad_scale;
      const float c = max_norm / (total + 1e-6f);
      coef *= c < 1.f ? c : 1.f;
    }
    s_coef = coef;
  }
  __syncthreads();
  const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float w1 = 1.f - beta1, w2 = 1.f - beta2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = m[i] + (gi - m[i]) * w1;               // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * beta2 + w2 * gi * gi;           // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);                 // param.addcdiv_(exp_avg, denom, value=-step_size)
  }
}

// The same update with every hyper-parameter read from DEVICE memory, hyper = {grad_scale, max_norm, lr, beta1, beta2,
// eps}: a CUDA-graph-captured step then follows learning-rate schedulers (the host rewrites the six floats between
// replays) instead of replaying the values that were current at capture time.
__global__ void __launch_bounds__(256) clip_adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                            float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                            const float* __restrict__ norm2, const float* __restrict__ hyper,
                                                            const int64_t* __restrict__ step) {
  pdl_enter();
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  const float grad_scale = hyper[0], max_norm = hyper[1], lr = hyper[2], beta1 = hyper[3], beta2 = hyper[4], eps = hyper[5];
  if (threadIdx.x == 0) {
    const double t = static_cast<double>(step[0]);
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), t);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), t);
    s_step_size = static_cast<float>(static_cast<double>(lr) / bc1);
    s_bc2_sqrt = static_cast<float>(sqrt(bc2));
    float coef = grad_scale;
    if (max_norm > 0.f) {
      const float total = sqrtf(norm2[0]) * grad_scale;
      const float c = max_norm / (total + 1e-6f);
      coef *= c < 1.f ? c : 1.f;
    }
    s_coef = coef;
  }
  __syncthreads();
  const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float w1 = 1.f - beta1, w2 = 1.f - beta2;
  auto upd = [&](float gi, float& mi, float& vi, float& pi) {
    gi *= coef;
    mi = mi + (gi - mi) * w1;
    vi = vi * beta2 + w2 * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
  };
  // 128-bit accesses over the flat arena (the same per-element arithmetic); scalar tail / misaligned arenas below
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  const int64_t n4 = vec ? n >> 2 : 0;
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x, nthr = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = tid; i < n4; i += nthr) {
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    upd(g4.x, m4.x, v4.x, p4.x);
    upd(g4.y, m4.y, v4.y, p4.y);
    upd(g4.z, m4.z, v4.z, p4.z);
    upd(g4.w, m4.w, v4.w, p4.w);
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    reinterpret_cast<float4*>(p)[i] = p4;
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += nthr) {
    float mi = m[i], vi = v[i], pi = p[i];
    upd(g[i], mi, vi, pi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
  }
}

// ----------------------------------------------------------------------------------------- loss
__global__ void __launch_bounds__(1024) weighted_loss_kernel(const float* __restrict__ pred,
                                                             const float* __restrict__ target,
                                                             const float* __restrict__ weights, int64_t B, int T,
                                                             int kind, float* __restrict__ loss,
                                                             float* __restrict__ g_pred) {
  pdl_enter();
  __shared__ float red[32];
  const float invB = 1.f / static_cast<float>(B);
  float s = 0.f;
  // 32-bit index arithmetic (the host checks B * T < 2^31): a 64-bit modulo per element made this tiny kernel 20 us
  const int total = static_cast<int>(B) * T;
  int col = threadIdx.x % T;                       // weight index of element i, advanced without a modulo per element
  const int step = blockDim.x % T;
#pragma unroll 4
  for (int i = threadIdx.x; i < total; i += blockDim.x, col = col + step >= T ? col + step - T : col + step) {
    const float d = __ldg(pred + i) - __ldg(target + i);
    const float w = __ldg(weights + col);
    if (kind == 0) {
      s += fabsf(d) * w;
      if (g_pred != nullptr) g_pred[i] = (d > 0.f ? w : (d < 0.f ? -w : 0.f)) * invB;
    } else {
      s += d * d * w;
      if (g_pred != nullptr) g_pred[i] = 2.f * d * w * invB;
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    loss[0] = t * invB;
  }
}

}  // namespace ax2d

using namespace ax2d;

extern "C" int ax2d_embed_fwd(const float* const* tables, const int64_t* const* indices, int n_tables, int emb_dim,
                              int64_t N, float* out, int64_t ldo, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(n_tables >= 1 && n_tables <= kMaxTables, "ax2d_embed_fwd: n_tables=%d not in [1,%d]", n_tables, kMaxTables);
  AX2D_CHECK_ARG(emb_dim > 0 && emb_dim % 4 == 0 && ldo % 4 == 0, "ax2d_embed_fwd: emb_dim and ldo must be multiples of 4");
  AX2D_CHECK_ALIGN(out);
  if (N <= 0) return AX2D_OK;
  EmbedArgs a;
  for (int t = 0; t < n_tables; ++t) {
    AX2D_CHECK_ALIGN(tables[t]);
    a.table[t] = tables[t];
    a.index[t] = indices[t];
  }
  launch_k(embed_fwd_kernel, dim3(static_cast<unsigned>((N + kEmbedRowsPerCta - 1) / kEmbedRowsPerCta)), dim3(64, 4), 0, reinterpret_cast<cudaStream_t>(stream), a, n_tables, emb_dim / 4, N, out, ldo);
  return launch_status("ax2d_embed_fwd");
}

extern "C" int ax2d_embed_fwd_bf16(const float* const* tables, const int64_t* const* indices, int n_tables, int emb_dim,
                                   int64_t N, void* out, int64_t ldo, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(n_tables >= 1 && n_tables <= kMaxTables, "ax2d_embed_fwd_bf16: n_tables=%d not in [1,%d]", n_tables, kMaxTables);
  AX2D_CHECK_ARG(emb_dim > 0 && emb_dim % 8 == 0 && ldo % 8 == 0, "ax2d_embed_fwd_bf16: emb_dim and ldo must be multiples of 8");
  AX2D_CHECK_ALIGN(out);
  if (N <= 0) return AX2D_OK;
  EmbedArgs a;
  for (int t = 0; t < n_tables; ++t) {
    AX2D_CHECK_ALIGN(tables[t]);
    a.table[t] = tables[t];
    a.index[t] = indices[t];
  }
  launch_k(embed_fwd_bf16_kernel, dim3(static_cast<unsigned>((N + kEmbedRowsPerCta - 1) / kEmbedRowsPerCta)), dim3(64, 4), 0, reinterpret_cast<cudaStream_t>(stream), a, n_tables, emb_dim / 8, N, static_cast<uint16_t*>(out), ldo);
  return launch_status("ax2d_embed_fwd_bf16");
}

static int embed_bwd_all_ctas(int64_t N) {
  const int64_t by_rows = (N + 63) / 64;               // at least 64 atoms per CTA
  return static_cast<int>(by_rows < kNumSMs ? (by_rows < 1 ? 1 : by_rows) : kNumSMs);
}
extern "C" int64_t ax2d_embed_bwd_all_workspace(int64_t N, int64_t total_rows, int emb_dim) {
  return static_cast<int64_t>(embed_bwd_all_ctas(N)) * total_rows * emb_dim * 4;
}
static int embed_bwd_all_impl(const void* g_out, int g_bf16, int64_t ldg, int n_tables, int emb_dim, int64_t N,
                              const int64_t* const* indices, const int64_t* vocab, float* const* g_tables, void* workspace,
                              ax2d_stream_t stream) {
  AX2D_CHECK_ARG(n_tables > 0 && n_tables <= kMaxTables && emb_dim > 0 && N >= 0, "ax2d_embed_bwd_all: bad arguments");
  AX2D_CHECK_ARG(g_out != nullptr && workspace != nullptr, "ax2d_embed_bwd_all: null operand");
  EmbedBwdArgs a;
  int rows = 0;
  for (int t = 0; t < n_tables; ++t) {
    AX2D_CHECK_ARG(indices[t] != nullptr && g_tables[t] != nullptr && vocab[t] > 0, "ax2d_embed_bwd_all: null table operand");
    a.index[t] = indices[t];
    a.g_table[t] = g_tables[t];
    a.row_off[t] = rows;
    rows += static_cast<int>(vocab[t]);
  }
  a.row_off[n_tables] = rows;
  const size_t smem = static_cast<size_t>(rows) * emb_dim * 4;
  AX2D_CHECK_ARG(smem <= 200 * 1024, "ax2d_embed_bwd_all: tables of %d rows x %d do not fit shared memory (use ax2d_embed_bwd)",
                 rows, emb_dim);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(embed_bwd_all_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(embed_bwd_all_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
      set_error("ax2d_embed_bwd_all: cannot raise the dynamic shared memory limit to %zu: %s", smem, cudaGetErrorString(e));
      return AX2D_ERR_LAUNCH;
    }
    configured = smem;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int ctas = embed_bwd_all_ctas(N);
  const int64_t rows_per_cta = (N + ctas - 1) / ctas;
  if (g_bf16)
    launch_k(embed_bwd_all_kernel<uint16_t>, dim3(ctas), dim3(256), smem, st, a, static_cast<const uint16_t*>(g_out), ldg, n_tables, emb_dim, N,
                                                            rows_per_cta, static_cast<float*>(workspace));
  else
    launch_k(embed_bwd_all_kernel<float>, dim3(ctas), dim3(256), smem, st, a, static_cast<const float*>(g_out), ldg, n_tables, emb_dim, N, rows_per_cta,
                                                         static_cast<float*>(workspace));
  int rc = launch_status("ax2d_embed_bwd_all(partial)");
  if (rc != AX2D_OK) return rc;
  const int total = rows * emb_dim;
  launch_k(embed_bwd_all_final_kernel, dim3((8 * total + 255) / 256), dim3(256), 0, st, a, static_cast<const float*>(workspace), n_tables, emb_dim,
                                                                      ctas);
  return launch_status("ax2d_embed_bwd_all(final)");
}

extern "C" int ax2d_embed_bwd_all(const float* g_out, int64_t ldg, int n_tables, int emb_dim, int64_t N,
                                  const int64_t* const* indices, const int64_t* vocab, float* const* g_tables, void* workspace,
                                  ax2d_stream_t stream) {
  return embed_bwd_all_impl(g_out, 0, ldg, n_tables, emb_dim, N, indices, vocab, g_tables, workspace, stream);
}
extern "C" int ax2d_embed_bwd_all_bf16(const void* g_out, int64_t ldg, int n_tables, int emb_dim, int64_t N,
                                       const int64_t* const* indices, const int64_t* vocab, float* const* g_tables, void* workspace,
                                       ax2d_stream_t stream) {
  return embed_bwd_all_impl(g_out, 1, ldg, n_tables, emb_dim, N, indices, vocab, g_tables, workspace, stream);
}

constexpr int kEmbedSplits = 64;
extern "C" int64_t ax2d_embed_bwd_workspace(int64_t vocab, int emb_dim) {
  return vocab * kEmbedSplits * static_cast<int64_t>(emb_dim) * 4;
}
extern "C" int ax2d_embed_bwd(const float* g_out, int64_t ldg, int table, int emb_dim, int64_t vocab,
                              const int32_t* order, const int32_t* ptr, float* g_table, void* workspace,
                              ax2d_stream_t stream) {
  AX2D_CHECK_ARG(emb_dim > 0 && emb_dim % 4 == 0 && ldg % 4 == 0, "ax2d_embed_bwd: emb_dim and ldg must be multiples of 4");
  AX2D_CHECK_ARG(workspace != nullptr && vocab > 0, "ax2d_embed_bwd: workspace required");
  AX2D_CHECK_ALIGN(g_out);
  AX2D_CHECK_ALIGN(workspace);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid(static_cast<unsigned>(vocab), kEmbedSplits);
  launch_k(embed_bwd_partial_kernel, dim3(grid), dim3(256), 0, st, g_out, ldg, table * emb_dim, emb_dim / 4, order, ptr, kEmbedSplits,
                                                static_cast<float*>(workspace));
  int rc = launch_status("ax2d_embed_bwd(partial)");
  if (rc != AX2D_OK) return rc;
  const int64_t total = vocab * emb_dim;
  launch_k(embed_bwd_final_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, 
      static_cast<const float*>(workspace), kEmbedSplits, emb_dim, vocab, g_table);
  return launch_status("ax2d_embed_bwd(final)");
}

extern "C" int64_t ax2d_sqnorm_workspace(int64_t n) {
  (void)n;
  return kNormBlocks * 4;
}
extern "C" int ax2d_sqnorm(const float* g, int64_t n, float* norm2, int64_t* step_inc, void* workspace,
                           ax2d_stream_t stream) {
  AX2D_CHECK_ARG(n >= 0 && workspace != nullptr, "ax2d_sqnorm: bad arguments");
  AX2D_CHECK_ALIGN(g);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int64_t blocks = (n / 4 + 255) / 256;
  blocks = blocks < 1 ? 1 : (blocks > kNormBlocks ? kNormBlocks : blocks);
  launch_k(sqnorm_partial_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, st, g, n, static_cast<float*>(workspace));
  launch_k(sqnorm_final_kernel, dim3(1), dim3(256), 0, st, static_cast<const float*>(workspace), static_cast<int>(blocks), norm2);
  if (step_inc != nullptr) launch_k(step_inc_kernel, dim3(1), dim3(1), 0, st, step_inc);
  return launch_status("ax2d_sqnorm", step_inc != nullptr ? 3 : 2);
}

extern "C" int ax2d_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* norm2,
                              float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                              const int64_t* step, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(n >= 0 && step != nullptr && (max_norm <= 0.f || norm2 != nullptr), "ax2d_clip_adam: bad arguments");
  if (n == 0) return AX2D_OK;
  int64_t blocks = (n + 255) / 256;
  blocks = blocks > kNumSMs * 8 ? kNumSMs * 8 : blocks;
  launch_k(clip_adam_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      p, g, m, v, n, norm2, grad_scale, max_norm, lr, beta1, beta2, eps, step);
  return launch_status("ax2d_clip_adam");
}

extern "C" int ax2d_clip_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* norm2,
                                  const float* hyper, const int64_t* step, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(n >= 0 && step != nullptr && hyper != nullptr && norm2 != nullptr, "ax2d_clip_adam_dev: bad arguments");
  if (n == 0) return AX2D_OK;
  int64_t blocks = (n + 255) / 256;
  blocks = blocks > kNumSMs * 8 ? kNumSMs * 8 : blocks;
  launch_k(clip_adam_dev_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p, g, m, v, n, norm2,
                                                                                                            hyper, step);
  return launch_status("ax2d_clip_adam_dev");
}

extern "C" int ax2d_weighted_loss(const float* pred, const float* target, const float* weights, int64_t B, int T,
                                  int kind, float* loss, float* g_pred, ax2d_stream_t stream) {
  AX2D_CHECK_ARG(B > 0 && T > 0 && (kind == 0 || kind == 1) && B * T < (1ll << 31), "ax2d_weighted_loss: bad arguments");
  launch_k(weighted_loss_kernel, dim3(1), dim3(1024), 0, reinterpret_cast<cudaStream_t>(stream), pred, target, weights, B, T, kind, loss,
                                                                               g_pred);
  return launch_status("ax2d_weighted_loss");
}

// ---------------------------------------------------------------------------------------------- packed weights
// One launch per step instead of ~190 tiny ones: every projection weight of the model is gathered from its
// reference-shaped parameter into the padded / packed layout the kernels consume (zero padding lives in the
// destination and is never written), split into the two TF32 terms of the tensor-core path (plain and transposed:
// forward / data-gradient operands), and -- the other way round -- the packed weight gradients are accumulated
// into the parameters' .grad storage.
struct PackDesc {           // one rectangular block; 128 bytes (packed.py mirrors the layout)
  const float* src;         // parameter block [rows, cols], leading dimension src_ld
  float* grad_dst;          // same block inside the parameter's .grad (may be null: frozen parameter)
  float* w;                 // packed fp32 copy              [.., dst_ld]
  float* hi; float* lo;     // TF32 split of it (may be null: bias vectors)
  float* hiT; float* loT;   // transposed split              [.., dstT_ld] (may be null)
  const float* g;           // packed gradient block          [.., dst_ld]
  const float* part;        // split-K partials of the whole packed matrix [split][p_rows][p_cols] left behind by the
                            // weight-gradient kernel (null / split == 0: none); summed here in split order
  int32_t rows, cols, src_ld, dst_ld, dstT_ld;
  int32_t split, p_rows, p_cols, dr, dc;   // (dr, dc): position of this block inside the packed matrix
  uint16_t* wb;             // bf16 copy (operand of ax2d_gemm_bf16)            [.., dst_ld]   (may be null)
  uint16_t* wbT;            // transposed bf16 copy (data-gradient operand)      [.., dstT_ld]  (may be null)
};
__device__ __forceinline__ uint16_t pk_bf16(float v) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(v));
  return static_cast<uint16_t>(r & 0xFFFFu);
}
__device__ __forceinline__ float pk_round_tf32(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);
}
// 32 x 32 tiles through shared memory so that the transposed copies are written as coalesced rows too
__global__ void __launch_bounds__(256) pack_weights_kernel(const PackDesc* __restrict__ table) {
  pdl_enter();
  __shared__ float tile[32][33];
  const PackDesc d = table[blockIdx.y];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int tiles_c = (d.cols + 31) / 32, tiles_r = (d.rows + 31) / 32;
  for (int t = blockIdx.x; t < tiles_r * tiles_c; t += gridDim.x) {
    const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + ty + 8 * k, c = c0 + tx;
      float v = 0.f;
      if (r < d.rows && c < d.cols) {
        v = d.src[static_cast<int64_t>(r) * d.src_ld + c];
        d.w[static_cast<int64_t>(r) * d.dst_ld + c] = v;
        if (d.hi != nullptr) {
          const float h = pk_round_tf32(v);
          d.hi[static_cast<int64_t>(r) * d.dst_ld + c] = h;
          d.lo[static_cast<int64_t>(r) * d.dst_ld + c] = pk_round_tf32(v - h);
        }
        if (d.wb != nullptr) d.wb[static_cast<int64_t>(r) * d.dst_ld + c] = pk_bf16(v);
      }
      tile[ty + 8 * k][tx] = v;
    }
    __syncthreads();
    if (d.hiT != nullptr || d.wbT != nullptr) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;      // transposed: row c of the output, column r
        if (r < d.rows && c < d.cols) {
          const float v = tile[tx][ty + 8 * k];
          if (d.hiT != nullptr) {
            const float h = pk_round_tf32(v);
            d.hiT[static_cast<int64_t>(c) * d.dstT_ld + r] = h;
            d.loT[static_cast<int64_t>(c) * d.dstT_ld + r] = pk_round_tf32(v - h);
          }
          if (d.wbT != nullptr) d.wbT[static_cast<int64_t>(c) * d.dstT_ld + r] = pk_bf16(v);
        }
      }
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256) unpack_grads_kernel(const PackDesc* __restrict__ table) {
  pdl_enter();
  __shared__ float4 s_part[3][64];
  const PackDesc d = table[blockIdx.y];
  if (d.grad_dst == nullptr) return;
  // four consecutive columns per thread: the split-K partials (the bulk of the traffic) are read as float4 whenever the
  // packed side is 16-byte aligned; the parameter side (.grad, arbitrary width) is written element by element
  const int quads = (d.cols + 3) >> 2;
  const int total = d.rows * quads;
  const int64_t plane = static_cast<int64_t>(d.p_rows) * d.p_cols;
  const bool vec = d.split > 0 && (d.p_cols & 3) == 0 && (d.dc & 3) == 0 && (reinterpret_cast<uintptr_t>(d.part) & 15u) == 0;
  if (vec && d.split >= 32) {
    // many partials per element (the 148-way splits of the 160-row matrices): one thread per quad would chain
    // split / 8 dependent batches of loads.  Four thread groups take the planes z = k, k + 4, ... of the same 64 quads
    // and the first group adds the four sums in a fixed order (deterministic; another association than z = 0, 1, ...).
    const int q_in = threadIdx.x & 63, sl = threadIdx.x >> 6;
    for (int base = blockIdx.x * 64; base < total; base += gridDim.x * 64) {      // uniform over the CTA
      const int i = base + q_in;
      const bool live = i < total;
      const int r = live ? i / quads : 0, c0 = live ? (i - r * quads) * 4 : 0;
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        const float* p = d.part + static_cast<int64_t>(d.dr + r) * d.p_cols + d.dc + c0;
#pragma unroll 8
        for (int z = sl; z < d.split; z += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p + z * plane));
          sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
      }
      if (sl > 0) s_part[sl - 1][q_in] = sum;
      __syncthreads();
      if (sl == 0 && live) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 v = s_part[k][q_in];
          sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        const float* gp = d.g + static_cast<int64_t>(r) * d.dst_ld + c0;
        float* dst = d.grad_dst + static_cast<int64_t>(r) * d.src_ld + c0;
        // (a quad that sticks out of the block reads the padding of the packed row: p_cols and dc are multiples of 4)
        const float a4[4] = {sum.x, sum.y, sum.z, sum.w};
        const int nc = d.cols - c0 < 4 ? d.cols - c0 : 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < nc) dst[j] += gp[j] + a4[j];
      }
      __syncthreads();
    }
    return;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / quads, c0 = (i - r * quads) * 4;
    const int nc = d.cols - c0 < 4 ? d.cols - c0 : 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (d.split > 0) {      // deferred split-K reduction: fixed order z = 0, 1, ... (deterministic)
      const float* p = d.part + static_cast<int64_t>(d.dr + r) * d.p_cols + d.dc + c0;
      if (vec && nc == 4) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int z = 0; z < d.split; ++z) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p + z * plane));
          s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        acc[0] = s.x; acc[1] = s.y; acc[2] = s.z; acc[3] = s.w;
      } else {
        for (int j = 0; j < nc; ++j) {
          float s = 0.f;
#pragma unroll 8
          for (int z = 0; z < d.split; ++z) s += __ldg(p + z * plane + j);
          acc[j] = s;
        }
      }
    }
    const float* gp = d.g + static_cast<int64_t>(r) * d.dst_ld + c0;
    float* dst = d.grad_dst + static_cast<int64_t>(r) * d.src_ld + c0;
    for (int j = 0; j < nc; ++j) dst[j] += gp[j] + acc[j];
  }
}

extern "C" int ax2d_pack_weights(const void* table, int n_blocks, int max_block_elems, ax2d_stream_t stream) {
  using namespace ax2d;
  AX2D_CHECK_ARG(table != nullptr && n_blocks > 0 && max_block_elems > 0, "ax2d_pack_weights: bad arguments");
  int gx = (max_block_elems + 1023) / 1024;       // one CTA per 32 x 32 tile of the largest block (rough: ragged
  gx = gx > 128 ? 128 : (gx < 1 ? 1 : gx);        // blocks loop)
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(n_blocks));
  launch_k(pack_weights_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), static_cast<const PackDesc*>(table));
  return launch_status("ax2d_pack_weights");
}
extern "C" int ax2d_unpack_grads(const void* table, int n_blocks, int max_block_elems, ax2d_stream_t stream) {
  using namespace ax2d;
  AX2D_CHECK_ARG(table != nullptr && n_blocks > 0 && max_block_elems > 0, "ax2d_unpack_grads: bad arguments");
  int gx = (max_block_elems / 4 + 255) / 256;
  gx = gx > 100 ? 100 : (gx < 1 ? 1 : gx);     // ~one quad per thread for the large blocks: their partial sums are latency-bound
  dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(n_blocks));
  launch_k(unpack_grads_kernel, dim3(grid), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), static_cast<const PackDesc*>(table));
  return launch_status("ax2d_unpack_grads");
}
