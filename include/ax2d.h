/*
 * ax2d.h -- C ABI of libax2d.so, the B200 (sm_100a) hot path of AIMNet-X2D.
 *
 * The reference (mahdi-shafiei/AIMNet-X2D) is 100 % Python and has no FFI layer; its hot path bottoms
 * out in torch / torch_scatter calls.  Every entry point below replaces one of those call sites (cited
 * as file:line relative to /root/reference/src) and is what a ctypes / torch-extension binding on the
 * reference side would bind.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  Pointers are DEVICE pointers unless the function
 *     name starts with ax2d_host_.
 *   - the caller owns and allocates every buffer (outputs, scratch); scratch sizes come from the
 *     *_workspace() queries.  No hidden global state, no allocation, no synchronisation inside:
 *     everything is enqueued on `stream` and is re-entrant.
 *   - feature matrices are row-major fp32 with a leading dimension `ld` in elements.  "width" arguments
 *     are PADDED widths; pad columns must hold exact zeros (every kernel here preserves that).
 *   - return 0 on success, a negative AX2D_ERR_* code otherwise; ax2d_last_error() gives a message.
 *     Index validity is checked once at collation (ax2d_host_*), never inside the hot kernels.
 *   - no atomics anywhere on floating-point data: all reductions have a fixed order.
 */
#ifndef AX2D_H_
#define AX2D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* ax2d_stream_t;

enum { AX2D_OK = 0, AX2D_ERR_ARG = -1, AX2D_ERR_ALIGN = -2, AX2D_ERR_DTYPE = -3,
       AX2D_ERR_LAUNCH = -4, AX2D_ERR_UNSUPPORTED = -5 };
enum { AX2D_F32 = 0, AX2D_BF16 = 1 };
/* utils/activation.py:23-29 */
enum { AX2D_ACT_NONE = 0, AX2D_ACT_RELU = 1, AX2D_ACT_LEAKYRELU = 2, AX2D_ACT_ELU = 3,
       AX2D_ACT_GELU = 4, AX2D_ACT_SILU = 5 };
enum { AX2D_SEG_SUM = 0, AX2D_SEG_MEAN = 1, AX2D_SEG_MAX = 2 };

int         ax2d_abi_version(void);
const char* ax2d_error_string(int code);
const char* ax2d_last_error(void);
/* number of CUDA kernels this library has enqueued in this process so far (bench.py's gpu_launches). */
uint64_t    ax2d_launch_count(void);
/* Programmatic dependent launch of the step's kernels (off unless AX2D_PDL=1 or ax2d_set_pdl(1)): inside a captured step
 * graph the next kernel's prologue overlaps the tail of the current one; every kernel waits (griddepcontrol.wait)
 * before it touches memory, so results do not depend on the setting.  Returns the previous setting. */
int         ax2d_set_pdl(int on);

/* ------------------------------------------------------------------------------------------------
 * Collation-time integer work (HOST).  Replaces the Python loops of
 * datasets/molecular.py:427-438 (edge concatenation) and prepares what layers.py:154-163 needs.
 * ---------------------------------------------------------------------------------------------- */

/* Stable counting sort of the E edges by key -> CSR.  edges is [E,2] int64 with element strides
 * (stride_e, stride_c) (the reference hands out a transposed, non-contiguous view, molecular.py:436).
 * transpose == 0: key = edges[e,0] (target, in [0,R)),      val = edges[e,1] mod N   (layers.py:154)
 * transpose == 1: key = edges[e,1] mod N (src, R must be N), val = edges[e,0]
 * Outputs: rowptr[R+1], col[E], perm[E] (perm may be NULL).  Returns AX2D_ERR_ARG on an out-of-range index. */
int ax2d_host_csr_build(const int64_t* edges, int64_t E, int64_t stride_e, int64_t stride_c,
                        int64_t N, int64_t R, int transpose,
                        int32_t* rowptr, int32_t* col, int32_t* perm);

/* unique[0] = 1 iff no (row, column) pair occurs twice in the CSR, i.e. no (target, source) edge of
 * multi_hop_edge_indices is repeated (a dense tile product could not count a repeated edge twice). */
int ax2d_host_csr_rows_unique(const int32_t* rowptr, const int32_t* col, int64_t R, int64_t N, int32_t* unique);
/* Greedy packing of whole molecules into row tiles of at most `cap` rows (a molecule larger than cap
 * gets a tile of its own).  seg_ptr[B+1] are the molecule row offsets.  tile_ptr must hold B+1 entries.
 * On return *n_tiles tiles, *max_rows = largest tile.  If rowptr/col are given (R == N rows), *tile_local
 * is set to 1 iff every col of every row lies inside the row's own tile (true for the reference
 * collation, where no edge leaves its molecule, molecular.py:429-433). */
int ax2d_host_tile_plan(const int32_t* seg_ptr, int64_t B, int64_t cap,
                        const int32_t* rowptr, const int32_t* col,
                        int32_t* tile_ptr, int64_t* n_tiles, int64_t* max_rows, int32_t* tile_local);

/* Shell-edge BFS for a batch of molecules + edge collation in one pass (datasets/features.py:97-150 and
 * molecular.py:427-438).  atom_ptr[B+1] global atom offsets, bond_ptr[B+1] offsets into bonds[.,2]
 * (molecule-local atom indices, undirected, each bond once).  hop_counts[B*H] receives the number of
 * directed pairs per (molecule, hop).  If edges_out != NULL it receives the [E,2] int64 (target, src)
 * list with atom offsets applied, molecules then hops then BFS discovery order; capacity = rows of
 * edges_out.  Returns the total E (>= 0) or a negative error. */
int64_t ax2d_host_shell_edges(int64_t B, const int64_t* atom_ptr, const int64_t* bond_ptr,
                              const int32_t* bonds, int num_hops,
                              int64_t* hop_counts, int64_t* edges_out, int64_t capacity);

/* f-1: the same BFS emitting the CSR of the aggregation kernels DIRECTLY (features.py:82-150 + the collation of
 * molecular.py:426-438 + the stable sort of ax2d_host_csr_build in one pass, no edge list): rowptr[N+1] (shared by both
 * CSRs: shell relations are symmetric), col[E] = sources of every target row, hop-major and in BFS discovery order inside
 * a hop (the reference's stable edge order), col_t[E] = targets of every source row, hop-major and ascending.  Call with
 * col == NULL to get rowptr and E, then with col / col_t of capacity >= E.  Returns E (>= 0) or a negative error.
 * Bit-identical to ax2d_host_csr_build over the edge list of ax2d_host_shell_edges. */
int64_t ax2d_host_shell_csr(int64_t B, const int64_t* atom_ptr, const int64_t* bond_ptr, const int32_t* bonds, int num_hops,
                            int32_t* rowptr, int32_t* col, int32_t* col_t, int64_t capacity);

/* ... and on the device (one warp per molecule, adjacency bit matrices in shared memory, BFS on register bitsets; molecules
 * of at most 256 atoms).  All pointers are device pointers: atom_ptr / bond_ptr [B+1] int32, bonds [nb,2] int32 (molecule-
 * local indices), err one int32 zeroed by the caller (non-zero afterwards: 1 molecule larger than max_atoms, 2 bond index
 * out of range, 3 more than 2^31 edges).  _count writes rowptr[N+1] (workspace: N int32); read E = rowptr[N], then _fill
 * writes col / col_t [E].  Bit-identical to ax2d_host_shell_csr. */
int ax2d_shell_csr_count(const int32_t* atom_ptr, const int32_t* bond_ptr, const int32_t* bonds, int64_t B, int64_t N,
                         int max_atoms, int num_hops, int32_t* rowptr, int32_t* workspace, int32_t* err, ax2d_stream_t stream);
int ax2d_shell_csr_fill(const int32_t* atom_ptr, const int32_t* bond_ptr, const int32_t* bonds, int64_t B, int max_atoms,
                        int num_hops, const int32_t* rowptr, int32_t* col, int32_t* col_t, int32_t* err, ax2d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a1  ShellConvolutionLayer.message_passing  (models/layers.py:133-167): gather rows + scatter_add.
 *   out[r,:] = (addend ? addend[r,:] : 0) + sum_{k in [rowptr[r], rowptr[r+1])} x[col[k],:]
 * summed sequentially in CSR order (bit-identical to the reference CPU scatter_add).  Backward is the
 * same call with the transposed CSR.  width % 4 == 0.
 * Tiled mode (tile_info != NULL): requires n_out_rows == n_src_rows, ldx == width, width % 32 == 0 and tile-local
 * columns (ax2d_host_tile_plan).  tile_info is [n_tiles][4] int32 = {row0, row1, rowptr[row0], rowptr[row1]} per tile
 * (16-byte aligned); max_tile_rows / max_tile_edges bound row1 - row0 and the edge count of a tile.  A persistent
 * kernel stages each tile's rows of x AND its rowptr / col windows in shared memory with bulk async copies issued by
 * a producer warp several tiles ahead and gathers from there; rowptr and col must be readable up to 3 entries past
 * their logical end (16-byte windows).  Otherwise a global-gather kernel is used.
 * ---------------------------------------------------------------------------------------------- */
int ax2d_agg(const void* x, int64_t ldx, int64_t n_src_rows,
             void* out, int64_t ldo, int64_t n_out_rows,
             const int32_t* rowptr, const int32_t* col,
             const void* addend, int64_t ld_addend, int width,
             const int32_t* tile_info, int64_t n_tiles, int max_tile_rows, int max_tile_edges,
             int dtype, ax2d_stream_t stream);

/* The same operation for whole-molecule tiles as a dense product on the tensor cores: out_tile = A_tile x_tile with A the
 * tile's 0/1 adjacency (models/layers.py:154-163 restricted to the shipped collation, where no edge leaves its molecule).
 * Contract of the tiled mode of ax2d_agg PLUS: the edge list holds no duplicate (target, source) pair (checked at collation).
 * bf16 features: fp32 accumulation, one rounding per output.  fp32 features: x is split into three bf16 terms (exact), so
 * the result differs from the sequential sum only by the order of the fp32 additions (~1e-7 relative); ax2d_agg remains
 * the bit-exact path.  ..._supported: 1 iff the tile shape fits (width % 16 == 0, <= 128 rows per tile, shared memory). */
int ax2d_agg_tiles_mma_supported(int width, int max_tile_rows, int max_tile_edges, int dtype);
int ax2d_agg_tiles_mma(const void* x, int64_t ldx, int64_t n_rows, void* out, int64_t ldo,
                       const int32_t* rowptr, const int32_t* col, const void* addend, int64_t ld_addend, int width,
                       const int32_t* tile_info, int64_t n_tiles, int max_tile_rows, int max_tile_edges,
                       int dtype, ax2d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a7  MultiHeadAttentionPoolingLayer.forward  (models/pooling.py:122-172)
 *   z[h,n] = (w_h . x_n + b_h) / T ; a = per-(head,molecule) softmax ; pooled[g] = mean_h sum_n a[h,n] x_n
 * x [N,F] (ld), seg_ptr[B+1], w [heads,F], b [heads], temperature: device scalar.
 * Outputs pooled [B,F], attn [heads,N], z [heads,N] (saved scores for backward).
 * x must be contiguous (ldx == F).  max_rows_hint: largest molecule of the batch (known at collation;
 * 0 = unknown) -- sizes the shared-memory chunk of the backward and of the two-phase forward kernels; larger
 * molecules are still handled (chunked).  The default forward (F <= 512, heads x F <= 2048) streams the rows
 * once and does not use it.
 * ---------------------------------------------------------------------------------------------- */
int ax2d_attn_pool_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int64_t B, int64_t N,
                       int F, int heads, const float* w, const float* b, const float* temperature,
                       float* pooled, float* attn, float* z, int max_rows_hint, ax2d_stream_t stream);
/* Development / test switch of the forward: mode 0 (default) = single-pass streaming kernel (online softmax, x read once)
 * whenever F <= 512 and heads x F <= 2048, 1 = the two-phase kernels only, 2 = the streaming kernel with the head weights
 * in registers instead of shared memory (measured slower); group = warps per molecule of the streaming kernel
 * (0 = chosen from the batch, or 1 / 2 / 4). */
int ax2d_attn_pool_fwd_config(int mode, int group);
int64_t ax2d_attn_pool_bwd_workspace(int64_t B, int F, int heads);
/* g_attn may be NULL (no gradient flows into the returned attention weights).  gx [N,F] (ld gx).
 * gw [heads,F], gb [heads], gT [1] are OVERWRITTEN.  workspace: bytes from the query above. */
int ax2d_attn_pool_bwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int64_t B, int64_t N,
                       int F, int heads, const float* w, const float* temperature,
                       const float* attn, const float* z, const float* g_pooled, const float* g_attn,
                       float* gx, int64_t ldgx, float* gw, float* gb, float* gT,
                       void* workspace, int max_rows_hint, ax2d_stream_t stream);

/* a8  Mean/Max/SumPoolingLayer (models/pooling.py:15-80, torch_scatter semantics: empty segment -> 0,
 * max keeps the FIRST maximal row and back-propagates to it only).  arg [B,F] int32 (max only). */
int ax2d_seg_reduce_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int64_t B, int F, int mode,
                        float* out, int32_t* arg, ax2d_stream_t stream);
int ax2d_seg_reduce_bwd(const float* g_out, const int32_t* seg_ptr, int64_t B, int F, int mode,
                        const int32_t* arg, float* gx, int64_t ldgx, ax2d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a4  GNN._partial_charge_calculation  (models/gnn.py:622-658).  Columns 0 (q) and 1 (f) of x change,
 * the rest is copied.  stats [B,2] receives (dQ_g, Fu_g) for the backward.
 * ---------------------------------------------------------------------------------------------- */
int ax2d_charge_eq_fwd(const float* x, int64_t ldx, const int32_t* seg_ptr, int64_t B, int width,
                       const float* total_charges, float* out, int64_t ldo, float* stats,
                       ax2d_stream_t stream);
int ax2d_charge_eq_bwd(const float* x, int64_t ldx, const float* out, int64_t ldo, const int32_t* seg_ptr,
                       int64_t B, int width, const float* stats, const float* g_out, int64_t ldg,
                       float* gx, int64_t ldgx, ax2d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a5  stereo features (models/gnn.py:387-509).
 * tetra: idx [M,4] int32 neighbour rows of each centre; slot CSR (slot_ptr[N+1], slot_idx[4M]) lists,
 * per atom, the (m*4+i) slots that reference it in ascending order (built at collation; replaces
 * torch.unique + index_add_).  contrib [4M,width] scratch.  out = mask * (x + sum contrib) (quirk Q4).
 * cis/trans (quirk Q3): out = x; out[tgt[j]] += sign[j] * x[src[j]] for j < n_upd (<= 4), sequential.
 * ---------------------------------------------------------------------------------------------- */
int ax2d_tetra_fwd(const float* x, int64_t ldx, int64_t N, int width, int true_width,
                   const int32_t* idx, int64_t M, const int32_t* slot_ptr, const int32_t* slot_idx,
                   float* contrib, float* out, int64_t ldo, ax2d_stream_t stream);
int ax2d_tetra_bwd(const float* x, int64_t ldx, int64_t N, int width, int true_width,
                   const int32_t* idx, int64_t M, const int32_t* slot_ptr, const int32_t* slot_idx,
                   const float* g_out, int64_t ldg, float* g_rows, float* gx, int64_t ldgx,
                   ax2d_stream_t stream);
int ax2d_cistrans(const float* x, int64_t ldx, int64_t N, int width,
                  const int32_t* upd_src, const int32_t* upd_tgt, const float* upd_sign, int n_upd,
                  int transpose, float* out, int64_t ldo, ax2d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a6  embedding lookups + concat (models/gnn.py:262-274): out[n, t*E:(t+1)*E] = table_t[idx_t[n], :].
 * Backward: per-table segment sum over atoms sorted (stably) by index value -- order[t] / ptr[t]
 * built at collation -- so it matches the CPU index_add order without atomics.
 * ---------------------------------------------------------------------------------------------- */
int ax2d_embed_fwd(const float* const* tables, const int64_t* const* indices, int n_tables, int emb_dim,
                   int64_t N, float* out, int64_t ldo, ax2d_stream_t stream);
/* Backward of ALL lookups in one pass (tables small enough for shared memory: sum(vocab) * emb_dim * 4 <= 200 KB):
 * g_tables[t][v, :] = sum over atoms n with indices[t][n] == v of g_out[n, t * emb_dim : (t + 1) * emb_dim], summed in a
 * fixed order (atom order inside a CTA's slice, CTA order across slices).  Replaces nn.Embedding's backward
 * (models/gnn.py:262-274 under autograd). */
int64_t ax2d_embed_bwd_all_workspace(int64_t N, int64_t total_rows, int emb_dim);
int ax2d_embed_bwd_all(const float* g_out, int64_t ldg, int n_tables, int emb_dim, int64_t N, const int64_t* const* indices,
                       const int64_t* vocab, float* const* g_tables, void* workspace, ax2d_stream_t stream);
int64_t ax2d_embed_bwd_workspace(int64_t vocab, int emb_dim);
int ax2d_embed_bwd(const float* g_out, int64_t ldg, int table, int emb_dim, int64_t vocab,
                   const int32_t* order, const int32_t* ptr, float* g_table, void* workspace,
                   ax2d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a2 / a5 / a6  dense projections (every nn.Linear on the path: layers.py:47-61, gnn.py:96-146,190-195)
 *
 * A logical operand is a matrix made of up to AX2D_MAX_SEG column segments (so `cat` is never
 * materialised: layers.py:79, gnn.py:245,257,323).  C = epilogue(op(A) * op(B)):
 *   trans_a == 0: A[m,k] = a.seg(k)[m, k']         (reduce over A's columns)
 *   trans_a == 1: A[m,k] = a.seg(m)[k, m']         (reduce over A's rows   -- weight gradients)
 *   trans_b == 0: B[k,n] = b.seg(n)[k, n']         (reduce over B's rows)
 *   trans_b == 1: B[k,n] = b.seg(k)[n, k']         (reduce over B's columns -- y = x W^T, torch layout)
 * Epilogue, in order:  v = acc (+ bias[n]);  pre[m,n] = v;  v = act(v) for n < act_cols;
 *   v *= mask[m,n] (explicit mask) or hash-dropout(seed, p);  v += sum_r resid_r[m,n];
 *   v *= act'(dact_pre[m,n]) * dropout(m,n) (backward of the two steps above);  c[m,n] (+)= v.
 * Segment widths and K must be multiples of 4; all pointers 16-byte aligned.
 * split_k > 1 (weight gradients): partial products go to `workspace` and are reduced in a fixed order.
 * ---------------------------------------------------------------------------------------------- */
#define AX2D_MAX_SEG 8
typedef struct { const float* ptr[AX2D_MAX_SEG]; int64_t ld[AX2D_MAX_SEG]; int32_t width[AX2D_MAX_SEG]; int32_t n_seg; } ax2d_cmat;
typedef struct { float* ptr[AX2D_MAX_SEG]; int64_t ld[AX2D_MAX_SEG]; int32_t width[AX2D_MAX_SEG]; int32_t n_seg; } ax2d_mat;
typedef struct {
  const float* bias;            /* [n] or NULL */
  ax2d_mat     pre;             /* optional pre-activation copy (n_seg == 0 -> none) */
  int32_t      act;             /* AX2D_ACT_* applied to columns < act_cols */
  int32_t      act_cols;
  const float* mask;  int64_t ld_mask;      /* explicit dropout mask (already scaled) or NULL */
  float        drop_p; uint64_t drop_seed;  /* hash dropout if drop_p > 0 and mask == NULL */
  const uint64_t* drop_tick;    /* optional DEVICE counter added to drop_seed (CUDA-graph friendly; ax2d_tick) */
  ax2d_cmat    resid;           /* n_seg residual matrices; matrix r is added to columns [0, width[r]) (0 = all) */
  const float* dact_pre; int64_t ld_dact; int32_t dact; /* columns < dact_cols: multiply by act'(dact_pre) (and the dropout above) */
  int32_t      dact_cols;
  int32_t      accumulate;      /* c += v instead of c = v */
} ax2d_epilogue;

int64_t ax2d_gemm_workspace(int64_t M, int64_t N, int64_t K, int trans_a, int split_k);
int ax2d_gemm(const ax2d_cmat* a, int trans_a, const ax2d_cmat* b, int trans_b,
              const ax2d_mat* c, int64_t M, int64_t N, int64_t K,
              const ax2d_epilogue* ep, int split_k, void* workspace, ax2d_stream_t stream);
/* Tensor-core path (tcgen05 / TMEM / TMA) for the same contraction, fp32-faithful by a 3xTF32 operand split:
 *   C[M,N] = epilogue( A[M,K] * B[N,K]^T ),  A column-segmented activations, B = weights given as two fp32 arrays
 *   b_hi + b_lo == B produced by ax2d_split_tf32 (transpose = 1 turns a [K,N] row-major operand -- the weight of a
 *   data-gradient product dX = dY W -- into the [N,K] K-major form).  Same epilogue semantics as ax2d_gemm.
 * Supported when every A segment width and K are multiples of 16 and N % 4 == 0 (ax2d_gemm_tc_supported). */
int ax2d_split_tf32(const float* w, int64_t ldw, int rows, int cols, int transpose, float* hi, float* lo, int64_t ldo,
                    ax2d_stream_t stream);
int ax2d_gemm_tc_supported(const ax2d_cmat* a, int64_t M, int64_t N, int64_t K);
int ax2d_gemm_tc(const ax2d_cmat* a, const float* b_hi, const float* b_lo, int64_t ldb, const ax2d_mat* c,
                 int64_t M, int64_t N, int64_t K, const ax2d_epilogue* ep, ax2d_stream_t stream);
/* Weight gradients on the tensor cores: C[M,N] (+)= A^T B, A = a [K rows, M columns], B = b [K rows, N columns], both
 * column-segmented activations (dW = dY^T X of every nn.Linear; contraction over the K atom / molecule rows, split
 * over CTAs and reduced in a fixed order).  Segment widths must be multiples of 32, M, N >= 32.
 * bias_grad (optional, [M]): receives the column sums of A (db = sum_rows dY) computed from the tiles the kernel
 * already stages -- replaces a separate ax2d_colsum pass. */
int     ax2d_gemm_tc_wgrad_supported(const ax2d_cmat* a, const ax2d_cmat* b, int64_t M, int64_t N, int64_t K);
int64_t ax2d_gemm_tc_wgrad_workspace(int64_t M, int64_t N, int64_t K);
/* Number of row splits the kernel uses for this shape.  With accumulate == 2 (and more than one split) the call leaves
 * the partial tiles [splits][M][N] followed by the partial bias vectors [splits][M] in `workspace` and does not reduce
 * them: ax2d_unpack_grads sums them, in split order, while it scatters the gradients (c and bias_grad are not written). */
int     ax2d_gemm_tc_wgrad_splits(int64_t M, int64_t N, int64_t K);
int     ax2d_gemm_tc_wgrad(const ax2d_cmat* a, const ax2d_cmat* b, const ax2d_mat* c, int64_t M, int64_t N, int64_t K,
                           int accumulate, float* bias_grad, void* workspace, ax2d_stream_t stream);
/* ------------------------------------------------------------------------------------------------
 * bf16 configuration (BASELINE configs[3]; reference mixed precision: training/trainer.py:133-149, cli.py:236).
 * Activations and the packed weight copies are bf16, accumulation is fp32 (TMEM), master weights / weight gradients /
 * optimiser state stay fp32.  In these entry points the `float*` members of ax2d_cmat / ax2d_mat / ax2d_epilogue.pre /
 * .resid / .dact_pre point to BF16 data (leading dimensions in bf16 elements, multiples of 8); `bias` stays fp32.
 *   ax2d_gemm_bf16:  C[M,N] = epilogue(A B^T), A column-segmented, B = bf16 [N,K] (K-major; ldb), C = c_dtype (AX2D_BF16 or
 *     AX2D_F32) in up to 4 column segments.  Epilogue as ax2d_gemm (bias, pre copy, act, hash dropout, <= 3 residuals,
 *     act' * dropout); explicit masks and `accumulate` are not supported.  act_cols / dact_cols / residual widths / pre
 *     segment boundaries must coincide with output-segment boundaries or multiples of the column tile.
 *   ax2d_gemm_bf16_wgrad:  C[M,N] (fp32) (+)= A^T B over the K rows, segment widths % 32 == 0.  The workspace
 *     (ax2d_gemm_bf16_wgrad_workspace) is always required; accumulate: 0 write, 1 add, 2 leave the fp32 partials
 *     [splits][M][N] + [splits][M] (bias) in the workspace for ax2d_unpack_grads.  bias_grad: optional [M] column sums of A.
 * ---------------------------------------------------------------------------------------------- */
int     ax2d_gemm_bf16_supported(const ax2d_cmat* a, int64_t M, int64_t N, int64_t K);
int     ax2d_gemm_bf16(const ax2d_cmat* a, const void* b, int64_t ldb, const ax2d_mat* c, int c_dtype,
                       int64_t M, int64_t N, int64_t K, const ax2d_epilogue* ep, ax2d_stream_t stream);
int     ax2d_gemm_bf16_wgrad_supported(const ax2d_cmat* a, const ax2d_cmat* b, int64_t M, int64_t N, int64_t K);
int     ax2d_gemm_bf16_wgrad_splits(int64_t M, int64_t N, int64_t K);
int64_t ax2d_gemm_bf16_wgrad_workspace(int64_t M, int64_t N, int64_t K);
int     ax2d_gemm_bf16_wgrad(const ax2d_cmat* a, const ax2d_cmat* b, const ax2d_mat* c, int64_t M, int64_t N, int64_t K,
                             int accumulate, float* bias_grad, void* workspace, ax2d_stream_t stream);
/* bf16 variants of the embedding kernels: out / g_out are bf16 (emb_dim, ld multiples of 8); tables and their gradients fp32. */
int     ax2d_embed_fwd_bf16(const float* const* tables, const int64_t* const* indices, int n_tables, int emb_dim,
                            int64_t N, void* out, int64_t ldo, ax2d_stream_t stream);
int     ax2d_embed_bwd_all_bf16(const void* g_out, int64_t ldg, int n_tables, int emb_dim, int64_t N,
                                const int64_t* const* indices, const int64_t* vocab, float* const* g_tables, void* workspace,
                                ax2d_stream_t stream);
/* fp32 <-> bf16 copy of an [M, width] matrix (width, ldi, ldo multiples of 8). */
int     ax2d_convert(const void* in, int64_t ldi, int in_dtype, void* out, int64_t ldo, int out_dtype, int64_t M, int width,
                     ax2d_stream_t stream);
/* bf16 variant of ax2d_act_bwd (width % 8 == 0). */
int     ax2d_act_bwd_bf16(const void* g, int64_t ldg, const void* pre, int64_t ldp, void* out, int64_t ldo, int64_t M,
                          int width, int act, ax2d_stream_t stream);

/* elementwise g_pre[m,n] = g[m,n] * act'(pre[m,n]) for n < width (width % 4 == 0). */
int ax2d_act_bwd(const float* g, int64_t ldg, const float* pre, int64_t ldp, float* out, int64_t ldo,
                 int64_t M, int width, int act, ax2d_stream_t stream);
/* counter[0] += 1 on the device (dropout tick). */
int ax2d_tick(uint64_t* counter, ax2d_stream_t stream);
/* column sums (bias gradients): out[n] (+)= sum_m a[m,n], fixed order, two passes. */
int64_t ax2d_colsum_workspace(int64_t M, int64_t N);
int ax2d_colsum(const ax2d_cmat* a, int64_t M, int64_t N, float* out, int accumulate,
                void* workspace, ax2d_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a10 / a11  after the gradient all-reduce (training/trainer.py:164-165, runner.py:703-707):
 * one flat fp32 arena of n floats.  sqnorm: partials[blocks] then norm2[0] = sum (fixed order).
 * clip_adam: g *= grad_scale (1/world); coef = min(1, max_norm/(sqrt(norm2)*grad_scale + 1e-6));
 * Adam(lr, b1, b2, eps) with bias corrections computed from the device step counter step[0]
 * (1-based, already incremented), no host synchronisation.
 * ---------------------------------------------------------------------------------------------- */
int64_t ax2d_sqnorm_workspace(int64_t n);
/* norm2[0] = sum g^2.  If step_inc != NULL the device step counter step_inc[0] is incremented afterwards
 * (so a CUDA-graph-captured training step needs no host-side counter). */
int ax2d_sqnorm(const float* g, int64_t n, float* norm2, int64_t* step_inc, void* workspace,
                ax2d_stream_t stream);
int ax2d_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* norm2,
                   float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps,
                   const int64_t* step, ax2d_stream_t stream);

/* The same update with the hyper-parameters in DEVICE memory: hyper[6] = {grad_scale, max_norm, lr, beta1, beta2, eps}.
 * A captured CUDA graph of the step then follows the reference's learning-rate schedulers
 * (training/trainer.py:60-93, 268-271: ReduceLROnPlateau & co. rewrite param_group['lr'] between steps). */
int ax2d_clip_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* norm2,
                       const float* hyper, const int64_t* step, ax2d_stream_t stream);

/* Packed weights: ONE launch gathers every projection weight / bias from its reference-shaped parameter into the
 * padded, packed layouts the kernels read (plus the two TF32 terms, plain and transposed), and one launch adds the
 * packed weight gradients back into the parameters' gradients.  `table` is a DEVICE array of n_blocks 128-byte block
 * descriptors (layout: struct PackDesc in csrc/embed_optim.cu, built by aimnet_x2d_b200/packed.py):
 *   { const float* src; float* grad_dst; float* w; float* hi; float* lo; float* hiT; float* loT; const float* g;
 *     const float* part; int32 rows, cols, src_ld, dst_ld, dstT_ld, split, p_rows, p_cols, dr, dc;
 *     bf16* wb; bf16* wbT; }   (wb / wbT: bf16 copies for the bf16 configuration, may be null)
 * Replaces the per-call torch pad / cat / slice / accumulate kernels around every nn.Linear (layers.py:47-61,
 * gnn.py:96-146). */
int ax2d_pack_weights(const void* table, int n_blocks, int max_block_elems, ax2d_stream_t stream);
int ax2d_unpack_grads(const void* table, int n_blocks, int max_block_elems, ax2d_stream_t stream);

/* a12  WeightedL1Loss / WeightedMSELoss (models/losses.py:14-87): loss[0] = mean_b sum_t w_t |p-y|
 * (kind 0) or w_t (p-y)^2 (kind 1); g_pred = d loss / d pred * upstream (upstream = 1 if NULL). */
int ax2d_weighted_loss(const float* pred, const float* target, const float* weights, int64_t B, int T,
                       int kind, float* loss, float* g_pred, ax2d_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AX2D_H_ */
