/*
 * b200gcn — C ABI of the Blackwell-native bipartite graph-convolution engine.
 *
 * The reference (RUCAIBox/RecBole-GNN @ 632ef888) is pure Python and has no FFI of its own; its
 * arithmetic for this path lives in third-party wheels (PyG MessagePassing.propagate,
 * torch_sparse.matmul, PyG gcn_norm).  The entry points below are what a binding for THIS path has
 * to provide; each one names the reference interface it stands in for (paths relative to the
 * reference checkout).  INTEGRATION.md shows the ctypes stub a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless its name starts with h_;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call is
 *     asynchronous on that stream and performs no host synchronisation unless stated;
 *   - float rows must be 16-byte aligned: base pointers 16-byte aligned, leading dimensions (`ld*`,
 *     in floats) multiples of 4, `dim` a multiple of 4 and <= 512;
 *   - indices: reference COO is int64 (`edge_index`); the engine's CSR is int64 rowptr / int32 col;
 *   - return value 0 = success; otherwise a b200gcn_status and b200gcn_last_error() (thread-local,
 *     valid until the next call on the thread) describes it.  Inputs are never modified.
 *   - there is no CPU implementation behind any of these symbols.
 */
#ifndef B200GCN_H_
#define B200GCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GCN_ABI_VERSION 10

typedef enum b200gcn_status {
  B200GCN_OK = 0,
  B200GCN_ERR_INVALID = 1,   /* bad argument (shape, alignment, NULL) -> ValueError on the Python side */
  B200GCN_ERR_CUDA = 2,      /* a CUDA runtime call or kernel launch failed */
  B200GCN_ERR_WORKSPACE = 3, /* workspace too small */
  B200GCN_ERR_RANGE = 4      /* an index is outside [0, n) (only raised by the checking entry points) */
} b200gcn_status;

int b200gcn_abi_version(void);
const char* b200gcn_last_error(void);

/* sm count, L2 bytes, and total HBM bytes of the current device (used to size grids and column tiles). */
int b200gcn_device_info(int32_t* sm_count, int64_t* l2_bytes, int64_t* hbm_bytes, int32_t* cc_major,
                        int32_t* cc_minor);

/* ---------------------------------------------------------------------------------------------
 * Graph build.
 *
 * b200gcn_csr_from_coo: COO (int64 source ids `src` = edge_index[0], destination ids `dst` =
 * edge_index[1], optional fp32 weights) -> CSR keyed by destination, entries of a row ordered by
 * (source id, original position), duplicates kept.  Stands in for
 * GeneralGraphDataset.edge_index_to_adj_t = SparseTensor(row, col, value, sizes).t()
 * (recbole_gnn/data/dataset.py:41-47) and for the implicit COO->CSR conversion inside
 * torch_sparse.matmul (recbole_gnn/model/layers.py:19-20).
 *   rowptr [n_dst+1] int64, col [nnz] int32, val [nnz] f32 (may be NULL iff w is NULL: unit weights),
 *   perm [nnz] int64 (optional, may be NULL): COO position of every CSR entry.
 * Workspace: query the size with b200gcn_csr_from_coo_workspace, pass a device buffer of that size.
 * If `check` != 0 the ids are validated on the device first; this costs one host synchronisation
 * and returns B200GCN_ERR_RANGE for ids outside [0, n_src) x [0, n_dst).
 */
int b200gcn_csr_from_coo_workspace(int64_t nnz, int64_t n_dst, int64_t n_src, size_t* bytes);
int b200gcn_csr_from_coo(const int64_t* src, const int64_t* dst, const float* w, int64_t nnz,
                         int64_t n_dst, int64_t n_src, int64_t* rowptr, int32_t* col, float* val,
                         int64_t* perm, void* workspace, size_t workspace_bytes, int check,
                         void* stream);

/* Bipartite interaction COO -> symmetric (U+I)^2 CSR without materialising the int64 COO:
 * rows/cols of [[0,R],[R^T,0]] from the inter_feat columns uid[E], iid[E] (int64, ids incl. [PAD]=0).
 * Stands in for the COO assembly of GeneralGraphDataset.get_norm_adj_mat
 * (recbole_gnn/data/dataset.py:60-66) followed by edge_index_to_adj_t (dataset.py:73).
 * nnz = 2E.  Same outputs/workspace protocol as b200gcn_csr_from_coo (val = unit weights, not written
 * when NULL). */
int b200gcn_csr_from_interactions_workspace(int64_t n_inter, int64_t user_num, int64_t item_num,
                                            size_t* bytes);
int b200gcn_csr_from_interactions(const int64_t* uid, const int64_t* iid, int64_t n_inter,
                                  int64_t user_num, int64_t item_num, int64_t* rowptr, int32_t* col,
                                  void* workspace, size_t workspace_bytes, int check, void* stream);

/* gcn_norm(add_self_loops=False) on a square CSR (PyG; call sites recbole_gnn/data/dataset.py:74,77,
 * sgl.py:121,124): deg[r] = sum of row r (val_in, or the entry count when val_in is NULL),
 * dis = deg^-1/2 with inf -> 0, val_out[e] = (dis[col[e]] * val_in[e]) * dis[row(e)].
 * val_in may alias val_out.  dis_out [n] is optional (NULL to skip). */
int b200gcn_gcn_norm_csr(const int64_t* rowptr, const int32_t* col, const float* val_in,
                         float* val_out, float* dis_out, int64_t n, void* stream);

/* Edge weights of GeneralGraphDataset.get_bipartite_inter_mat (recbole_gnn/data/dataset.py:81-106)
 * for COO ids row_ids[E] in [0,n_row), col_ids[E] in [0,n_col):
 *   row_norm != 0: w = 1 / max(deg_row,1)[row]        (dataset.py:93-96)
 *   row_norm == 0: w = deg_row^-1/2[row] * deg_col^-1/2[col] with zero degrees read as 1 (dataset.py:97-104)
 * workspace: (n_row + n_col) * 4 bytes. */
int b200gcn_bipartite_norm_coo(const int64_t* row_ids, const int64_t* col_ids, int64_t nnz,
                               int64_t n_row, int64_t n_col, int row_norm, float* w_out,
                               void* workspace, size_t workspace_bytes, void* stream);

/* Transpose of a CSR (needed for the backward product of non-symmetric graphs: dL/dx = A^T dL/dy).
 * workspace size from b200gcn_csr_transpose_workspace. val/val_t may both be NULL (unit weights). */
int b200gcn_csr_transpose_workspace(int64_t nnz, int64_t n_rows, int64_t n_cols, size_t* bytes);
int b200gcn_csr_transpose(const int64_t* rowptr, const int32_t* col, const float* val, int64_t n_rows,
                          int64_t n_cols, int64_t nnz, int64_t* rowptr_t, int32_t* col_t, float* val_t,
                          void* workspace, size_t workspace_bytes, void* stream);

/* CSR -> COO row ids (for GraphHandle.coo(), used by NGCF's node-dropout branch ngcf.py:79). */
int b200gcn_csr_row_ids(const int64_t* rowptr, int64_t n_rows, int64_t nnz, int64_t* row_ids,
                        void* stream);

/* Edge dropout on an existing CSR without re-sorting (PyG dropout_adj semantics: keep edge e iff
 * keep[e] != 0, NO rescale; call sites ngcf.py:81,89).  Compacts col/val and rebuilds rowptr.
 * Returns the kept count through h_nnz_out (one host synchronisation). workspace from the query. */
int b200gcn_csr_mask_workspace(int64_t nnz, int64_t n_rows, size_t* bytes);
int b200gcn_csr_mask(const int64_t* rowptr, const int32_t* col, const float* val,
                     const uint8_t* keep, int64_t n_rows, int64_t nnz, int64_t* rowptr_out,
                     int32_t* col_out, float* val_out, int64_t* h_nnz_out, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Propagation: y = A x, the body of LightGCNConv.forward / BipartiteGCNConv.forward /
 * the propagate() step of BiGNNConv.forward (recbole_gnn/model/layers.py:13-20, 31-35, 55),
 * with optional fused epilogues for the model loops that call it.
 *
 *   p[r, :] = sum_{e in row r} val[e] * X[col[e], :]          (rowptr == NULL: identity mode, p[r, :] = X[r, :] —
 *     the epilogues below applied to rows that are already known: perturbed SimGCL views of a shared first
 *     layer, publishing a rank's rows to its peers)
 *     X = x for source ids < x_split, x2 (indexed by id - x_split) otherwise; x2 == NULL -> single
 *     table.  The two-table form reads user_embedding.weight / item_embedding.weight in place of
 *     LightGCN.get_ego_embeddings' torch.cat (lightgcn.py:60-68).
 *   SimGCL perturbation (simgcl.py:30-32), when eps != 0:
 *     p += sign(p) * noise[r,:] / max(||noise[r,:]||_2, 1e-12) * eps
 *     noise = the caller's rand_like draw, or, when noise == NULL, U[0,1) drawn in-kernel from
 *     Philox4x32-10 keyed by (seed, row, column block).
 *   y[r, :] = p                                   (skipped when y == NULL)
 *   layer combine (lightgcn.py:77-78, simgcl.py:34-35), when acc_out != NULL:
 *     acc_out[r, :] = ((acc_in ? acc_in[r, :] : 0) + p) * acc_scale     (acc_in may alias acc_out)
 *     acc_in may also be given as two tables like x (acc_in2 / same split) for the first layer.
 */
typedef struct b200gcn_spmm_args {
  int64_t n_rows;        /* destination rows computed by this call */
  int32_t dim;           /* embedding dimension D */
  int32_t flags;         /* 0 = engine defaults; otherwise a tuning word (kernel variant, gathers in flight,
                            prefetch distance, rows per warp — see `dispatch` in csrc/spmm.cu); never changes results
                            beyond fp32 summation order */
  const int64_t* rowptr; /* [n_rows + 1] */
  const int32_t* col;    /* [nnz] source ids */
  const float* val;      /* [nnz] or NULL (unit weights) */
  const float* x;        /* source table */
  const float* x2;       /* optional second source table (ids >= x_split) */
  int64_t x_split;
  int64_t ldx;           /* leading dimension of x and x2, floats */
  float* y;              /* [n_rows, ldy] or NULL */
  int64_t ldy;
  const float* noise;    /* [n_rows, ldn] or NULL */
  int64_t ldn;
  float eps;             /* 0 = no perturbation */
  float acc_scale;
  uint64_t seed;         /* Philox key when eps != 0 and noise == NULL */
  const float* acc_in;   /* optional */
  const float* acc_in2;  /* optional second table for rows >= acc_split */
  int64_t acc_split;
  int64_t ld_acc_in;
  float* acc_out;        /* optional */
  int64_t ld_acc_out;
  /* Row-sharded multi-GPU exchange fused into the epilogue (one process per GPU, NVLink peer memory):
   * every finished row p is ALSO stored at row (y_peer_row0 + r) of each of the n_peers buffers in
   * y_peers (device array of n_peers device pointers: the next-layer gather table on every rank,
   * peer-mapped over NVLink, this rank's own buffer included), or, when y_mc != NULL, once through the
   * NVSwitch multicast address y_mc (multimem.st; the switch replicates to all ranks).  Leading dimension
   * ld_peer.  The caller orders layers with a cross-GPU barrier after the launch.  n_peers == 0 and
   * y_mc == NULL: single-GPU behaviour. */
  float* const* y_peers;
  float* y_mc;
  int64_t y_peer_row0;
  int64_t ld_peer;
  int32_t n_peers;
  int32_t n_acc_extra;   /* 0..3 further addends of the layer combine, see acc_extra */
  /* acc_out[r] = (acc_in[r] + acc_extra[0][r] + ... + p[r]) * acc_scale, added in that order: with the earlier
   * layers' outputs passed here the LAST layer alone forms mean(x_0 .. x_L) and the earlier layers write only y
   * (1 GB fewer writes per 3-layer step at config 2 than a running sum).  All extras share ld_acc_extra. */
  const float* acc_extra[3];
  int64_t ld_acc_extra;
  /* Halo-only exchange (graphs whose partition has locality): bit q of peer_need[r] says whether rank q gathers
   * destination row r of this launch at all; rows are stored only to the peers that need them (y_peers path; the
   * multicast path always reaches every rank).  NULL = every peer needs every row. */
  const uint32_t* peer_need;
} b200gcn_spmm_args;

int b200gcn_spmm(const b200gcn_spmm_args* args, void* stream);

/* Hub-row plan for heavily skewed graphs.  b200gcn_plan_hubs lists the rows holding more than
 * `long_row` entries (at most `cap` of them are stored in hub_rows; the true count comes back through
 * h_count; one host synchronisation).  b200gcn_spmm_planned then gives each listed row a whole CTA
 * whose partial sums are combined in a fixed order (deterministic), and all other rows the row kernel.
 * b200gcn_spmm == b200gcn_spmm_planned with an empty plan. */
int b200gcn_plan_hubs(const int64_t* rowptr, int64_t n_rows, int64_t long_row, int64_t* hub_rows,
                      int32_t cap, int32_t* h_count, void* stream);
int b200gcn_spmm_planned(const b200gcn_spmm_args* args, int64_t long_row, const int64_t* hub_rows,
                         int32_t n_hubs, void* stream);

/* Chunked hub rows: for graphs whose hub rows hold millions of entries (Zipf item popularity) one CTA per row
 * is not enough.  The caller cuts every listed hub row into chunks of its choosing (the Python host uses 8192
 * entries): chunk c covers entries [chunk_beg[c], chunk_end[c]) of row hub_rows[h] for
 * hub_chunk_ptr[h] <= c < hub_chunk_ptr[h+1].  One CTA per chunk writes a partial row to scratch [n_chunks, dim];
 * a second kernel adds the partials of each hub in chunk order (deterministic) and runs the same epilogues as
 * b200gcn_spmm.  Use after b200gcn_spmm_planned(args, long_row, NULL, 0, stream), which skips the rows above
 * long_row. */
typedef struct b200gcn_hub_plan {
  int32_t n_hubs;
  int32_t n_chunks;
  const int64_t* hub_rows;      /* [n_hubs] */
  const int32_t* hub_chunk_ptr; /* [n_hubs + 1] */
  const int64_t* chunk_beg;     /* [n_chunks] absolute entry positions */
  const int64_t* chunk_end;     /* [n_chunks] */
  float* scratch;               /* [n_chunks, dim] workspace */
} b200gcn_hub_plan;
int b200gcn_spmm_hubs(const b200gcn_spmm_args* args, const b200gcn_hub_plan* plan, void* stream);

/* Backward of b200gcn_bignn_tail for d_in = d_out = 64 (the autograd of layers.py:56-58 + ngcf.py:96-98), one pass on
 * the tcgen05 tensor cores: given t = pre-activation (pre_out of the forward), the dropout mask and g_out = dL/d out,
 *   g_t = dL/dt (row-local backward of normalise / mask / LeakyReLU)              -> g_t [n, 64]
 *   [g_a | g_m] = g_t [W1 | W2] ;  g_p = g_a + g_m * x ;  g_x = g_a + g_m * p      -> g_p, g_x [n, 64]
 *   am = [p + x | p * x]                                                          -> am [n, 128] (optional)
 * The weight gradients are then ONE plain GEMM, [g_W1 | g_W2] = g_t^T am, and g_b1 = g_b2 = column sums of g_t. */
int b200gcn_bignn_tail_backward(const float* p, int64_t ldp, const float* x, int64_t ldx, const float* w1,
                                const float* w2, const float* t, int64_t ld_t, const uint8_t* keep, float drop_p,
                                float slope, int normalize, const float* g_out, int64_t ld_g, int64_t n, int32_t d_in,
                                int32_t d_out, float* g_p, float* g_x, float* g_t, float* am, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Phase chain: several b200gcn_spmm launches (each possibly a row range of one layer, or an identity-mode
 * publish) executed by ONE persistent cooperative kernel, ordered by device-side flags that also cross GPUs.
 * This is the row-sharded K-layer LightGCN.forward (lightgcn.py:70-81 extended to P ranks) as a
 * single launch per step: CTAs take row tiles from an atomic counter in phase order; a tile of phase p starts
 * once phase wait_phase[p] is complete on EVERY rank (and wait_local[p] on this one); when the last CTA of a rank leaves phase p it writes
 * `epoch` into flags[p][rank] of every rank (peer-mapped stores over NVLink, after a system-scope fence that
 * covers the rows the phase published through y_mc / y_peers).  No host-ordered barrier, no launch gaps, and
 * the flag latency of one phase hides behind the tiles of the next independent one.
 *   - every phase must share dim, val != NULL or == NULL, x2 == NULL, and have no hub rows (rows above the
 *     engine's long-row bound are NOT skipped here: the caller routes graphs with a hub plan to the per-launch
 *     path);
 *   - gathers of phase p read tables that other GPUs wrote during the same kernel: they use coherent loads;
 *   - all ranks must call with the same phase structure and epoch.  epoch starts at 1 and grows by 1 per call.
 */
#define B200GCN_CHAIN_MAX_PHASES 12
#define B200GCN_CHAIN_MAX_RANKS 16
#define B200GCN_CHAIN_SCRATCH_BYTES 256
typedef struct b200gcn_chain_sync {
  int32_t n_ranks;
  int32_t rank;
  uint32_t epoch;
  int32_t start_wait_phase;       /* phase whose flags of (epoch - 1) must have arrived from all ranks before the
                                     first store of this call (buffer reuse across calls); -1 = none */
  uint32_t* flags;                /* this rank's [MAX_PHASES][MAX_RANKS] words, zero-initialised once */
  uint32_t* const* flags_peers;   /* device array [n_ranks]: every rank's `flags` as mapped on this rank */
  int32_t* scratch;               /* device, B200GCN_CHAIN_SCRATCH_BYTES, zeroed by the call: int32 tile counter, int32
                                     arrival counter per phase, then (offset 64) uint64 %globaltimer stamps: [0] = first
                                     tile taken, [1 + p] = last CTA of this rank left phase p (diagnostics) */
  int8_t wait_phase[B200GCN_CHAIN_MAX_PHASES];   /* -1 = no dependency */
  int8_t wait_local[B200GCN_CHAIN_MAX_PHASES];   /* phase that must be complete on THIS rank only (rows of a local
                                                    buffer, e.g. the running layer sum, written by the phase right
                                                    before); -1 = none */
  int8_t merge_next[B200GCN_CHAIN_MAX_PHASES];   /* != 0: the tiles of phase p are interleaved evenly into the tile range
                                                    of phase p + 1 (a publish that only the phase AFTER p + 1 needs
                                                    shares the NVLink with the rows p + 1 produces instead of running
                                                    in front of it); both phases are left together */
} b200gcn_chain_sync;
int b200gcn_spmm_chain(const b200gcn_spmm_args* phases, int32_t n_phases, const b200gcn_chain_sync* sync,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * NGCF layer tail: everything of BiGNNConv.forward after propagate() (layers.py:56-58) plus the
 * per-layer ops of NGCF.forward (ngcf.py:96-98), one pass over the node rows:
 *   t = (p + x) W1^T + b1 + (p * x) W2^T + b2 ;  t = leaky_relu(t, slope) ;
 *   t = t * keep / (1 - drop_p) when keep != NULL ;  out = t / max(||t||_2, 1e-12) when normalize != 0
 * p, x: [n, d_in]; W1, W2: [d_out, d_in] row-major (nn.Linear.weight); b1, b2: [d_out];
 * keep: uint8 [n, d_out] or NULL; out: [n, ldo] (may be a column slice of the concat buffer, ngcf.py:100).
 * out2 (optional, [n, ldo2]) receives a second copy of `out`: the contiguous next-layer gather table, while
 * `out` is the strided concat slice (gathering from a 1 KB-strided slice costs ~60 % more: it uses a quarter
 * of the L2 sets / DRAM banks).
 * pre_out (optional, [n, ld_pre]) receives t before the activation (what BiGNNConv.forward returns).
 * d_in, d_out multiples of 4, <= 256. */
int b200gcn_bignn_tail(const float* p, int64_t ldp, const float* x, int64_t ldx, const float* w1,
                       const float* b1, const float* w2, const float* b2, int64_t n, int32_t d_in,
                       int32_t d_out, float slope, const uint8_t* keep, float drop_p, int normalize,
                       float* out, int64_t ldo, float* out2, int64_t ldo2, float* pre_out, int64_t ld_pre,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * The training step around the propagation.
 *
 * b200gcn_bpr_loss: what LightGCN.calculate_loss (lightgcn.py:83-110) / NGCF.calculate_loss (ngcf.py:106-123) do with
 * the propagated tables for one mini-batch of (user, pos_item, neg_item) ids, forward AND backward in one pass:
 *   s+ = <u_all[user], i_all[pos]>,  s- = <u_all[user], i_all[neg]>
 *   mf  = mean(-log(gamma + sigmoid(s+ - s-)))                                   recbole BPRLoss (gamma 1e-10)
 *   reg = (||reg_u[user]|| + ||reg_i[pos]|| + ||reg_i[neg]||) / B               recbole EmbLoss, require_pow == 0
 *       = (||.||^2 + ||.||^2 + ||.||^2) / B / 2                                  require_pow != 0
 *   loss_out[0] = mf + reg_weight * reg   (loss_out[1] = mf, loss_out[2..4] = the three batch norms; 5 floats)
 * reg_u / reg_i are the ego tables for LightGCN (lightgcn.py:103-107) and the propagated tables themselves for NGCF
 * (ngcf.py:121).  Gradients are ACCUMULATED (fp32 atomics) into g_u_all / g_i_all (d loss / d propagated rows) and
 * g_reg_u / g_reg_i (d loss / d EmbLoss rows); pass NULL pairs to skip them; the caller zeroes the tables.
 * Workspace: b200gcn_bpr_loss_workspace bytes. */
int b200gcn_bpr_loss_workspace(int64_t batch, size_t* bytes);
int b200gcn_bpr_loss(const float* u_all, int64_t ld_u, const float* i_all, int64_t ld_i, const float* reg_u,
                     int64_t ld_ru, const float* reg_i, int64_t ld_ri, const int64_t* user, const int64_t* pos,
                     const int64_t* neg, int64_t batch, int32_t dim, float gamma, float reg_weight, int require_pow,
                     float* g_u_all, int64_t ld_gu, float* g_i_all, int64_t ld_gi, float* g_reg_u, int64_t ld_gru,
                     float* g_reg_i, int64_t ld_gri, float* loss_out, void* workspace, size_t workspace_bytes,
                     void* stream);

/* torch.optim.Adam (amsgrad off; weight_decay = L2 added to the gradient) over one contiguous fp32 tensor, in place:
 * the optimizer.step() of recbole's Trainer over an embedding table.  `step` is the 1-based step count. */
int b200gcn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int64_t step, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Full-sort evaluation: `scores = restore_user_e[user] @ restore_item_e^T` of
 * LightGCN.full_sort_predict / NGCF.full_sort_predict (lightgcn.py:123-133, ngcf.py:138-150) on the tcgen05 tensor
 * cores with an error-compensated TF32 split (fp32-accurate scores).
 *
 * b200gcn_fullsort_topk: the top-k items per user row WITHOUT materialising the [n_users, n_items] matrix — what
 * RecBole's full-sort evaluator computes from full_sort_predict + history masking + torch.topk.  users: the gathered
 * user rows [n_users, dim]; items [n_items, dim]; items with id < first_item are excluded ([PAD] = 0 -> first_item 1);
 * hist_ptr [n_users + 1] / hist_items (ascending per user) = CSR of already-seen items to exclude (both NULL: none).
 * out_scores / out_ids [n_users, k], descending score, equal scores by ascending id; missing candidates (fewer than k
 * admissible items) are -inf / -1.  k <= 64, dim %% 8 == 0 and <= 512, n_items < 2^31.
 *
 * b200gcn_fullsort_scores: the dense [n_users, n_items] matrix (row stride ld_out) the reference API returns. */
int b200gcn_fullsort_topk_workspace(int64_t n_users, int64_t n_items, int32_t k, size_t* bytes);
int b200gcn_fullsort_topk(const float* users, int64_t ld_u, int64_t n_users, const float* items, int64_t ld_i,
                          int64_t n_items, int32_t dim, int32_t k, int64_t first_item, const int64_t* hist_ptr,
                          const int64_t* hist_items, float* out_scores, int64_t* out_ids, void* workspace,
                          size_t workspace_bytes, void* stream);
int b200gcn_fullsort_scores(const float* users, int64_t ld_u, int64_t n_users, const float* items, int64_t ld_i,
                            int64_t n_items, int32_t dim, float* out, int64_t ld_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * `.inter` atomic-file ingest (HOST code, h_ pointers are host memory): the tab-separated file format of
 * the reference's fixture tests/test_data/test/test.inter (header `user_id:token<TAB>item_id:token...`).  Tokens are
 * remapped to ids in first-appearance order starting at 1 (0 = RecBole's [PAD]) — what RecBole's Dataset hands to
 * get_norm_adj_mat through inter_feat (dataset.py:60-61).  open parses the file (mmap, one pass) and reports the
 * sizes; read copies the two id columns into caller buffers of n_inter int64 each; close frees the handle. */
int b200gcn_inter_open(const char* path, void** handle, int64_t* n_inter, int64_t* user_num, int64_t* item_num);
int b200gcn_inter_read(void* handle, int64_t* h_uid, int64_t* h_iid);
void b200gcn_inter_close(void* handle);

#ifdef __cplusplus
}
#endif
#endif /* B200GCN_H_ */
