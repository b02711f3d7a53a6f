// Graph build on the device: COO -> CSR (radix sort on packed (dst, src) keys), gcn_norm, bipartite
// normalisation, transpose, row ids, edge masking.  One-time work per model (per epoch for the
// augmentation models); everything is HBM-streaming integer work, so the kernels are plain
// grid-stride loops with coalesced accesses and the sort is cub's onesweep radix sort restricted to
// the significant key bits.
//
// Reference behaviour restated here: recbole_gnn/data/dataset.py:41-47 (edge_index_to_adj_t),
// :60-66 (COO assembly), :74,77 (gcn_norm call sites), :81-106 (get_bipartite_inter_mat);
// ngcf.py:81,89 (dropout_adj).
#include <cub/cub.cuh>
#include <stdarg.h>

#include "common.cuh"

namespace b200gcn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t n, int per_block = kThreads) {
  int64_t g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > (int64_t(1) << 30)) g = int64_t(1) << 30;
  return static_cast<int>(g);
}

// ---- key construction -------------------------------------------------------------------------
__global__ void make_keys_coo(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                              int64_t nnz, int src_bits, uint64_t* __restrict__ keys,
                              uint32_t* __restrict__ idx) {
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < nnz;
       e += int64_t(gridDim.x) * blockDim.x) {
    keys[e] = (uint64_t(dst[e]) << src_bits) | uint64_t(src[e]);
    idx[e] = uint32_t(e);
  }
}

// entry e < E is (dst = uid, src = iid + U); entry e >= E is (dst = iid + U, src = uid)   dataset.py:60-64
__global__ void make_keys_inter(const int64_t* __restrict__ uid, const int64_t* __restrict__ iid,
                                int64_t n_inter, int64_t user_num, int src_bits,
                                uint64_t* __restrict__ keys) {
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < n_inter;
       e += int64_t(gridDim.x) * blockDim.x) {
    uint64_t u = uint64_t(uid[e]), i = uint64_t(iid[e] + user_num);
    // edge_index1 = [row; col] -> source u, target i ; edge_index2 = [col; row] -> source i, target u
    keys[e] = (i << src_bits) | u;
    keys[e + n_inter] = (u << src_bits) | i;
  }
}

__global__ void fill_f32(float* p, int64_t n, float v) {
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < n;
       e += int64_t(gridDim.x) * blockDim.x)
    p[e] = v;
}

__global__ void check_range(const int64_t* __restrict__ a, int64_t n, int64_t lo, int64_t hi,
                            int* __restrict__ flag) {
  bool bad = false;
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < n;
       e += int64_t(gridDim.x) * blockDim.x) {
    int64_t v = a[e];
    bad |= (v < lo) | (v >= hi);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

// sorted keys -> rowptr + col (+ val/perm through the sorted original positions)
__global__ void keys_to_csr(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                            const float* __restrict__ w, int64_t nnz, int64_t n_dst, int src_bits,
                            int64_t* __restrict__ rowptr, int32_t* __restrict__ col,
                            float* __restrict__ val, int64_t* __restrict__ perm) {
  const uint64_t mask = (uint64_t(1) << src_bits) - 1;
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e <= nnz;
       e += int64_t(gridDim.x) * blockDim.x) {
    int64_t d = (e < nnz) ? int64_t(keys[e] >> src_bits) : n_dst;
    int64_t prev = (e > 0) ? int64_t(keys[e - 1] >> src_bits) : -1;
    for (int64_t r = prev + 1; r <= d; ++r) rowptr[r] = e;  // rows prev+1..d start at e
    if (e < nnz) {
      col[e] = int32_t(keys[e] & mask);
      if (idx != nullptr) {
        uint32_t p = idx[e];
        if (val != nullptr) val[e] = w[p];
        if (perm != nullptr) perm[e] = int64_t(p);
      }
    }
  }
}

struct SortPlan {
  size_t keys_bytes, idx_bytes, cub_bytes;
  int src_bits, end_bit;
};

int plan_sort(int64_t nnz, int64_t n_dst, int64_t n_src, bool with_idx, SortPlan* p) {
  p->src_bits = bits_for(n_src);
  p->end_bit = p->src_bits + bits_for(n_dst);
  p->keys_bytes = align_up(size_t(nnz) * 8);
  p->idx_bytes = with_idx ? align_up(size_t(nnz) * 4) : 0;
  size_t tmp = 0;
  cudaError_t e;
  if (with_idx) {
    e = cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                        (const uint32_t*)nullptr, (uint32_t*)nullptr, nnz, 0,
                                        p->end_bit);
  } else {
    e = cub::DeviceRadixSort::SortKeys(nullptr, tmp, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                       nnz, 0, p->end_bit);
  }
  if (e != cudaSuccess) {
    set_error("cub radix sort sizing failed: %s", cudaGetErrorString(e));
    return B200GCN_ERR_CUDA;
  }
  p->cub_bytes = align_up(tmp);
  return B200GCN_OK;
}

size_t sort_total(const SortPlan& p) {
  return 2 * p.keys_bytes + 2 * p.idx_bytes + p.cub_bytes + 256 /*flag*/;
}

}  // namespace
}  // namespace b200gcn

using namespace b200gcn;

extern "C" int b200gcn_abi_version(void) { return B200GCN_ABI_VERSION; }
extern "C" const char* b200gcn_last_error(void) { return g_err; }

extern "C" int b200gcn_device_info(int32_t* sm_count, int64_t* l2_bytes, int64_t* hbm_bytes,
                                   int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  B200_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (l2_bytes) *l2_bytes = prop.l2CacheSize;
  if (hbm_bytes) *hbm_bytes = int64_t(prop.totalGlobalMem);
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return B200GCN_OK;
}

extern "C" int b200gcn_csr_from_coo_workspace(int64_t nnz, int64_t n_dst, int64_t n_src,
                                              size_t* bytes) {
  B200_CHECK_ARG(bytes != nullptr, "bytes is NULL");
  B200_CHECK_ARG(nnz >= 0 && nnz < (int64_t(1) << 31), "nnz=%lld outside [0, 2^31)", (long long)nnz);
  B200_CHECK_ARG(n_dst >= 0 && n_src >= 0 && n_src < (int64_t(1) << 31) && n_dst < (int64_t(1) << 31),
                 "n_dst=%lld / n_src=%lld outside [0, 2^31)", (long long)n_dst, (long long)n_src);
  SortPlan p;
  int rc = plan_sort(nnz, n_dst, n_src, true, &p);
  if (rc) return rc;
  *bytes = sort_total(p);
  return B200GCN_OK;
}

static int sort_and_emit(uint64_t* keys_in, uint64_t* keys_out, uint32_t* idx_in, uint32_t* idx_out,
                         void* cub_ws, size_t cub_bytes, const SortPlan& p, const float* w,
                         int64_t nnz, int64_t n_dst, int64_t* rowptr, int32_t* col, float* val,
                         int64_t* perm, cudaStream_t st) {
  if (nnz > 0) {
    if (idx_in != nullptr) {
      B200_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys_in, keys_out, idx_in,
                                                      idx_out, nnz, 0, p.end_bit, st));
    } else {
      B200_CHECK_CUDA(
          cub::DeviceRadixSort::SortKeys(cub_ws, cub_bytes, keys_in, keys_out, nnz, 0, p.end_bit, st));
    }
  }
  keys_to_csr<<<grid_for(nnz + 1), kThreads, 0, st>>>(keys_out, idx_in ? idx_out : nullptr, w, nnz,
                                                      n_dst, p.src_bits, rowptr, col, val, perm);
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}

static int run_range_check(const int64_t* a, int64_t n, int64_t lo, int64_t hi, int* flag,
                           const char* what, cudaStream_t st) {
  if (n == 0) return B200GCN_OK;
  B200_CHECK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  check_range<<<grid_for(n, kThreads * 8), kThreads, 0, st>>>(a, n, lo, hi, flag);
  B200_CHECK_LAUNCH();
  int h = 0;
  B200_CHECK_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  B200_CHECK_CUDA(cudaStreamSynchronize(st));
  if (h) {
    set_error("%s holds an id outside [%lld, %lld)", what, (long long)lo, (long long)hi);
    return B200GCN_ERR_RANGE;
  }
  return B200GCN_OK;
}

extern "C" int b200gcn_csr_from_coo(const int64_t* src, const int64_t* dst, const float* w,
                                    int64_t nnz, int64_t n_dst, int64_t n_src, int64_t* rowptr,
                                    int32_t* col, float* val, int64_t* perm, void* workspace,
                                    size_t workspace_bytes, int check, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t need = 0;
  int rc = b200gcn_csr_from_coo_workspace(nnz, n_dst, n_src, &need);
  if (rc) return rc;
  B200_CHECK_ARG(rowptr && (nnz == 0 || (src && dst && col)), "NULL src/dst/rowptr/col");
  B200_CHECK_ARG(!(w != nullptr && val == nullptr), "w given but val is NULL");
  if (workspace_bytes < need || workspace == nullptr) {
    set_error("workspace %zu < required %zu", workspace_bytes, need);
    return B200GCN_ERR_WORKSPACE;
  }
  SortPlan p;
  rc = plan_sort(nnz, n_dst, n_src, true, &p);
  if (rc) return rc;
  Carver c(workspace);
  uint64_t* keys_in = c.take<uint64_t>(nnz);
  uint64_t* keys_out = c.take<uint64_t>(nnz);
  uint32_t* idx_in = c.take<uint32_t>(nnz);
  uint32_t* idx_out = c.take<uint32_t>(nnz);
  char* cub_ws = c.take<char>(p.cub_bytes);
  int* flag = c.take<int>(1);
  if (check) {
    rc = run_range_check(src, nnz, 0, n_src, flag, "edge_index[0] (source ids)", st);
    if (rc) return rc;
    rc = run_range_check(dst, nnz, 0, n_dst, flag, "edge_index[1] (destination ids)", st);
    if (rc) return rc;
  }
  if (nnz > 0) {
    make_keys_coo<<<grid_for(nnz, kThreads * 4), kThreads, 0, st>>>(src, dst, nnz, p.src_bits, keys_in,
                                                                   idx_in);
    B200_CHECK_LAUNCH();
  }
  rc = sort_and_emit(keys_in, keys_out, idx_in, idx_out, cub_ws, p.cub_bytes, p, w, nnz, n_dst, rowptr,
                     col, w ? val : nullptr, perm, st);
  if (rc) return rc;
  if (w == nullptr && val != nullptr && nnz > 0) {
    // caller asked for an explicit value array without weights: ones
    fill_f32<<<grid_for(nnz, kThreads * 4), kThreads, 0, st>>>(val, nnz, 1.0f);
    B200_CHECK_LAUNCH();
  }
  return B200GCN_OK;
}

extern "C" int b200gcn_csr_from_interactions_workspace(int64_t n_inter, int64_t user_num,
                                                       int64_t item_num, size_t* bytes) {
  B200_CHECK_ARG(bytes != nullptr, "bytes is NULL");
  int64_t nnz = 2 * n_inter, n = user_num + item_num;
  B200_CHECK_ARG(n_inter >= 0 && nnz < (int64_t(1) << 31), "2*n_inter=%lld outside [0, 2^31)",
                 (long long)nnz);
  B200_CHECK_ARG(user_num >= 0 && item_num >= 0 && n < (int64_t(1) << 31), "user_num+item_num too large");
  SortPlan p;
  int rc = plan_sort(nnz, n, n, false, &p);
  if (rc) return rc;
  *bytes = sort_total(p);
  return B200GCN_OK;
}

extern "C" int b200gcn_csr_from_interactions(const int64_t* uid, const int64_t* iid, int64_t n_inter,
                                             int64_t user_num, int64_t item_num, int64_t* rowptr,
                                             int32_t* col, void* workspace, size_t workspace_bytes,
                                             int check, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t need = 0;
  int rc = b200gcn_csr_from_interactions_workspace(n_inter, user_num, item_num, &need);
  if (rc) return rc;
  B200_CHECK_ARG(rowptr && (n_inter == 0 || (uid && iid && col)), "NULL uid/iid/rowptr/col");
  if (workspace_bytes < need || workspace == nullptr) {
    set_error("workspace %zu < required %zu", workspace_bytes, need);
    return B200GCN_ERR_WORKSPACE;
  }
  const int64_t nnz = 2 * n_inter, n = user_num + item_num;
  SortPlan p;
  rc = plan_sort(nnz, n, n, false, &p);
  if (rc) return rc;
  Carver c(workspace);
  uint64_t* keys_in = c.take<uint64_t>(nnz);
  uint64_t* keys_out = c.take<uint64_t>(nnz);
  char* cub_ws = c.take<char>(p.cub_bytes);
  int* flag = c.take<int>(1);
  if (check) {
    rc = run_range_check(uid, n_inter, 0, user_num, flag, "inter_feat[uid]", st);
    if (rc) return rc;
    rc = run_range_check(iid, n_inter, 0, item_num, flag, "inter_feat[iid]", st);
    if (rc) return rc;
  }
  if (n_inter > 0) {
    make_keys_inter<<<grid_for(n_inter, kThreads * 4), kThreads, 0, st>>>(uid, iid, n_inter, user_num,
                                                                         p.src_bits, keys_in);
    B200_CHECK_LAUNCH();
  }
  return sort_and_emit(keys_in, keys_out, nullptr, nullptr, cub_ws, p.cub_bytes, p, nullptr, nnz, n,
                       rowptr, col, nullptr, nullptr, st);
}

// ---- gcn_norm ----------------------------------------------------------------------------------
namespace b200gcn {
namespace {

// 8 lanes per row, shuffle-reduced row sum; val == NULL -> entry count
__global__ void row_degree(const int64_t* __restrict__ rowptr, const float* __restrict__ val, int64_t n,
                           float* __restrict__ dis) {
  constexpr int G = 8;
  int64_t row = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) / G;
  int lig = threadIdx.x & (G - 1);
  if (row >= n) return;
  int64_t b = rowptr[row], e = rowptr[row + 1];
  float s = 0.f;
  if (val == nullptr) {
    s = float(e - b);
  } else {
    for (int64_t k = b + lig; k < e; k += G) s += val[k];
    unsigned gm = 0xffu << ((threadIdx.x & 31) & ~(G - 1));
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(gm, s, o, G);
  }
  if (lig == 0) {
    // deg.pow(-0.5); inf -> 0  (PyG gcn_norm).  IEEE-rounded sqrt and division: bit-identical to ATen's CPU
    // result on the golden fixtures, within 1 ulp in general (ATen's own pow(-0.5) and 1/sqrt paths differ
    // from each other by 1 ulp on ~0.7 % of degrees).
    float d = 1.0f / sqrtf(s);
    dis[row] = isinf(d) ? 0.f : d;
  }
}

__global__ void scale_entries(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                              const float* __restrict__ val_in, const float* __restrict__ dis, int64_t n,
                              float* __restrict__ val_out) {
  constexpr int G = 8;
  int64_t row = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) / G;
  int lig = threadIdx.x & (G - 1);
  if (row >= n) return;
  int64_t b = rowptr[row], e = rowptr[row + 1];
  float dr = dis[row];
  for (int64_t k = b + lig; k < e; k += G) {
    float w = val_in ? val_in[k] : 1.0f;
    val_out[k] = (dis[col[k]] * w) * dr;  // dis[row0] * w * dis[row1], row0 = source, row1 = target
  }
}

}  // namespace
}  // namespace b200gcn

extern "C" int b200gcn_gcn_norm_csr(const int64_t* rowptr, const int32_t* col, const float* val_in,
                                    float* val_out, float* dis_out, int64_t n, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // val_out may be NULL only for a graph without entries (no kernel dereferences it then)
  B200_CHECK_ARG(rowptr && n >= 0, "NULL rowptr");
  if (n == 0) return B200GCN_OK;
  float* dis = dis_out;
  bool own = false;
  if (dis == nullptr) {
    B200_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&dis), size_t(n) * 4, st));
    own = true;
  }
  row_degree<<<grid_for(n * 8), kThreads, 0, st>>>(rowptr, val_in, n, dis);
  B200_CHECK_LAUNCH();
  scale_entries<<<grid_for(n * 8), kThreads, 0, st>>>(rowptr, col, val_in, dis, n, val_out);
  B200_CHECK_LAUNCH();
  if (own) B200_CHECK_CUDA(cudaFreeAsync(dis, st));
  return B200GCN_OK;
}

// ---- get_bipartite_inter_mat weights -------------------------------------------------------------
namespace b200gcn {
namespace {
__global__ void count_ids(const int64_t* __restrict__ ids, int64_t nnz, int* __restrict__ cnt) {
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < nnz;
       e += int64_t(gridDim.x) * blockDim.x)
    atomicAdd(&cnt[ids[e]], 1);
}
__global__ void bip_weights(const int64_t* __restrict__ row_ids, const int64_t* __restrict__ col_ids,
                            int64_t nnz, const int* __restrict__ rcnt, const int* __restrict__ ccnt,
                            int row_norm, float* __restrict__ w) {
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < nnz;
       e += int64_t(gridDim.x) * blockDim.x) {
    int rc = rcnt[row_ids[e]];
    float rd = rc == 0 ? 1.0f : float(rc);
    if (row_norm) {
      w[e] = 1.0f / rd;  // dataset.py:95-96
    } else {
      int cc = ccnt[col_ids[e]];
      float cd = cc == 0 ? 1.0f : float(cc);
      w[e] = (1.0f / sqrtf(rd)) * (1.0f / sqrtf(cd));  // dataset.py:101-104
    }
  }
}
}  // namespace
}  // namespace b200gcn

extern "C" int b200gcn_bipartite_norm_coo(const int64_t* row_ids, const int64_t* col_ids, int64_t nnz,
                                          int64_t n_row, int64_t n_col, int row_norm, float* w_out,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(nnz >= 0 && n_row >= 0 && n_col >= 0, "negative size");
  if (nnz == 0) return B200GCN_OK;
  B200_CHECK_ARG(row_ids && col_ids && w_out, "NULL ids/w_out");
  size_t need = size_t(n_row + n_col) * 4;
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("workspace %zu < required %zu", workspace_bytes, need);
    return B200GCN_ERR_WORKSPACE;
  }
  int* rcnt = static_cast<int*>(workspace);
  int* ccnt = rcnt + n_row;
  B200_CHECK_CUDA(cudaMemsetAsync(rcnt, 0, need, st));
  count_ids<<<grid_for(nnz, kThreads * 4), kThreads, 0, st>>>(row_ids, nnz, rcnt);
  B200_CHECK_LAUNCH();
  if (!row_norm) {
    count_ids<<<grid_for(nnz, kThreads * 4), kThreads, 0, st>>>(col_ids, nnz, ccnt);
    B200_CHECK_LAUNCH();
  }
  bip_weights<<<grid_for(nnz, kThreads * 4), kThreads, 0, st>>>(row_ids, col_ids, nnz, rcnt, ccnt,
                                                               row_norm, w_out);
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}

// ---- row ids / transpose ----------------------------------------------------------------------------
namespace b200gcn {
namespace {
__device__ __forceinline__ int64_t row_of_entry(const int64_t* __restrict__ rowptr, int64_t n_rows,
                                                int64_t e) {
  // largest r with rowptr[r] <= e
  int64_t lo = 0, hi = n_rows;
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (rowptr[mid] <= e) lo = mid; else hi = mid;
  }
  return lo;
}
__global__ void expand_rows(const int64_t* __restrict__ rowptr, int64_t n_rows, int64_t nnz,
                            int64_t* __restrict__ row_ids) {
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < nnz;
       e += int64_t(gridDim.x) * blockDim.x)
    row_ids[e] = row_of_entry(rowptr, n_rows, e);
}
__global__ void make_keys_transpose(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                    int64_t n_rows, int64_t nnz, int src_bits,
                                    uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < nnz;
       e += int64_t(gridDim.x) * blockDim.x) {
    uint64_t r = uint64_t(row_of_entry(rowptr, n_rows, e));
    keys[e] = (uint64_t(uint32_t(col[e])) << src_bits) | r;  // new destination = old source
    idx[e] = uint32_t(e);
  }
}
}  // namespace
}  // namespace b200gcn

extern "C" int b200gcn_csr_row_ids(const int64_t* rowptr, int64_t n_rows, int64_t nnz,
                                   int64_t* row_ids, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(nnz >= 0 && n_rows >= 0, "negative size");
  if (nnz == 0) return B200GCN_OK;
  B200_CHECK_ARG(rowptr && row_ids, "NULL rowptr/row_ids");
  expand_rows<<<grid_for(nnz, kThreads * 2), kThreads, 0, st>>>(rowptr, n_rows, nnz, row_ids);
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}

extern "C" int b200gcn_csr_transpose_workspace(int64_t nnz, int64_t n_rows, int64_t n_cols,
                                               size_t* bytes) {
  return b200gcn_csr_from_coo_workspace(nnz, n_cols, n_rows, bytes);
}

extern "C" int b200gcn_csr_transpose(const int64_t* rowptr, const int32_t* col, const float* val,
                                     int64_t n_rows, int64_t n_cols, int64_t nnz, int64_t* rowptr_t,
                                     int32_t* col_t, float* val_t, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t need = 0;
  int rc = b200gcn_csr_transpose_workspace(nnz, n_rows, n_cols, &need);
  if (rc) return rc;
  B200_CHECK_ARG(rowptr && rowptr_t && (nnz == 0 || (col && col_t)), "NULL rowptr/col");
  B200_CHECK_ARG(!(val != nullptr && val_t == nullptr), "val given but val_t is NULL");
  if (workspace_bytes < need || workspace == nullptr) {
    set_error("workspace %zu < required %zu", workspace_bytes, need);
    return B200GCN_ERR_WORKSPACE;
  }
  SortPlan p;
  rc = plan_sort(nnz, n_cols, n_rows, true, &p);
  if (rc) return rc;
  Carver c(workspace);
  uint64_t* keys_in = c.take<uint64_t>(nnz);
  uint64_t* keys_out = c.take<uint64_t>(nnz);
  uint32_t* idx_in = c.take<uint32_t>(nnz);
  uint32_t* idx_out = c.take<uint32_t>(nnz);
  char* cub_ws = c.take<char>(p.cub_bytes);
  if (nnz > 0) {
    make_keys_transpose<<<grid_for(nnz, kThreads * 2), kThreads, 0, st>>>(rowptr, col, n_rows, nnz,
                                                                         p.src_bits, keys_in, idx_in);
    B200_CHECK_LAUNCH();
  }
  return sort_and_emit(keys_in, keys_out, idx_in, idx_out, cub_ws, p.cub_bytes, p, val, nnz, n_cols,
                       rowptr_t, col_t, val ? val_t : nullptr, nullptr, st);
}

// ---- edge masking (dropout_adj without re-sort) ---------------------------------------------------------
namespace b200gcn {
namespace {
struct U8ToI64 {
  __host__ __device__ int64_t operator()(uint8_t v) const { return v ? 1 : 0; }
};
__global__ void compact_entries(const int32_t* __restrict__ col, const float* __restrict__ val,
                                const uint8_t* __restrict__ keep, const int64_t* __restrict__ pos,
                                int64_t nnz, int32_t* __restrict__ col_out, float* __restrict__ val_out) {
  for (int64_t e = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; e < nnz;
       e += int64_t(gridDim.x) * blockDim.x) {
    if (keep[e]) {
      int64_t p = pos[e];
      col_out[p] = col[e];
      if (val_out) val_out[p] = val ? val[e] : 1.0f;
    }
  }
}
__global__ void remap_rowptr(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ pos,
                             const uint8_t* __restrict__ keep, int64_t n_rows, int64_t nnz,
                             int64_t* __restrict__ rowptr_out) {
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r <= n_rows;
       r += int64_t(gridDim.x) * blockDim.x) {
    int64_t e = rowptr[r];
    rowptr_out[r] = (e < nnz) ? pos[e] : (nnz > 0 ? pos[nnz - 1] + (keep[nnz - 1] ? 1 : 0) : 0);
  }
}
}  // namespace
}  // namespace b200gcn

extern "C" int b200gcn_csr_mask_workspace(int64_t nnz, int64_t n_rows, size_t* bytes) {
  B200_CHECK_ARG(bytes != nullptr && nnz >= 0 && n_rows >= 0, "bad args");
  size_t tmp = 0;
  cub::TransformInputIterator<int64_t, U8ToI64, const uint8_t*> it(nullptr, U8ToI64());
  B200_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, it, (int64_t*)nullptr, nnz));
  *bytes = align_up(size_t(nnz) * 8) + align_up(tmp) + 256;
  return B200GCN_OK;
}

extern "C" int b200gcn_csr_mask(const int64_t* rowptr, const int32_t* col, const float* val,
                                const uint8_t* keep, int64_t n_rows, int64_t nnz, int64_t* rowptr_out,
                                int32_t* col_out, float* val_out, int64_t* h_nnz_out, void* workspace,
                                size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t need = 0;
  int rc = b200gcn_csr_mask_workspace(nnz, n_rows, &need);
  if (rc) return rc;
  B200_CHECK_ARG(rowptr && rowptr_out && h_nnz_out && (nnz == 0 || (col && keep && col_out)),
                 "NULL argument");
  if (workspace_bytes < need || workspace == nullptr) {
    set_error("workspace %zu < required %zu", workspace_bytes, need);
    return B200GCN_ERR_WORKSPACE;
  }
  Carver c(workspace);
  int64_t* pos = c.take<int64_t>(nnz);
  size_t tmp = need - c.off;
  char* cub_ws = c.take<char>(0);
  if (nnz > 0) {
    cub::TransformInputIterator<int64_t, U8ToI64, const uint8_t*> it(keep, U8ToI64());
    B200_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(cub_ws, tmp, it, pos, nnz, st));
    compact_entries<<<grid_for(nnz, kThreads * 2), kThreads, 0, st>>>(col, val, keep, pos, nnz, col_out,
                                                                     val_out);
    B200_CHECK_LAUNCH();
  }
  remap_rowptr<<<grid_for(n_rows + 1), kThreads, 0, st>>>(rowptr, pos, keep, n_rows, nnz, rowptr_out);
  B200_CHECK_LAUNCH();
  B200_CHECK_CUDA(cudaMemcpyAsync(h_nnz_out, rowptr_out + n_rows, sizeof(int64_t),
                                  cudaMemcpyDeviceToHost, st));
  B200_CHECK_CUDA(cudaStreamSynchronize(st));
  return B200GCN_OK;
}
