// Full-sort evaluation: the step right after the propagation at eval time,
//   scores = restore_user_e[user] @ restore_item_e^T        lightgcn.py:123-133, ngcf.py:138-150
// followed, in RecBole's evaluator, by masking the seen items and torch.topk over the [batch, n_items] matrix
// (4096 x 1 M x 4 B = 16 GB per batch if materialised).  Here the contraction runs on the 5th-generation tensor cores
// (tcgen05.mma.kind::tf32, accumulators in TMEM) with the same error-compensated split as the NGCF tail
// (a = a_hi + a_lo, four TF32 products per fp32 product -> fp32-accurate scores, so the ids agree with an fp32
// matmul), and the top-k selection happens in the epilogue straight out of TMEM: the score matrix never exists.
//
//   grid = (user tiles of 128) x (item splits).  One CTA keeps its 128 user rows as the A operand, streams its slice of
//   the item table through the B operand (128 items per tile), and after every tile 128 threads (one per user = one
//   TMEM lane) scan the 128 fresh scores against that user's running k-th best; the rare survivors are checked
//   against the user's seen-item list (binary search) and replace the minimum of a k-slot candidate list in shared
//   memory.  Item tile t's MMAs run while tile t-1 is being scanned (two accumulators in TMEM).  A second small
//   kernel merges the per-split candidate lists of every user into the sorted top-k.
//
// `fullsort_scores` is the same kernel with a store epilogue: the dense matrix `full_sort_predict` returns.
#include "common.cuh"

#include <math_constants.h>
#include <stdint.h>

namespace b200gcn {
namespace {

constexpr int kFM = 128;                 // users per CTA = UMMA M = TMEM lanes
constexpr int kFN = 128;                 // items per tile = UMMA N
constexpr int kFKC = 64;                 // K elements per operand chunk held in shared memory
constexpr int kFChunks = kFKC / 4;       // 16-byte K chunks per row
constexpr int kFThreads = 512;
constexpr int kFRowsPerPass = kFThreads / 16;
constexpr int kFIters = kFM / kFRowsPerPass;   // 4
constexpr uint32_t kFSBO = 128;
constexpr uint32_t kFLbo = kFM * 16 + 16;      // 2064: +16 keeps the 16-lanes-per-row stores conflict-free
constexpr uint32_t kFBytesOp = kFChunks * kFLbo;   // 33,024 per hi / lo part
constexpr uint32_t kFOffAhi = 0, kFOffAlo = kFBytesOp, kFOffBhi = 2 * kFBytesOp, kFOffBlo = 3 * kFBytesOp;
constexpr uint32_t kFOffMisc = 4 * kFBytesOp;      // mbarrier, tmem slot
constexpr uint32_t kFOffHeap = kFOffMisc + 32;     // [k][128] scores, [k][128] ids
constexpr int kFMaxK = 64;
constexpr uint32_t kFTmemCols = 2 * kFN;
constexpr int kFDenseLd = kFN + 1;               // dense-mode staging row stride (floats): scalar accesses conflict-free

// order-preserving float <-> int code (atomicMax on floats of either sign)
__device__ __forceinline__ int fs_enc(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float fs_dec(int e) { return __int_as_float(e >= 0 ? e : e ^ 0x7fffffff); }

struct FsArgs {
  const float* users; int64_t ld_u; int64_t n_users;
  const float* items; int64_t ld_i; int64_t n_items;
  int32_t dim; int32_t k;
  int64_t first_item;
  const int64_t* hist_ptr; const int64_t* hist_items;
  int32_t n_splits; int64_t items_per_split;     // a multiple of kFN
  float* part_scores; int32_t* part_ids;          // [n_splits, n_users, k]
  int32_t* thr_g;                                 // [n_users] order-preserving int code of a lower bound of the user's k-th best
  float* dense; int64_t ld_dense;                 // dense-scores mode
};

__device__ __forceinline__ uint32_t fs_smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float fs_tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ uint64_t fs_desc(uint32_t smem_addr) {
  return uint64_t((smem_addr & 0x3ffffu) >> 4) | (uint64_t(kFLbo >> 4) << 16) | (uint64_t(kFSBO >> 4) << 32) |
         (uint64_t(1) << 46);
}
// D = F32, A = B = TF32, K-major both, N = 128, M = 128
constexpr uint32_t kFIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kFN >> 3) << 17) | (uint32_t(kFM >> 4) << 24);

__device__ __forceinline__ void fs_umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kFIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void fs_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool fs_elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// One 64-wide K chunk of 128 rows of `tbl` (rows row0 .., columns k0 ..) -> hi / lo operand parts in shared memory.
__device__ __forceinline__ void fs_load_rows(const float* tbl, int64_t ld, int64_t n_rows, int64_t row0, int k0,
                                             int width, float4 (&v)[kFIters]) {
  const int my_chunk = threadIdx.x & 15, my_row0 = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < kFIters; ++i) {
    const int64_t row = row0 + my_row0 + kFRowsPerPass * i;
    v[i] = (row < n_rows && my_chunk * 4 < width) ? ld_gather_f4(tbl + row * ld + k0 + my_chunk * 4)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void fs_store_operand(char* smem, uint32_t off_hi, uint32_t off_lo,
                                                 const float4 (&v)[kFIters]) {
  const int my_chunk = threadIdx.x & 15, my_row0 = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < kFIters; ++i) {
    const int r = my_row0 + kFRowsPerPass * i;
    const float4 h = make_float4(fs_tf32_hi(v[i].x), fs_tf32_hi(v[i].y), fs_tf32_hi(v[i].z), fs_tf32_hi(v[i].w));
    const uint32_t off = my_chunk * kFLbo + (r >> 3) * kFSBO + (r & 7) * 16;
    *reinterpret_cast<float4*>(smem + off_hi + off) = h;
    *reinterpret_cast<float4*>(smem + off_lo + off) = make_float4(v[i].x - h.x, v[i].y - h.y, v[i].z - h.z, v[i].w - h.w);
  }
}

template <bool TOPK>
__global__ void __launch_bounds__(kFThreads, 1) fullsort_kernel(const FsArgs a) {
  extern __shared__ __align__(1024) char smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + kFOffMisc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kFOffMisc + 16);
  float* heap_s = reinterpret_cast<float*>(smem + kFOffHeap);
  int* heap_i = reinterpret_cast<int*>(smem + kFOffHeap + size_t(a.k) * kFM * 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t u0 = int64_t(blockIdx.x) * kFM;                 // first user row of this CTA
  const int split = blockIdx.y;
  const int64_t it_lo = int64_t(split) * a.items_per_split;
  const int64_t it_hi = min(a.n_items, it_lo + a.items_per_split);
  const int n_tiles = it_hi > it_lo ? int((it_hi - it_lo + kFN - 1) / kFN) : 0;
  const int n_chunks = (a.dim + kFKC - 1) / kFKC;
  const bool single = n_chunks == 1;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fs_smem_u32(tmem_slot)),
                 "r"(kFTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fs_smem_u32(mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float4 v[kFIters];
  if (single) {   // the user rows are the A operand of every tile: staged once
    fs_load_rows(a.users, a.ld_u, a.n_users, u0, 0, a.dim, v);
    fs_store_operand(smem, kFOffAhi, kFOffAlo, v);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // ---- per-user candidate list (TOPK): thread t < 128 owns user u0 + t = TMEM lane t
  int cnt = 0, minpos = 0;
  float thr = -CUDART_INF_F;
  const int64_t my_user = u0 + tid;
  // seen items: a cursor into the user's ascending history list walks along with the item tiles (the split's items
  // ascend too), so the mask of a tile is four 32-bit words built from a value that is already in a register
  int64_t h_cur = 0, h_end = 0, h_next = INT64_MAX;
  if (TOPK && tid < kFM && my_user < a.n_users && a.hist_ptr != nullptr) {
    int64_t lo = a.hist_ptr[my_user], hi = a.hist_ptr[my_user + 1];
    h_end = hi;
    while (lo < hi) {   // first seen item >= it_lo
      const int64_t mid = (lo + hi) >> 1;
      if (a.hist_items[mid] < it_lo) lo = mid + 1; else hi = mid;
    }
    h_cur = lo;
    if (h_cur < h_end) h_next = a.hist_items[h_cur];
  }

  auto rescan = [&]() {   // minimum of the k candidates; among equal scores evict the largest id
    float m = heap_s[tid];
    int mi = heap_i[tid], mp = 0;
    for (int j = 1; j < a.k; ++j) {
      const float s = heap_s[j * kFM + tid];
      const int id = heap_i[j * kFM + tid];
      if (s < m || (s == m && id > mi)) { m = s; mi = id; mp = j; }
    }
    thr = m;
    minpos = mp;
    // any split's k-th best is a lower bound of the user's global k-th best: share it with the other splits
    atomicMax(a.thr_g + my_user, fs_enc(m));
  };
  auto offer = [&](float s, int64_t item) {
    if (item < a.first_item) return;
    if (cnt < a.k) {
      heap_s[cnt * kFM + tid] = s;
      heap_i[cnt * kFM + tid] = int(item);
      if (++cnt == a.k) rescan();
    } else {
      heap_s[minpos * kFM + tid] = s;
      heap_i[minpos * kFM + tid] = int(item);
      rescan();
    }
  };

  auto epilogue = [&](int t, uint32_t buf) {
    const int64_t item0 = it_lo + int64_t(t) * kFN;
    if (TOPK) {
      if (warp < 4) {   // warp w reads TMEM lanes 32 w .. 32 w + 31 (= users), all 128 columns, 32 at a time
        const bool live = my_user < a.n_users;
        // lower bound from the other splits (read early, used after the TMEM loads)
        const float g_thr = live ? fs_dec(__ldcg(a.thr_g + my_user)) : CUDART_INF_F;
        uint32_t seen[kFN / 32] = {0u, 0u, 0u, 0u};
        while (h_next < item0 + kFN) {   // seen items inside this tile
          seen[(h_next - item0) >> 5] |= 1u << ((h_next - item0) & 31);
          h_next = (++h_cur < h_end) ? a.hist_items[h_cur] : INT64_MAX;
        }
#pragma unroll
        for (int cg = 0; cg < kFN / 32; ++cg) {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16) + buf * uint32_t(kFN) + uint32_t(cg * 32);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr)
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // candidates of this column group: better than both bounds, inside the split, not seen
          const float bound = fmaxf(g_thr, cnt == a.k ? thr : -CUDART_INF_F);
          uint32_t cand = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) cand |= (__uint_as_float(r[j]) > bound ? 1u : 0u) << j;
          const int64_t left = it_hi - (item0 + cg * 32);
          if (left < 32) cand &= left <= 0 ? 0u : ((1u << left) - 1u);
          cand &= ~seen[cg];
          if (!live) cand = 0;
          while (cand) {
            const int j = __ffs(cand) - 1;
            cand &= cand - 1;
            float s = 0.f;
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) s = (jj == j) ? __uint_as_float(r[jj]) : s;   // no dynamic register index
            if (cnt < a.k || s > thr) offer(s, item0 + cg * 32 + j);
          }
        }
      }
    } else {   // dense scores: warp (q, h) stages columns 32 h .. of users 32 q .., then rows leave coalesced
      const int q = warp & 3, h = warp >> 2;
      uint32_t r[32];
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * uint32_t(kFN) + uint32_t(h * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
            "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
            "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* stg = heap_s;   // [128][kFDenseLd]
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[(q * 32 + lane) * kFDenseLd + h * 32 + j] = __uint_as_float(r[j]);
      __syncthreads();
      const int64_t n_cols = min(int64_t(kFN), it_hi - item0);
      for (int rr = warp; rr < kFM; rr += kFThreads / 32) {
        const int64_t user = u0 + rr;
        if (user >= a.n_users) break;
        float* dst = a.dense + user * a.ld_dense + item0;
#pragma unroll
        for (int c = lane; c < kFN; c += 32)
          if (c < n_cols) dst[c] = stg[rr * kFDenseLd + c];
      }
      __syncthreads();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  uint32_t n_commits = 0;
  if (single && n_tiles > 0) fs_load_rows(a.items, a.ld_i, it_hi, it_lo, 0, a.dim, v);
  for (int t = 0; t < n_tiles; ++t) {
    const int64_t item0 = it_lo + int64_t(t) * kFN;
    for (int kc = 0; kc < n_chunks; ++kc) {
      const int k0 = kc * kFKC;
      const int width = min(kFKC, a.dim - k0);
      if (n_commits > 0) {   // the previous group of MMAs has finished reading the operand buffers
        fs_mbar_wait(fs_smem_u32(mbar), (n_commits - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      if (!single) {
        fs_load_rows(a.users, a.ld_u, a.n_users, u0, k0, width, v);
        fs_store_operand(smem, kFOffAhi, kFOffAlo, v);
        fs_load_rows(a.items, a.ld_i, it_hi, item0, k0, width, v);
      }
      fs_store_operand(smem, kFOffBhi, kFOffBlo, v);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t base = fs_smem_u32(smem);
        const uint32_t tmem_d = tmem_base + uint32_t(t & 1) * uint32_t(kFN);
        if (fs_elect_one()) {
          uint32_t acc = kc > 0 ? 1u : 0u;
          const int ksteps = width / 8;
#pragma unroll
          for (int term = 0; term < 4; ++term) {   // smallest terms first: lo.lo, lo.hi, hi.lo, hi.hi
            const uint32_t a_off = (term < 2) ? kFOffAlo : kFOffAhi;
            const uint32_t b_off = (term == 0 || term == 2) ? kFOffBlo : kFOffBhi;
#pragma unroll 8
            for (int ks = 0; ks < ksteps; ++ks) {
              fs_umma(tmem_d, fs_desc(base + a_off + 2 * ks * kFLbo), fs_desc(base + b_off + 2 * ks * kFLbo), acc);
              acc = 1;
            }
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           fs_smem_u32(mbar))
                       : "memory");
        }
        __syncwarp();
      }
      ++n_commits;
    }
    // item rows of the next tile leave HBM / L2 while the tensor core works (single-chunk case)
    if (single && t + 1 < n_tiles) fs_load_rows(a.items, a.ld_i, it_hi, item0 + kFN, 0, a.dim, v);
    // scan / store the PREVIOUS tile under this tile's MMAs (its last commit was waited on above)
    if (t > 0) epilogue(t - 1, uint32_t((t - 1) & 1));
  }
  if (n_tiles > 0) {
    fs_mbar_wait(fs_smem_u32(mbar), (n_commits - 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue(n_tiles - 1, uint32_t((n_tiles - 1) & 1));
  }

  if (TOPK && tid < kFM && my_user < a.n_users) {
    float* ps = a.part_scores + (int64_t(split) * a.n_users + my_user) * a.k;
    int32_t* pi = a.part_ids + (int64_t(split) * a.n_users + my_user) * a.k;
    for (int j = 0; j < a.k; ++j) {
      ps[j] = j < cnt ? heap_s[j * kFM + tid] : -CUDART_INF_F;
      pi[j] = j < cnt ? heap_i[j * kFM + tid] : -1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kFTmemCols) : "memory");
  }
}

// Per user: the sorted top-k of the n_splits * k candidates (descending score; equal scores by ascending id).
constexpr int kMergeThreads = 128;
__global__ void __launch_bounds__(kMergeThreads) fullsort_merge_kernel(const float* __restrict__ part_scores,
                                                                       const int32_t* __restrict__ part_ids,
                                                                       int64_t n_users, int n_splits, int k,
                                                                       float* __restrict__ out_scores,
                                                                       int64_t* __restrict__ out_ids) {
  extern __shared__ char msm[];
  const int total = n_splits * k;
  float* cs = reinterpret_cast<float*>(msm);
  int* ci = reinterpret_cast<int*>(msm + size_t(total) * 4);
  __shared__ float red_s[kMergeThreads / 32];
  __shared__ int red_i[kMergeThreads / 32], red_p[kMergeThreads / 32];
  const int64_t user = blockIdx.x;
  for (int c = threadIdx.x; c < total; c += kMergeThreads) {
    const int sp = c / k, j = c % k;
    cs[c] = part_scores[(int64_t(sp) * n_users + user) * k + j];
    ci[c] = part_ids[(int64_t(sp) * n_users + user) * k + j];
  }
  __syncthreads();
  for (int round = 0; round < k; ++round) {
    float bs = -CUDART_INF_F;
    int bi = 0x7fffffff, bp = -1;
    for (int c = threadIdx.x; c < total; c += kMergeThreads) {
      const float s = cs[c];
      const int id = ci[c];
      if (id >= 0 && (s > bs || (s == bs && id < bi))) { bs = s; bi = id; bp = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, bs, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (op >= 0 && (bp < 0 || os > bs || (os == bs && oi < bi))) { bs = os; bi = oi; bp = op; }
    }
    if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = bs; red_i[threadIdx.x >> 5] = bi; red_p[threadIdx.x >> 5] = bp; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < kMergeThreads / 32; ++w)
        if (red_p[w] >= 0 && (bp < 0 || red_s[w] > bs || (red_s[w] == bs && red_i[w] < bi))) {
          bs = red_s[w]; bi = red_i[w]; bp = red_p[w];
        }
      out_scores[user * k + round] = bp >= 0 ? bs : -CUDART_INF_F;
      out_ids[user * k + round] = bp >= 0 ? int64_t(bi) : int64_t(-1);
      if (bp >= 0) ci[bp] = -1;   // taken
    }
    __syncthreads();
  }
}

int plan_splits(int64_t n_users, int64_t n_items, int sms, int32_t* n_splits, int64_t* items_per_split) {
  const int64_t user_tiles = (n_users + kFM - 1) / kFM;
  const int64_t item_tiles = (n_items + kFN - 1) / kFN;
  int64_t s = (int64_t(sms) * 2 + user_tiles - 1) / user_tiles;   // about two waves of CTAs
  if (s > item_tiles) s = item_tiles;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  const int64_t tiles_per_split = (item_tiles + s - 1) / s;
  *items_per_split = tiles_per_split * kFN;
  *n_splits = int32_t((item_tiles + tiles_per_split - 1) / tiles_per_split);
  return 0;
}

}  // namespace
}  // namespace b200gcn

using namespace b200gcn;

static int fs_sm_count(int* sms) {
  int dev = 0;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  B200_CHECK_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
  return B200GCN_OK;
}

static int fs_check_common(const float* users, int64_t ld_u, int64_t n_users, const float* items, int64_t ld_i,
                           int64_t n_items, int32_t dim) {
  B200_CHECK_ARG(n_users >= 0 && n_items >= 0 && n_items < (int64_t(1) << 31), "n_users / n_items out of range");
  B200_CHECK_ARG(dim > 0 && dim % 8 == 0 && dim <= 512, "dim=%d must be a multiple of 8 in [8, 512]", dim);
  B200_CHECK_ARG(users && items && aligned16(users) && aligned16(items) && ld_u % 4 == 0 && ld_i % 4 == 0 &&
                     ld_u >= dim && ld_i >= dim,
                 "users / items must be 16-byte aligned, leading dimensions %% 4 == 0 and >= dim");
  return B200GCN_OK;
}

extern "C" int b200gcn_fullsort_topk_workspace(int64_t n_users, int64_t n_items, int32_t k, size_t* bytes) {
  B200_CHECK_ARG(bytes && n_users >= 0 && n_items >= 0 && k >= 1 && k <= kFMaxK, "k must be in [1, %d]", kFMaxK);
  int sms = 148;
  int rc = fs_sm_count(&sms);
  if (rc) return rc;
  int32_t s = 1;
  int64_t ips = 0;
  plan_splits(n_users > 0 ? n_users : 1, n_items > 0 ? n_items : 1, sms, &s, &ips);
  *bytes = align_up(size_t(s) * size_t(n_users) * size_t(k) * 4) * 2 + align_up(size_t(n_users) * 4) + 256;
  return B200GCN_OK;
}

extern "C" int b200gcn_fullsort_topk(const float* users, int64_t ld_u, int64_t n_users, const float* items,
                                     int64_t ld_i, int64_t n_items, int32_t dim, int32_t k, int64_t first_item,
                                     const int64_t* hist_ptr, const int64_t* hist_items, float* out_scores,
                                     int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = fs_check_common(users, ld_u, n_users, items, ld_i, n_items, dim);
  if (rc) return rc;
  B200_CHECK_ARG(k >= 1 && k <= kFMaxK, "k must be in [1, %d]", kFMaxK);
  B200_CHECK_ARG(out_scores && out_ids && workspace, "NULL output / workspace");
  B200_CHECK_ARG((hist_ptr == nullptr) == (hist_items == nullptr) || hist_ptr != nullptr, "hist_ptr / hist_items");
  if (n_users == 0) return B200GCN_OK;
  size_t need = 0;
  rc = b200gcn_fullsort_topk_workspace(n_users, n_items, k, &need);
  if (rc) return rc;
  if (workspace_bytes < need) {
    set_error("workspace %zu < required %zu", workspace_bytes, need);
    return B200GCN_ERR_WORKSPACE;
  }
  int sms = 148;
  rc = fs_sm_count(&sms);
  if (rc) return rc;
  FsArgs a{};
  a.users = users; a.ld_u = ld_u; a.n_users = n_users; a.items = items; a.ld_i = ld_i; a.n_items = n_items;
  a.dim = dim; a.k = k; a.first_item = first_item; a.hist_ptr = hist_ptr; a.hist_items = hist_items;
  plan_splits(n_users, n_items > 0 ? n_items : 1, sms, &a.n_splits, &a.items_per_split);
  Carver cv(workspace);
  a.part_scores = cv.take<float>(size_t(a.n_splits) * n_users * k);
  a.part_ids = cv.take<int32_t>(size_t(a.n_splits) * n_users * k);
  a.thr_g = cv.take<int32_t>(size_t(n_users));
  B200_CHECK_CUDA(cudaMemsetAsync(a.thr_g, 0x80, size_t(n_users) * 4, st));   // code of a hugely negative float
  const size_t smem = kFOffHeap + size_t(k) * kFM * 8;
  B200_CHECK_CUDA(cudaFuncSetAttribute(fullsort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const dim3 grid(unsigned((n_users + kFM - 1) / kFM), unsigned(a.n_splits));
  fullsort_kernel<true><<<grid, kFThreads, smem, st>>>(a);
  B200_CHECK_LAUNCH();
  const size_t msmem = size_t(a.n_splits) * k * 8;
  fullsort_merge_kernel<<<unsigned(n_users), kMergeThreads, msmem, st>>>(a.part_scores, a.part_ids, n_users, a.n_splits,
                                                                          k, out_scores, out_ids);
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}

extern "C" int b200gcn_fullsort_scores(const float* users, int64_t ld_u, int64_t n_users, const float* items,
                                       int64_t ld_i, int64_t n_items, int32_t dim, float* out, int64_t ld_out,
                                       void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = fs_check_common(users, ld_u, n_users, items, ld_i, n_items, dim);
  if (rc) return rc;
  B200_CHECK_ARG(out && ld_out >= n_items, "out / ld_out");
  if (n_users == 0 || n_items == 0) return B200GCN_OK;
  int sms = 148;
  rc = fs_sm_count(&sms);
  if (rc) return rc;
  FsArgs a{};
  a.users = users; a.ld_u = ld_u; a.n_users = n_users; a.items = items; a.ld_i = ld_i; a.n_items = n_items;
  a.dim = dim; a.k = 0; a.dense = out; a.ld_dense = ld_out;
  plan_splits(n_users, n_items, sms, &a.n_splits, &a.items_per_split);
  const size_t smem = kFOffHeap + size_t(kFM) * kFDenseLd * 4;
  B200_CHECK_CUDA(cudaFuncSetAttribute(fullsort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const dim3 grid(unsigned((n_users + kFM - 1) / kFM), unsigned(a.n_splits));
  fullsort_kernel<false><<<grid, kFThreads, smem, st>>>(a);
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}
