// y = A x over a CSR keyed by destination row: the K-layer hot loop of LightGCN / NGCF / SimGCL
// (recbole_gnn/model/layers.py:13-20, 31-35, 55 in the reference).  sm_100a, HBM-gather bound, tensor cores off
// by design (a sparse gather, not a dense contraction).
//
// Kernels in this file
//   spmm_warp_kernel   (v2, default for dim <= 128)  one warp per destination row, see the comment above it;
//   spmm_rows_kernel   (v1, dim > 128 and the A/B baseline)  G lanes per row, V float4 per lane, (col, val)
//                      chunks read coalesced and broadcast with shuffles, 8 gathers in flight per lane;
//   spmm_hub_chunk_kernel + spmm_hub_finish_kernel  rows above `long_row` entries, cut into chunks (one CTA
//                      each), partial rows added in chunk order -> deterministic;
//   spmm_hub_kernel    one CTA per hub row (the plan-free hub path of b200gcn_spmm_planned);
//   rows_identity_kernel  rowptr == NULL: p = x, epilogues only.
// Every kernel ends in finish_row(): the epilogues run on the registers that hold the finished row p —
// SimGCL sign-noise, y store, running layer mean, and the NVLink stores of the row-sharded exchange — so no
// [N, D] intermediate is re-read.
#include "common.cuh"

#include <string.h>

namespace b200gcn {
namespace {

constexpr int kCta = 256;
constexpr int kUnroll = 8;  // gathers in flight per lane

struct Philox {
  // Philox4x32-10 (Salmon et al. 2011), counter = (c0, c1, 0, 0), key = 64-bit seed
  static __device__ __forceinline__ uint4 draw(uint64_t seed, uint32_t c0, uint32_t c1) {
    uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32);
    uint4 c = make_uint4(c0, c1, 0u, 0u);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
      c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    return c;
  }
  static __device__ __forceinline__ float u01(uint32_t v) { return float(v >> 8) * (1.0f / 16777216.0f); }
};

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (G == 32) {
    return 0xffffffffu;
  } else {
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << G) - 1u) << (lane & ~unsigned(G - 1));
  }
}

template <int G>
__device__ __forceinline__ float group_sum(float v, unsigned gm) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gm, v, o, G);
  return v;
}

__device__ __forceinline__ const float* src_row(const b200gcn_spmm_args& a, int c) {
  if (a.x2 != nullptr && int64_t(c) >= a.x_split) return a.x2 + (int64_t(c) - a.x_split) * a.ldx;
  return a.x + int64_t(c) * a.ldx;
}

// Accumulate entries [beg, end) of one row into acc[V] (lane owns float columns lig*4 + k*G*4 .. +3).
template <int G, int V, bool HAS_VAL, bool TWO_TABLES>
__device__ __forceinline__ void gather_row(const b200gcn_spmm_args& a, int64_t beg, int64_t end, int lig,
                                           unsigned gm, float4 (&acc)[V]) {
  const int cbase = lig * 4;
  const int D = a.dim;
  int c_next = 0;
  float w_next = 0.f;
  if (beg + lig < end) {
    c_next = ld_stream_i32(a.col + beg + lig);
    w_next = HAS_VAL ? ld_stream_f32(a.val + beg + lig) : 1.0f;
  }
  for (int64_t e = beg; e < end; e += G) {
    const int c_mine = c_next;
    const float w_mine = w_next;
    const int64_t nxt = e + G + lig;
    if (nxt < end) {  // prefetch the next chunk of the index stream
      c_next = ld_stream_i32(a.col + nxt);
      w_next = HAS_VAL ? ld_stream_f32(a.val + nxt) : 1.0f;
    }
    const int n = int(min(int64_t(G), end - e));
#pragma unroll
    for (int j0 = 0; j0 < G; j0 += kUnroll) {
      if (j0 >= n) break;
      float4 xv[kUnroll][V];
      float wj[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int cj = __shfl_sync(gm, c_mine, j0 + u, G);
        wj[u] = __shfl_sync(gm, w_mine, j0 + u, G);
        const bool live = (j0 + u) < n;
        const float* row;
        if (TWO_TABLES) row = src_row(a, cj);
        else row = a.x + int64_t(cj) * a.ldx;
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const int cc = cbase + k * G * 4;
          if (live && cc < D) xv[u][k] = ld_gather_f4(row + cc);
          else xv[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (!live) wj[u] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
        for (int k = 0; k < V; ++k) {
          acc[k].x = fmaf(wj[u], xv[u][k].x, acc[k].x);
          acc[k].y = fmaf(wj[u], xv[u][k].y, acc[k].y);
          acc[k].z = fmaf(wj[u], xv[u][k].z, acc[k].z);
          acc[k].w = fmaf(wj[u], xv[u][k].w, acc[k].w);
        }
      }
    }
  }
}

__device__ __forceinline__ float sgn(float v) { return float(v > 0.f) - float(v < 0.f); }

// Everything that happens to a finished row p (held in acc) before it leaves the registers.
template <int G, int V>
__device__ __forceinline__ void finish_row(const b200gcn_spmm_args& a, int64_t row, int lig, unsigned gm,
                                           float4 (&acc)[V]) {
  const int D = a.dim;
  const int cbase = lig * 4;
  if (a.eps != 0.f) {  // SimGCL: p += sign(p) * normalize(noise) * eps        simgcl.py:31-32
    float4 nz[V];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int cc = cbase + k * G * 4;
      nz[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (cc < D) {
        if (a.noise != nullptr) {
          nz[k] = *reinterpret_cast<const float4*>(a.noise + row * a.ldn + cc);
        } else {
          uint4 r = Philox::draw(a.seed, uint32_t(row), (uint32_t(uint64_t(row) >> 32) << 16) | uint32_t(cc >> 2));
          nz[k] = make_float4(Philox::u01(r.x), Philox::u01(r.y), Philox::u01(r.z), Philox::u01(r.w));
        }
        ss += nz[k].x * nz[k].x + nz[k].y * nz[k].y + nz[k].z * nz[k].z + nz[k].w * nz[k].w;
      }
    }
    ss = group_sum<G>(ss, gm);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
#pragma unroll
    for (int k = 0; k < V; ++k) {
      acc[k].x = acc[k].x + sgn(acc[k].x) * (nz[k].x / denom) * a.eps;
      acc[k].y = acc[k].y + sgn(acc[k].y) * (nz[k].y / denom) * a.eps;
      acc[k].z = acc[k].z + sgn(acc[k].z) * (nz[k].z / denom) * a.eps;
      acc[k].w = acc[k].w + sgn(acc[k].w) * (nz[k].w / denom) * a.eps;
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int cc = cbase + k * G * 4;
    if (cc >= D) continue;
    if (a.y != nullptr) st_stream_f4(a.y + row * a.ldy + cc, acc[k]);
    if (a.y_mc != nullptr) {  // one store, replicated to every rank by the NVSwitch
      st_multimem_f4(a.y_mc + (a.y_peer_row0 + row) * a.ld_peer + cc, acc[k]);
    } else if (a.n_peers > 0) {  // peer-mapped next-layer tables (own rank included)
      const int64_t off = (a.y_peer_row0 + row) * a.ld_peer + cc;
      const uint32_t need = a.peer_need != nullptr ? a.peer_need[row] : 0xffffffffu;   // halo-only: skip non-readers
      for (int q = 0; q < a.n_peers; ++q)
        if ((need >> q) & 1u) st_peer_f4(a.y_peers[q] + off, acc[k]);
    }
    if (a.acc_out != nullptr) {  // layer combine                                lightgcn.py:77-78
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
      bool have = false;
      if (a.acc_in != nullptr) {
        const float* ai = (a.acc_in2 != nullptr && row >= a.acc_split)
                              ? a.acc_in2 + (row - a.acc_split) * a.ld_acc_in
                              : a.acc_in + row * a.ld_acc_in;
        s = *reinterpret_cast<const float4*>(ai + cc);
        have = true;
      }
      for (int e = 0; e < a.n_acc_extra; ++e) {
        const float4 t = *reinterpret_cast<const float4*>(a.acc_extra[e] + row * a.ld_acc_extra + cc);
        if (have) { s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; } else { s = t; have = true; }
      }
      if (have) { s.x += acc[k].x; s.y += acc[k].y; s.z += acc[k].z; s.w += acc[k].w; } else { s = acc[k]; }
      s.x *= a.acc_scale; s.y *= a.acc_scale; s.z *= a.acc_scale; s.w *= a.acc_scale;
      st_stream_f4(a.acc_out + row * a.ld_acc_out + cc, s);
    }
  }
}

// Rows that left over NVLink must be visible system-wide before the caller's cross-GPU barrier.  Called
// ONCE per thread after its last row: a per-row MEMBAR.SC.SYS stalls the warp for a full NVLink round trip
// (measured at 2 GPUs: 4.6 ms per layer with the per-row fence vs 3.2 ms for the gathers alone).
__device__ __forceinline__ void publish_fence(const b200gcn_spmm_args& a) {
  if (a.n_peers > 0 || a.y_mc != nullptr) __threadfence_system();
}

template <int G, int V, bool HAS_VAL, bool TWO_TABLES>
__global__ void __launch_bounds__(kCta) spmm_rows_kernel(const b200gcn_spmm_args a, int64_t long_row) {
  const int lig = threadIdx.x & (G - 1);
  const unsigned gm = group_mask<G>();
  const int64_t row = (int64_t(blockIdx.x) * kCta + threadIdx.x) / G;
  if (row >= a.n_rows) return;
  const int64_t beg = a.rowptr[row], end = a.rowptr[row + 1];
  if (end - beg > long_row) return;  // hub row: handled by spmm_hub_kernel
  float4 acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  gather_row<G, V, HAS_VAL, TWO_TABLES>(a, beg, end, lig, gm, acc);
  finish_row<G, V>(a, row, lig, gm, acc);
  publish_fence(a);
}


// ---------------------------------------------------------------------------------------------------
// v2: one WARP per destination row, R consecutive rows per warp.
//
// ncu on v1 (profiles/r1_v1_spmm_rows_kernel.txt): DRAM only 58-65 % busy, long_scoreboard dominating:
// latency-bound (dependent chain rowptr -> col -> row, variable-mask shuffles costing a MATCH/REDUX
// sequence each).  v2:
//   * the 32/G sub-groups of the warp take alternating entries of the SAME row (two 256 B rows per
//     gather instruction at D = 64), so the loop is warp-uniform: no divergence, no shuffles in the loop;
//   * index entries are read by broadcast loads (one 4-byte request per sub-group) that hit L1;
//     gathered rows bypass L1 (`ld.global.nc.L1::no_allocate`, no reuse) so the index lines stay resident;
//   * U gathers are issued back to back before the first FMA;
//   * all loop arithmetic is 32-bit (entry index relative to the row start, one IMAD.WIDE per neighbour
//     address): the first v2 build spent 56 instructions per gather on 64-bit index math and was 67 %
//     issue-bound (profiles/r1_v2a_spmm_warp_kernel.txt);
//   * sub-group partial sums are combined with a fixed-order butterfly -> deterministic results.
// Measured dead ends (gpurun_out/tune*.log): an L2 prefetch stream running ahead of the
// gathers (prefetch.global.L2: no gain; cp.async.bulk.prefetch.L2: 40 % slower) and L2 cache-policy hints
// (evict_last on a fraction of the table, evict_first on the streams: 10-25 % slower).
__device__ __forceinline__ void prefetch_row_l2(const float* p, int bytes) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  if (bytes > 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32));
  if (bytes > 256) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 64));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 96));
  }
}

__device__ __forceinline__ float4 ld_gather_noalloc_f4(const char* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// (A CTA-staged variant that left as one TMA bulk store per peer, cp.async.bulk.global.shared::cta, was built in
// round 1 to test whether SM-issued peer stores were the reason a layer with exchange costs +0.2 ms at 8 GPUs:
// they are not — it measured 4-10 % slower (profiles/r1_bench_n8_fused_bulk.json) and was removed.)
// Gathered rows written earlier in the SAME kernel by other GPUs (the chain kernel below) are read with a weak
// ld.global (coherent at L2, the point where peer writes land) instead of the non-coherent path.
__device__ __forceinline__ float4 ld_gather_weak_f4(const char* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// The rows [r0, r0 + rows_per_warp) of one warp (body of spmm_warp_kernel and of every SpMM tile of the chain kernel).
template <int G, int U, bool HAS_VAL, bool TWO_TABLES, bool FULL, bool NC = true>
__device__ __forceinline__ void warp_rows(const b200gcn_spmm_args& a, const int64_t r0, const int64_t long_row,
                                          const int rows_per_warp, const int pf_edges) {
  constexpr int EPI = 32 / G;  // entries per gather instruction
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int lig = lane & (G - 1);
  const int sg = lane / G;
  const int cc = lig * 4;
  const bool col_ok = FULL || cc < a.dim;  // !FULL: lanes past the row end read column 0 and are discarded
  const int64_t r1 = min(a.n_rows, r0 + int64_t(rows_per_warp));
  const int nr = r0 < a.n_rows ? int(r1 - r0) : 0;
  const int64_t my_rp = a.rowptr[min(min(r0, a.n_rows) + lane, max(r1, min(r0, a.n_rows)))];  // <= 31 rows: one coalesced load
  // per-lane base pointers with the lane's column offset folded in
  const uint32_t ldb = uint32_t(a.ldx) * 4u;
  const int ccl = col_ok ? cc : 0;
  const char* xb = reinterpret_cast<const char*>(a.x) + ccl * 4;
  const int split = TWO_TABLES ? int(a.x_split) : 0;
  const char* xb2 = TWO_TABLES ? reinterpret_cast<const char*>(a.x2) + ccl * 4 - uint64_t(uint32_t(split)) * ldb : xb;
  // Prefetch stream over the warp's contiguous entry range [E0, E1): every 32 entries one COALESCED index
  // load (which also pulls the index lines into L1 for the broadcast loads that follow) and an L2
  // prefetch of the 32 neighbour rows, `pf_edges` entries ahead of the compute position.
  const int64_t E0 = __shfl_sync(kFull, my_rp, 0);
  const int etot = int(__shfl_sync(kFull, my_rp, nr) - E0);
  const int* __restrict__ col0 = a.col + E0;
  const float* __restrict__ val0 = HAS_VAL ? a.val + E0 : nullptr;
  const int row_bytes = a.dim * 4;
  int pf = 0;
  int cpf = (pf_edges > 0 && lane < etot) ? __ldg(col0 + lane) : -1;

  for (int rr = 0; rr < nr; ++rr) {
    const int64_t row = r0 + rr;
    const int64_t b = __shfl_sync(kFull, my_rp, rr), e = __shfl_sync(kFull, my_rp, rr + 1);
    if (e - b > long_row) continue;  // hub row: spmm_hub_kernel
    const int len = int(e - b);
    const int* __restrict__ colp = a.col + b;
    const float* __restrict__ valp = HAS_VAL ? a.val + b : nullptr;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // Index entries are fetched one batch AHEAD of the gathers that use them, and full batches run
    // without predicates: with either missing, ptxas schedules each FMA right behind its own gather
    // (load R8 -> fma R8 -> load R8 ...: one gather in flight; profiles/r1_v2b_sass_note.txt).
    int cn[U];
    float wn[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = sg + u * EPI;
      const bool live = j < len;
      cn[u] = live ? __ldg(colp + j) : -1;
      wn[u] = (HAS_VAL && live) ? __ldg(valp + j) : 1.0f;
    }
    const int rpos = int(b - E0);  // position of the row inside the warp's entry stream
    if (pf_edges > 0 && pf + 32 <= rpos) {  // the stream fell behind (hub rows were skipped): jump to this row
      pf = rpos;
      cpf = (pf + lane < etot) ? __ldg(col0 + pf + lane) : -1;
    }
    int i = 0;
    for (; i + U * EPI <= len; i += U * EPI) {  // full batches: every lane live, straight-line code
      if (pf_edges > 0 && pf < etot && pf < rpos + i + pf_edges) {  // warp-uniform
        if (cpf >= 0) {
          const char* prow = ((TWO_TABLES && cpf >= split) ? xb2 : xb) + uint64_t(uint32_t(cpf)) * ldb - ccl * 4;
          prefetch_row_l2(reinterpret_cast<const float*>(prow), row_bytes);
        }
        pf += 32;
        cpf = (pf + lane < etot) ? __ldg(col0 + pf + lane) : -1;
        if (HAS_VAL && lane == 0 && pf < etot) asm volatile("prefetch.global.L1 [%0];" ::"l"(val0 + pf));
      }
      int c[U];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { c[u] = cn[u]; w[u] = wn[u]; }
#pragma unroll
      for (int u = 0; u < U; ++u) {  // next batch's entries (may be partial)
        const int j = i + (U + u) * EPI + sg;
        const bool live = j < len;
        cn[u] = live ? __ldg(colp + j) : -1;
        wn[u] = (HAS_VAL && live) ? __ldg(valp + j) : 1.0f;
      }
      float4 xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const char* base = (TWO_TABLES && c[u] >= split) ? xb2 : xb;
        xv[u] = NC ? ld_gather_noalloc_f4(base + uint64_t(uint32_t(c[u])) * ldb) : ld_gather_weak_f4(base + uint64_t(uint32_t(c[u])) * ldb);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc.x = fmaf(w[u], xv[u].x, acc.x);
        acc.y = fmaf(w[u], xv[u].y, acc.y);
        acc.z = fmaf(w[u], xv[u].z, acc.z);
        acc.w = fmaf(w[u], xv[u].w, acc.w);
      }
    }
    if (i < len) {  // tail batch: predicated
      float4 xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const char* base = (TWO_TABLES && cn[u] >= split) ? xb2 : xb;
        if (cn[u] >= 0) xv[u] = NC ? ld_gather_noalloc_f4(base + uint64_t(uint32_t(cn[u])) * ldb) : ld_gather_weak_f4(base + uint64_t(uint32_t(cn[u])) * ldb);
        else { xv[u] = make_float4(0.f, 0.f, 0.f, 0.f); wn[u] = 0.f; }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc.x = fmaf(wn[u], xv[u].x, acc.x);
        acc.y = fmaf(wn[u], xv[u].y, acc.y);
        acc.z = fmaf(wn[u], xv[u].z, acc.z);
        acc.w = fmaf(wn[u], xv[u].w, acc.w);
      }
    }
#pragma unroll
    for (int o = 16; o >= G; o >>= 1) {  // combine the sub-groups (fixed order -> deterministic)
      acc.x += __shfl_xor_sync(kFull, acc.x, o);
      acc.y += __shfl_xor_sync(kFull, acc.y, o);
      acc.z += __shfl_xor_sync(kFull, acc.z, o);
      acc.w += __shfl_xor_sync(kFull, acc.w, o);
    }
    if (lane < G) {
      float4 accv[1] = {acc};
      finish_row<G, 1>(a, row, lig, G == 32 ? kFull : ((1u << (G & 31)) - 1u), accv);
    }
  }
}

template <int G, int U, bool HAS_VAL, bool TWO_TABLES, bool FULL>
__global__ void __launch_bounds__(kCta, (U >= 16 ? 2 : U >= 8 ? 3 : 5)) spmm_warp_kernel(const b200gcn_spmm_args a, int64_t long_row,
                                                         int rows_per_warp, int pf_edges) {
  const int64_t warp = (int64_t(blockIdx.x) * kCta + threadIdx.x) >> 5;
  const int64_t r0 = warp * rows_per_warp;
  if (r0 >= a.n_rows) return;
  warp_rows<G, U, HAS_VAL, TWO_TABLES, FULL>(a, r0, long_row, rows_per_warp, pf_edges);
  publish_fence(a);
}

// Hub rows (more than long_row entries): one CTA per hub row; the CTA's groups take interleaved
// chunks of the row, partial sums meet in shared memory and are added in group order.
template <int G, int V, bool HAS_VAL, bool TWO_TABLES>
__global__ void __launch_bounds__(kCta) spmm_hub_kernel(const b200gcn_spmm_args a,
                                                        const int64_t* __restrict__ hub_rows) {
  constexpr int kGroups = kCta / G;
  __shared__ float4 part[kGroups][V][G];
  const int lig = threadIdx.x & (G - 1);
  const int grp = threadIdx.x / G;
  const unsigned gm = group_mask<G>();
  const int64_t row = hub_rows[blockIdx.x];
  const int64_t beg = a.rowptr[row], end = a.rowptr[row + 1];
  const int64_t chunk = ((end - beg + kGroups - 1) / kGroups + G - 1) / G * G;
  float4 acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t b = min(end, beg + grp * chunk), e = min(end, b + chunk);
  gather_row<G, V, HAS_VAL, TWO_TABLES>(a, b, e, lig, gm, acc);
#pragma unroll
  for (int k = 0; k < V; ++k) part[grp][k][lig] = acc[k];
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float4 s = part[0][k][lig];
      for (int g = 1; g < kGroups; ++g) {
        const float4 t = part[g][k][lig];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      acc[k] = s;
    }
    finish_row<G, V>(a, row, lig, gm, acc);
    publish_fence(a);
  }
}

// Chunked hub path (b200gcn_spmm_hubs): a hub row with millions of entries (Zipf item popularity: the top item
// of a 100 M-interaction graph holds ~9 M) is cut into chunks of <= kHubChunk entries, one CTA per chunk writes a
// partial row to scratch, and a second kernel adds the partials of every hub in chunk order -> deterministic.
template <int G, int V, bool HAS_VAL, bool TWO_TABLES>
__global__ void __launch_bounds__(kCta) spmm_hub_chunk_kernel(const b200gcn_spmm_args a, const b200gcn_hub_plan hp) {
  constexpr int kGroups = kCta / G;
  __shared__ float4 part[kGroups][V][G];
  const int lig = threadIdx.x & (G - 1);
  const int grp = threadIdx.x / G;
  const unsigned gm = group_mask<G>();
  const int c = blockIdx.x;
  const int64_t beg = hp.chunk_beg[c], end = hp.chunk_end[c];
  const int64_t chunk = ((end - beg + kGroups - 1) / kGroups + G - 1) / G * G;
  float4 acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t b = min(end, beg + grp * chunk), e = min(end, b + chunk);
  gather_row<G, V, HAS_VAL, TWO_TABLES>(a, b, e, lig, gm, acc);
#pragma unroll
  for (int k = 0; k < V; ++k) part[grp][k][lig] = acc[k];
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float4 s = part[0][k][lig];
      for (int g = 1; g < kGroups; ++g) {
        const float4 t = part[g][k][lig];
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
      }
      const int cc = lig * 4 + k * G * 4;
      if (cc < a.dim) *reinterpret_cast<float4*>(hp.scratch + int64_t(c) * a.dim + cc) = s;
    }
  }
}

template <int G, int V>
__global__ void __launch_bounds__(kCta) spmm_hub_finish_kernel(const b200gcn_spmm_args a, const b200gcn_hub_plan hp) {
  const int lig = threadIdx.x & (G - 1);
  const unsigned gm = group_mask<G>();
  const int h = (blockIdx.x * kCta + threadIdx.x) / G;
  if (h >= hp.n_hubs) return;
  float4 acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = hp.hub_chunk_ptr[h]; c < hp.hub_chunk_ptr[h + 1]; ++c) {
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int cc = lig * 4 + k * G * 4;
      if (cc < a.dim) {
        const float4 t = *reinterpret_cast<const float4*>(hp.scratch + int64_t(c) * a.dim + cc);
        acc[k].x += t.x; acc[k].y += t.y; acc[k].z += t.z; acc[k].w += t.w;
      }
    }
  }
  finish_row<G, V>(a, hp.hub_rows[h], lig, gm, acc);
  publish_fence(a);
}

// Identity "propagation" (rowptr == NULL): p[r] = X[r].  Runs the same epilogue on rows that are already
// known — used to publish a rank's layer-0 rows into every peer's gather table over NVLink, and to derive the
// perturbed SimGCL views of a shared first layer without repeating the SpMM.
template <int G, int V>
__global__ void __launch_bounds__(kCta) rows_identity_kernel(const b200gcn_spmm_args a) {
  const int lig = threadIdx.x & (G - 1);
  const unsigned gm = group_mask<G>();
  const int64_t row = (int64_t(blockIdx.x) * kCta + threadIdx.x) / G;
  if (row >= a.n_rows) return;
  const float* src = (a.x2 != nullptr && row >= a.x_split) ? a.x2 + (row - a.x_split) * a.ldx : a.x + row * a.ldx;
  float4 acc[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int cc = lig * 4 + k * G * 4;
    acc[k] = cc < a.dim ? *reinterpret_cast<const float4*>(src + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  finish_row<G, V>(a, row, lig, gm, acc);
  publish_fence(a);
}

// ---------------------------------------------------------------------------------------------------
// Phase chain (b200gcn_spmm_chain): the K-layer row-sharded propagation as ONE persistent cooperative kernel.
// See include/b200gcn.h.  Tiles are handed out in phase order from an atomic counter; cross-GPU ordering is
// flag words in symmetric memory: writer = last CTA of a rank to leave a phase (system fence, then one 4-byte
// store per peer), reader = thread 0 of a CTA (ld.acquire.sys spin) followed by a CTA barrier.
constexpr int kChainIdRows = 256;  // rows per identity (publish) tile

struct ChainParams {
  b200gcn_spmm_args ph[B200GCN_CHAIN_MAX_PHASES];
  int32_t tile_end[B200GCN_CHAIN_MAX_PHASES];  // exclusive prefix of tiles (a merged pair shares its end)
  int32_t n_tiles[B200GCN_CHAIN_MAX_PHASES];   // tiles of the phase itself
  int32_t period[B200GCN_CHAIN_MAX_PHASES];    // merged pair: every period-th tile of the range belongs to the first phase
  b200gcn_chain_sync sync;
  int32_t n_ph;
  int32_t rpw;
  int32_t pf;
};

// A peer that never arrives (its process died, or the ranks disagree on the phase list) must not leave this GPU
// spinning forever: after kChainTimeoutNs (2 minutes) the kernel traps and the launch surfaces as a CUDA error on the host.
constexpr unsigned long long kChainTimeoutNs = 120ull * 1000 * 1000 * 1000;

__device__ __forceinline__ void chain_wait(const b200gcn_chain_sync& s, int phase, uint32_t epoch) {
  unsigned long long t0 = 0;
  for (int q = 0; q < s.n_ranks; ++q) {
    const uint32_t* f = s.flags + phase * B200GCN_CHAIN_MAX_RANKS + q;
    uint32_t v;
    for (unsigned spins = 0;; ++spins) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (int32_t(v - epoch) >= 0) break;
      __nanosleep(64);
      if ((spins & 0xfffu) == 0xfffu) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t0 == 0) t0 = t;
        else if (t - t0 > kChainTimeoutNs) __trap();
      }
    }
  }
}

template <int G>
__device__ __forceinline__ void chain_identity_tile(const b200gcn_spmm_args& a, int64_t row_lo) {
  const int lig = threadIdx.x & (G - 1);
  const unsigned gm = group_mask<G>();
  const int64_t row_hi = min(a.n_rows, row_lo + int64_t(kChainIdRows));
  for (int64_t row = row_lo + threadIdx.x / G; row < row_hi; row += kCta / G) {
    const int cc = lig * 4;
    float4 acc[1];
    acc[0] = cc < a.dim ? *reinterpret_cast<const float4*>(a.x + row * a.ldx + cc) : make_float4(0.f, 0.f, 0.f, 0.f);
    finish_row<G, 1>(a, row, lig, gm, acc);
  }
}

template <int G, int U, bool HAS_VAL, bool FULL>
__global__ void __launch_bounds__(kCta, 3) spmm_chain_kernel(const __grid_constant__ ChainParams P) {
  __shared__ int s_tile;
  const b200gcn_chain_sync& S = P.sync;
  int cur = 0;             // phases [0, cur) already left by this CTA
  unsigned ready = 0;      // (thread 0) phases known complete on all ranks
  unsigned ready_loc = 0;  // (thread 0) phases known complete on this rank
  if (threadIdx.x == 0 && S.start_wait_phase >= 0 && S.epoch > 1) chain_wait(S, S.start_wait_phase, S.epoch - 1);
  int next_tile = -1;      // (thread 0) tile index taken one tile ahead: the atomic's round trip hides under the rows
  for (;;) {
    __syncthreads();  // s_tile free; (first pass) the start wait is over
    if (threadIdx.x == 0) {
      s_tile = next_tile >= 0 ? next_tile : atomicAdd(&S.scratch[0], 1);
      next_tile = atomicAdd(&S.scratch[0], 1);   // consumed on the next pass (indices stay ascending per CTA)
    }
    __syncthreads();
    const int tile = s_tile;
    if (tile == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      reinterpret_cast<unsigned long long*>(S.scratch + 16)[0] = t;
    }
    int p = cur;
    while (p < P.n_ph && tile >= P.tile_end[p]) ++p;   // first phase of the group this tile belongs to
    if (p > cur) {  // leaving phases cur .. p-1: their published rows must be visible system-wide first
      bool published = false;
      for (int q = cur; q < p; ++q) published |= P.ph[q].n_peers > 0 || P.ph[q].y_mc != nullptr;
      // release pattern: every thread orders its own NVLink stores before the CTA's arrival (fence.acq_rel is enough:
      // the flag store below is the only thing a peer synchronises on); phases that stored nothing remote skip it
      if (published) asm volatile("fence.acq_rel.sys;" ::: "memory");
      else __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int q = cur; q < p; ++q) {
          const int old = atomicAdd(&S.scratch[1 + q], 1);
          if (old == int(gridDim.x) - 1) {  // last CTA of this rank out of phase q: tell every rank
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            reinterpret_cast<unsigned long long*>(S.scratch + 16)[1 + q] = t;
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            for (int r = 0; r < S.n_ranks; ++r) {
              uint32_t* f = S.flags_peers[r] + q * B200GCN_CHAIN_MAX_RANKS + S.rank;
              asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(S.epoch) : "memory");
            }
          }
        }
      }
      cur = p;
    }
    if (p >= P.n_ph) break;
    int t_in_phase = tile - (p == 0 ? 0 : P.tile_end[p - 1]);
    if (S.merge_next[p]) {   // interleaved pair (p, p + 1): positions 0, period, 2 period, .. are the tiles of p
      const int j = t_in_phase, per = P.period[p], np = P.n_tiles[p];
      if (j % per == 0 && j / per < np) {
        t_in_phase = j / per;
      } else {
        t_in_phase = j - min((j + per - 1) / per, np);
        ++p;   // (cur stays at the first phase of the pair: both are left together)
      }
    }
    const int w = S.wait_phase[p], wl = S.wait_local[p];
    if (w >= 0 || wl >= 0) {
      if (threadIdx.x == 0) {
        if (w >= 0 && !((ready >> w) & 1u)) {
          chain_wait(S, w, S.epoch);
          ready |= 1u << w;
        }
        if (wl >= 0 && !((ready_loc >> wl) & 1u)) {  // every CTA of this rank has left phase wl
          int v;
          for (;;) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(S.scratch + 1 + wl) : "memory");
            if (v >= int(gridDim.x)) break;
            __nanosleep(32);
          }
          ready_loc |= 1u << wl;
        }
      }
      __syncthreads();
    }
    const b200gcn_spmm_args& a = P.ph[p];
    const int t = t_in_phase;
    if (a.rowptr == nullptr) {
      chain_identity_tile<G>(a, int64_t(t) * kChainIdRows);
    } else {
      const int64_t r0 = (int64_t(t) * (kCta / 32) + (threadIdx.x >> 5)) * P.rpw;
      if (r0 < a.n_rows) warp_rows<G, U, HAS_VAL, false, FULL, false>(a, r0, INT64_MAX, P.rpw, P.pf);
    }
  }
}

// Collect rows with more than long_row entries (ascending order is not required).
__global__ void find_hubs(const int64_t* __restrict__ rowptr, int64_t n_rows, int64_t long_row,
                          int64_t* __restrict__ hub_rows, int* __restrict__ n_hubs, int cap) {
  for (int64_t r = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; r < n_rows;
       r += int64_t(gridDim.x) * blockDim.x) {
    if (rowptr[r + 1] - rowptr[r] > long_row) {
      int p = atomicAdd(n_hubs, 1);
      if (p < cap) hub_rows[p] = r;
    }
  }
}

}  // namespace
}  // namespace b200gcn

using namespace b200gcn;

namespace {

template <int G, int V>
int launch(const b200gcn_spmm_args& a, int64_t long_row, const int64_t* hubs, int n_hubs, cudaStream_t st) {
  const bool has_val = a.val != nullptr;
  const bool two = a.x2 != nullptr;
  const int64_t rows_per_cta = kCta / G;
  const int64_t grid = (a.n_rows + rows_per_cta - 1) / rows_per_cta;
  if (grid > 0x7fffffffLL) {
    set_error("n_rows too large for one launch");
    return B200GCN_ERR_INVALID;
  }
#define B200_LAUNCH(HV, TT)                                                                         \
  do {                                                                                              \
    if (n_hubs == 0)                                                                                \
      spmm_rows_kernel<G, V, HV, TT><<<unsigned(grid), kCta, 0, st>>>(a, long_row);                 \
    else                                                                                            \
      spmm_hub_kernel<G, V, HV, TT><<<unsigned(n_hubs), kCta, 0, st>>>(a, hubs);                    \
  } while (0)
  if (has_val && two) B200_LAUNCH(true, true);
  else if (has_val) B200_LAUNCH(true, false);
  else if (two) B200_LAUNCH(false, true);
  else B200_LAUNCH(false, false);
#undef B200_LAUNCH
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}

// flags (tuning word of b200gcn_spmm_args): bits 0-3 kernel (0 auto, 1 = v1 row-group, 2 = v2 warp-row),
// bits 4-7 gathers in flight per lane for v2 (0 or 8 -> 8; 4 -> 4; 1 -> 16), bits 8-15 L2 prefetch distance in
// entries / 8 (0 -> 16 entries; 255 -> prefetch stream off), bits 16-23 rows per warp (0 -> 4, or 2 at D > 64).
template <int G>
int launch_v2(const b200gcn_spmm_args& a, int64_t long_row, cudaStream_t st) {
  const int fl = a.flags;
  const int U = ((fl >> 4) & 15) == 4 ? 4 : ((fl >> 4) & 15) == 1 ? 16 : 8;  // nibble 1 selects U = 16
  int rpw = (fl >> 16) & 255;
  if (rpw == 0) rpw = G == 32 ? 2 : 4;  // sweeps: gpurun_out/tune_v2e_{64,128}.log
  if (rpw > 31) rpw = 31;
  int pf = ((fl >> 8) & 255) * 8;
  if (pf == 0) pf = 16;
  if (pf == 255 * 8) pf = 0;
  const bool has_val = a.val != nullptr, two = a.x2 != nullptr, full = a.dim == G * 4;
  const int64_t n_warps = (a.n_rows + rpw - 1) / rpw;
  const int64_t grid = (n_warps + kCta / 32 - 1) / (kCta / 32);
  if (grid > 0x7fffffffLL) {
    set_error("n_rows too large for one launch");
    return B200GCN_ERR_INVALID;
  }
  if (a.ldx * 4 > 0xffffffffLL || (two && a.x_split > 0x7fffffffLL)) {
    set_error("ldx / x_split too large for the 32-bit fast path");
    return B200GCN_ERR_INVALID;
  }
#define B200_V2(UU, HV, TT, FL) spmm_warp_kernel<G, UU, HV, TT, FL><<<unsigned(grid), kCta, 0, st>>>(a, long_row, rpw, pf)
#define B200_V2F(UU, HV, TT) do { if (full) B200_V2(UU, HV, TT, true); else B200_V2(UU, HV, TT, false); } while (0)
#define B200_V2U(HV, TT) do { if (U == 8) B200_V2F(8, HV, TT); else if (U == 16) B200_V2F(16, HV, TT); else B200_V2F(4, HV, TT); } while (0)
  if (has_val && two) B200_V2U(true, true);
  else if (has_val) B200_V2U(true, false);
  else if (two) B200_V2U(false, true);
  else B200_V2U(false, false);
#undef B200_V2U
#undef B200_V2F
#undef B200_V2
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}

template <int G, int V>
int launch_identity(const b200gcn_spmm_args& a, cudaStream_t st) {
  const int64_t rows_per_cta = kCta / G;
  const int64_t grid = (a.n_rows + rows_per_cta - 1) / rows_per_cta;
  if (grid > 0x7fffffffLL) {
    set_error("n_rows too large for one launch");
    return B200GCN_ERR_INVALID;
  }
  rows_identity_kernel<G, V><<<unsigned(grid), kCta, 0, st>>>(a);
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}

int dispatch(const b200gcn_spmm_args& a, int64_t long_row, const int64_t* hubs, int n_hubs, cudaStream_t st) {
  const int D = a.dim;
  if (a.rowptr == nullptr) {
    if (hubs != nullptr) return B200GCN_OK;
    if (D <= 32) return launch_identity<8, 1>(a, st);
    if (D <= 64) return launch_identity<16, 1>(a, st);
    if (D <= 128) return launch_identity<32, 1>(a, st);
    if (D <= 256) return launch_identity<32, 2>(a, st);
    return launch_identity<32, 4>(a, st);
  }
  const int kern = a.flags & 15;
  if (n_hubs == 0 && D <= 128 && kern != 1) {
    if (D <= 32) return launch_v2<8>(a, long_row, st);
    if (D <= 64) return launch_v2<16>(a, long_row, st);
    return launch_v2<32>(a, long_row, st);
  }
  if (D <= 32) return launch<8, 1>(a, long_row, hubs, n_hubs, st);
  if (D <= 64) return launch<16, 1>(a, long_row, hubs, n_hubs, st);
  if (D <= 128) return launch<32, 1>(a, long_row, hubs, n_hubs, st);
  if (D <= 256) return launch<32, 2>(a, long_row, hubs, n_hubs, st);
  return launch<32, 4>(a, long_row, hubs, n_hubs, st);
}

}  // namespace

// Hub-row plan: b200gcn_spmm looks for rows above kLongRow on every call only when the caller has
// not supplied a plan; the Python host caches the plan per graph through b200gcn_plan_hubs.
extern "C" int b200gcn_plan_hubs(const int64_t* rowptr, int64_t n_rows, int64_t long_row,
                                 int64_t* hub_rows, int32_t cap, int32_t* h_count, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(rowptr && h_count && long_row > 0 && cap >= 0, "bad arguments");
  B200_CHECK_ARG(cap == 0 || hub_rows != nullptr, "hub_rows is NULL");
  int* d_cnt = nullptr;
  B200_CHECK_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_cnt), sizeof(int), st));
  B200_CHECK_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int), st));
  if (n_rows > 0) {
    int64_t g = (n_rows + 255) / 256;
    if (g > 65535 * 16) g = 65535 * 16;
    find_hubs<<<unsigned(g), 256, 0, st>>>(rowptr, n_rows, long_row, hub_rows, d_cnt, cap);
    B200_CHECK_LAUNCH();
  }
  int h = 0;
  B200_CHECK_CUDA(cudaMemcpyAsync(&h, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
  B200_CHECK_CUDA(cudaStreamSynchronize(st));
  B200_CHECK_CUDA(cudaFreeAsync(d_cnt, st));
  *h_count = h;
  return B200GCN_OK;
}

static int validate(const b200gcn_spmm_args* a) {
  B200_CHECK_ARG(a != nullptr, "args is NULL");
  B200_CHECK_ARG(a->n_rows >= 0, "n_rows < 0");
  B200_CHECK_ARG(a->dim > 0 && a->dim % 4 == 0 && a->dim <= 512, "dim=%d must be a multiple of 4 in [4, 512]",
                 a->dim);
  if (a->n_rows == 0) return B200GCN_OK;
  // col may be NULL only for a graph without entries (it is never dereferenced then)
  B200_CHECK_ARG(a->x, "x is NULL");  // rowptr == NULL selects the identity mode (p = x)
  B200_CHECK_ARG(a->y || a->acc_out || a->n_peers > 0 || a->y_mc, "y, acc_out and the peer tables are all NULL: nothing to write");
  B200_CHECK_ARG(a->n_peers >= 0 && a->n_peers <= 64 && (a->n_peers == 0 || a->y_peers), "n_peers / y_peers");
  B200_CHECK_ARG((a->n_peers == 0 && !a->y_mc) || (a->ld_peer % 4 == 0 && a->ld_peer >= a->dim && a->y_peer_row0 >= 0),
                 "ld_peer / y_peer_row0");
  B200_CHECK_ARG(!a->y_mc || aligned16(a->y_mc), "y_mc must be 16-byte aligned");
  B200_CHECK_ARG(!a->peer_need || (a->n_peers > 0 && a->n_peers <= 32 && !a->y_mc), "peer_need needs the y_peers path, <= 32 ranks");
  B200_CHECK_ARG(aligned16(a->x) && a->ldx % 4 == 0 && a->ldx >= a->dim, "x must be 16-byte aligned, ldx %% 4 == 0, ldx >= dim");
  B200_CHECK_ARG(!a->x2 || aligned16(a->x2), "x2 must be 16-byte aligned");
  B200_CHECK_ARG(!a->y || (aligned16(a->y) && a->ldy % 4 == 0 && a->ldy >= a->dim), "y alignment / ldy");
  B200_CHECK_ARG(!a->noise || (aligned16(a->noise) && a->ldn % 4 == 0 && a->ldn >= a->dim), "noise alignment / ldn");
  B200_CHECK_ARG(!a->acc_in || (aligned16(a->acc_in) && a->ld_acc_in % 4 == 0 && a->ld_acc_in >= a->dim), "acc_in alignment / ld");
  B200_CHECK_ARG(!a->acc_in2 || (a->acc_in && aligned16(a->acc_in2)), "acc_in2 needs acc_in and 16-byte alignment");
  B200_CHECK_ARG(a->n_acc_extra >= 0 && a->n_acc_extra <= 3, "n_acc_extra outside [0,3]");
  for (int e = 0; e < a->n_acc_extra; ++e)
    B200_CHECK_ARG(a->acc_out && a->acc_extra[e] && aligned16(a->acc_extra[e]) && a->ld_acc_extra % 4 == 0 &&
                       a->ld_acc_extra >= a->dim, "acc_extra needs acc_out, 16-byte alignment, ld_acc_extra");
  B200_CHECK_ARG(!a->acc_out || (aligned16(a->acc_out) && a->ld_acc_out % 4 == 0 && a->ld_acc_out >= a->dim), "acc_out alignment / ld");
  return B200GCN_OK;
}

extern "C" int b200gcn_spmm_planned(const b200gcn_spmm_args* args, int64_t long_row,
                                    const int64_t* hub_rows, int32_t n_hubs, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = validate(args);
  if (rc) return rc;
  if (args->n_rows == 0) return B200GCN_OK;
  B200_CHECK_ARG(long_row > 0 && n_hubs >= 0 && (n_hubs == 0 || hub_rows), "bad hub plan");
  rc = dispatch(*args, long_row, nullptr, 0, st);
  if (rc) return rc;
  if (n_hubs > 0) rc = dispatch(*args, long_row, hub_rows, n_hubs, st);
  return rc;
}


namespace {
template <int G, int V>
int launch_hubs(const b200gcn_spmm_args& a, const b200gcn_hub_plan& hp, cudaStream_t st) {
  const bool has_val = a.val != nullptr, two = a.x2 != nullptr;
#define B200_HC(HV, TT) spmm_hub_chunk_kernel<G, V, HV, TT><<<unsigned(hp.n_chunks), kCta, 0, st>>>(a, hp)
  if (has_val && two) B200_HC(true, true);
  else if (has_val) B200_HC(true, false);
  else if (two) B200_HC(false, true);
  else B200_HC(false, false);
#undef B200_HC
  B200_CHECK_LAUNCH();
  const int groups_per_cta = kCta / G;
  spmm_hub_finish_kernel<G, V><<<unsigned((hp.n_hubs + groups_per_cta - 1) / groups_per_cta), kCta, 0, st>>>(a, hp);
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}
}  // namespace

extern "C" int b200gcn_spmm_hubs(const b200gcn_spmm_args* args, const b200gcn_hub_plan* plan, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = validate(args);
  if (rc) return rc;
  B200_CHECK_ARG(plan != nullptr && plan->n_hubs >= 0 && plan->n_chunks >= plan->n_hubs, "bad hub plan");
  if (plan->n_hubs == 0 || args->n_rows == 0) return B200GCN_OK;
  B200_CHECK_ARG(args->rowptr && plan->hub_rows && plan->hub_chunk_ptr && plan->chunk_beg && plan->chunk_end &&
                     plan->scratch && aligned16(plan->scratch),
                 "hub plan arrays / scratch are NULL or misaligned");
  const int D = args->dim;
  if (D <= 32) return launch_hubs<8, 1>(*args, *plan, st);
  if (D <= 64) return launch_hubs<16, 1>(*args, *plan, st);
  if (D <= 128) return launch_hubs<32, 1>(*args, *plan, st);
  if (D <= 256) return launch_hubs<32, 2>(*args, *plan, st);
  return launch_hubs<32, 4>(*args, *plan, st);
}

extern "C" int b200gcn_spmm(const b200gcn_spmm_args* args, void* stream) {
  // Plan-free entry: every row is taken by the row kernel whatever its length (correct for any
  // graph; the hub plan only matters for the speed of heavily skewed graphs).
  return b200gcn_spmm_planned(args, INT64_MAX, nullptr, 0, stream);
}


// ---------------------------------------------------------------------------------------------------
namespace {
template <int G>
int launch_chain(ChainParams& P, bool has_val, bool full, cudaStream_t st) {
  void* kern = nullptr;
  if (has_val) kern = full ? (void*)spmm_chain_kernel<G, 8, true, true> : (void*)spmm_chain_kernel<G, 8, true, false>;
  else kern = full ? (void*)spmm_chain_kernel<G, 8, false, true> : (void*)spmm_chain_kernel<G, 8, false, false>;
  int dev = 0, sms = 0, per_sm = 0;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  B200_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kCta, 0));
  B200_CHECK_ARG(per_sm > 0, "chain kernel does not fit on an SM");
  int64_t grid = int64_t(sms) * per_sm;
  const int64_t tiles = P.tile_end[P.n_ph - 1];
  if (grid > tiles) grid = tiles > 0 ? tiles : 1;
  void* args[] = {&P};
  B200_CHECK_CUDA(cudaLaunchCooperativeKernel(kern, dim3(unsigned(grid)), dim3(kCta), args, 0, st));
  return B200GCN_OK;
}
}  // namespace

extern "C" int b200gcn_spmm_chain(const b200gcn_spmm_args* phases, int32_t n_phases, const b200gcn_chain_sync* sync,
                                  void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(phases && sync && n_phases >= 1 && n_phases <= B200GCN_CHAIN_MAX_PHASES, "n_phases outside [1, %d]",
                 B200GCN_CHAIN_MAX_PHASES);
  B200_CHECK_ARG(sync->n_ranks >= 1 && sync->n_ranks <= B200GCN_CHAIN_MAX_RANKS && sync->rank >= 0 &&
                     sync->rank < sync->n_ranks && sync->epoch >= 1 && sync->flags && sync->flags_peers && sync->scratch,
                 "bad chain sync block");
  // (the start wait names the LAST phase of the previous call, which may have had more phases than this one)
  B200_CHECK_ARG(sync->start_wait_phase < B200GCN_CHAIN_MAX_PHASES, "start_wait_phase out of range");
  ChainParams P;
  memset(&P, 0, sizeof(P));
  P.sync = *sync;
  P.n_ph = n_phases;
  const int D = phases[0].dim;
  const bool has_val = phases[0].rowptr ? phases[0].val != nullptr : true;
  bool any_val_known = false, hv = true;
  int64_t tiles = 0;
  const int G = D <= 32 ? 8 : D <= 64 ? 16 : 32;
  P.rpw = G == 32 ? 2 : 4;
  P.pf = 16;
  for (int p = 0; p < n_phases; ++p) {
    int rc = validate(&phases[p]);
    if (rc) return rc;
    const b200gcn_spmm_args& a = phases[p];
    B200_CHECK_ARG(a.dim == D && D <= 128, "chain phases must share one dim <= 128");
    B200_CHECK_ARG(a.x2 == nullptr && a.acc_in2 == nullptr && a.eps == 0.f, "chain phases take single tables and no noise");
    B200_CHECK_ARG(sync->wait_phase[p] < p && sync->wait_local[p] < p, "wait_phase/wait_local[%d] must name an earlier phase", p);
    B200_CHECK_ARG(a.ldx * 4 <= 0xffffffffLL, "ldx too large for the 32-bit fast path");
    if (a.rowptr) {
      const bool v = a.val != nullptr;
      B200_CHECK_ARG(!any_val_known || v == hv, "chain phases must agree on val");
      any_val_known = true;
      hv = v;
      const int64_t rows_per_tile = int64_t(kCta / 32) * P.rpw;
      tiles += (a.n_rows + rows_per_tile - 1) / rows_per_tile;
    } else {
      tiles += (a.n_rows + kChainIdRows - 1) / kChainIdRows;
    }
    B200_CHECK_ARG(tiles < 0x7fffffffLL, "too many tiles");
    P.ph[p] = a;
    P.tile_end[p] = int32_t(tiles);
    P.n_tiles[p] = int32_t(tiles - (p == 0 ? 0 : P.tile_end[p - 1]));
    P.period[p] = 1;
  }
  for (int p = 0; p < n_phases; ++p) {
    if (!sync->merge_next[p]) continue;
    B200_CHECK_ARG(p + 1 < n_phases && !sync->merge_next[p + 1] && (p == 0 || !sync->merge_next[p - 1]),
                   "merge_next[%d]: pairs only, and a successor must exist", p);
    B200_CHECK_ARG(sync->wait_phase[p + 1] < p && sync->wait_local[p + 1] < p, "a merged pair cannot depend on itself");
    if (P.n_tiles[p] == 0 || P.n_tiles[p + 1] == 0) { P.sync.merge_next[p] = 0; continue; }
    const int total = P.n_tiles[p] + P.n_tiles[p + 1];
    P.period[p] = total / P.n_tiles[p];     // (n_tiles[p] - 1) * period < total: every tile of p has a position
    P.tile_end[p] = P.tile_end[p + 1];      // one shared range
  }
  (void)has_val;
  B200_CHECK_CUDA(cudaMemsetAsync(sync->scratch, 0, B200GCN_CHAIN_SCRATCH_BYTES, st));
  const bool full = D == G * 4;
  if (G == 8) return launch_chain<8>(P, hv, full, st);
  if (G == 16) return launch_chain<16>(P, hv, full, st);
  return launch_chain<32>(P, hv, full, st);
}
