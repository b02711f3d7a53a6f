// Backward of the NGCF layer tail (the autograd of layers.py:56-58 + ngcf.py:96-98 that `loss.backward()` runs in
// the reference), d_in = d_out = 64, as ONE pass over the node rows on the tcgen05 tensor cores:
//   forward:  t = (p + x) W1^T + b1 + (p * x) W2^T + b2 ;  z = leaky_relu(t) * keep_scale ;  out = z / max(||z||, eps)
//   given g = dL/d out, t and the dropout mask:
//     row-local    g_z = (g - z <z, g> / ||z||^2) / ||z||      (g / eps where ||z|| <= eps;  g itself without normalise)
//                  g_t = g_z * leaky'(t) * keep_scale
//     contraction  [g_a | g_m] = g_t (128 rows x K = 64) . [W1 | W2] (K = 64 x N = 128)       tcgen05.mma.kind::tf32,
//                  error-compensated split (four TF32 products per fp32 product), accumulator in TMEM
//     row-local    g_p = g_a + g_m * x ;  g_x = g_a + g_m * p
//   also written: g_t (for the bias gradient and the weight gradient) and am = [p + x | p * x], so that the weight
//   gradient [g_W1 | g_W2] = g_t^T am is ONE plain library GEMM ([64, n] x [n, 128]) instead of two GEMMs over freshly
//   materialised temporaries.
// Same pipeline as the forward kernel (bignn_tail_tc.cu): 512 threads, operand staged in the canonical K-major
// no-swizzle layout, MMAs issued warp-uniformly by an elected lane, two TMEM accumulators so that the epilogue of tile
// i-1 runs under the MMAs of tile i.
#include "common.cuh"

namespace b200gcn {
namespace {

constexpr int kBM = 128;                  // rows per tile
constexpr int kBD = 64;                   // d_in = d_out
constexpr int kBN = 2 * kBD;              // [g_a | g_m]
constexpr int kBChunks = kBD / 4;         // 16-byte K chunks per row (K = 64)
constexpr int kBThreads = 512;
constexpr int kBRowsPerPass = kBThreads / 16;   // 32
constexpr int kBIters = kBM / kBRowsPerPass;    // 4
constexpr uint32_t kBSBO = 128;
constexpr uint32_t kBLboA = kBM * 16 + 16;      // 2064: conflict-free 16-lanes-per-row stores
constexpr uint32_t kBLboB = kBN * 16;           // 2048 (staged once)
constexpr uint32_t kBBytesA = kBChunks * kBLboA;   // 33,024 per part
constexpr uint32_t kBBytesB = kBChunks * kBLboB;   // 32,768 per part
constexpr uint32_t kBOffAhi = 0, kBOffAlo = kBBytesA, kBOffBhi = 2 * kBBytesA, kBOffBlo = 2 * kBBytesA + kBBytesB;
constexpr uint32_t kBOffStage = 2 * kBBytesA + 2 * kBBytesB;   // [128][128] fp32, 16-byte slots XOR-swizzled by row & 7
constexpr uint32_t kBOffMisc = kBOffStage + kBM * kBN * 4;
constexpr uint32_t kBSmem = kBOffMisc + 32;
static_assert(kBSmem <= 227 * 1024, "shared memory budget");
constexpr uint32_t kBTmemCols = 2 * kBN;                        // two 128 x 128 fp32 accumulators

struct BwdArgs {
  const float* p; int64_t ldp; const float* x; int64_t ldx;
  const float* w1; const float* w2;
  const float* t; int64_t ld_t;            // pre-activation saved by the forward kernel
  const uint8_t* keep; float keep_scale;
  float slope; int normalize;
  const float* g; int64_t ld_g;            // dL/d out
  int64_t n;
  float* g_p; float* g_x; float* g_t; float* am;   // [n,64], [n,64], [n,64], [n,128] contiguous
};

__device__ __forceinline__ uint32_t b_smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float b_tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }
__device__ __forceinline__ uint64_t b_desc(uint32_t smem_addr, uint32_t lbo) {
  return uint64_t((smem_addr & 0x3ffffu) >> 4) | (uint64_t(lbo >> 4) << 16) | (uint64_t(kBSBO >> 4) << 32) |
         (uint64_t(1) << 46);
}
// D = F32, A = B = TF32, K-major both, N = 128, M = 128
constexpr uint32_t kBIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kBN >> 3) << 17) | (uint32_t(kBM >> 4) << 24);

__device__ __forceinline__ void b_umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kBIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void b_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ bool b_elect_one() {
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
__device__ __forceinline__ float b_sum16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o, 16);
  return v;
}

// B operand: N = 128 rows (input feature i of W1, then of W2) x K = 64 (output feature j): B[n][k] = W[k][n]
__device__ __forceinline__ void b_stage_weights(const BwdArgs& a, char* smem) {
  for (int idx = threadIdx.x; idx < kBN * kBChunks; idx += kBThreads) {
    const int n = idx / kBChunks, c = idx % kBChunks;
    const float* w = n < kBD ? a.w1 + n : a.w2 + (n - kBD);
    const float4 v = make_float4(w[(4 * c + 0) * kBD], w[(4 * c + 1) * kBD], w[(4 * c + 2) * kBD], w[(4 * c + 3) * kBD]);
    const float4 h = make_float4(b_tf32_hi(v.x), b_tf32_hi(v.y), b_tf32_hi(v.z), b_tf32_hi(v.w));
    const uint32_t off = c * kBLboB + (n >> 3) * kBSBO + (n & 7) * 16;
    *reinterpret_cast<float4*>(smem + kBOffBhi + off) = h;
    *reinterpret_cast<float4*>(smem + kBOffBlo + off) = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
  }
}

__global__ void __launch_bounds__(kBThreads, 1) bignn_tail_bwd_kernel(const BwdArgs a) {
  extern __shared__ __align__(1024) char smem[];
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + kBOffMisc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kBOffMisc + 16);
  char* stage = smem + kBOffStage;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (a.n + kBM - 1) / kBM;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(b_smem_u32(tmem_slot)),
                 "r"(kBTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b_smem_u32(mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  b_stage_weights(a, smem);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int my_chunk = tid & 15;              // 4 columns (j or i) of a row
  const int my_row0 = tid >> 4;               // rows my_row0 + 32 i
  const int q = warp & 3, h = warp >> 2;      // accumulator read: TMEM lanes 32 q .., columns 32 h ..
  float4 gv[kBIters], tv[kBIters];
  uchar4 kv[kBIters];

  auto load_tile = [&](int64_t tile) {
    const int64_t r0 = tile * kBM;
#pragma unroll
    for (int i = 0; i < kBIters; ++i) {
      const int64_t row = r0 + my_row0 + kBRowsPerPass * i;
      kv[i] = make_uchar4(1, 1, 1, 1);
      if (row < a.n) {
        gv[i] = ld_gather_f4(a.g + row * a.ld_g + my_chunk * 4);
        tv[i] = ld_gather_f4(a.t + row * a.ld_t + my_chunk * 4);
        if (a.keep != nullptr) kv[i] = *reinterpret_cast<const uchar4*>(a.keep + row * int64_t(kBD) + my_chunk * 4);
      } else {
        gv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        tv[i] = gv[i];
      }
    }
  };

  // g_a, g_m of `tile` (TMEM buffer `buf`) -> g_p, g_x, and the am rows for the weight gradient
  auto epilogue = [&](int64_t tile, uint32_t buf) {
    const int64_t r0 = tile * kBM;
    {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * uint32_t(kBN) + uint32_t(h * 32);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
            "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
            "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int r = q * 32 + lane;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int c4 = (h * 32 + j) >> 2;     // 16-byte slot 0..31 of the 128-float row
        *reinterpret_cast<float4*>(stage + r * (kBN * 4) + ((c4 ^ (r & 7)) << 4)) =
            make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                        __uint_as_float(v[j + 3]));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kBIters; ++i) {
      const int r = my_row0 + kBRowsPerPass * i;
      const int64_t row = r0 + r;
      if (row >= a.n) continue;
      const float4 ga = *reinterpret_cast<const float4*>(stage + r * (kBN * 4) + ((my_chunk ^ (r & 7)) << 4));
      const float4 gm = *reinterpret_cast<const float4*>(stage + r * (kBN * 4) + (((my_chunk + 16) ^ (r & 7)) << 4));
      const float4 pv = ld_gather_f4(a.p + row * a.ldp + my_chunk * 4);
      const float4 xv = ld_gather_f4(a.x + row * a.ldx + my_chunk * 4);
      st_stream_f4(a.g_p + row * kBD + my_chunk * 4,
                   make_float4(fmaf(gm.x, xv.x, ga.x), fmaf(gm.y, xv.y, ga.y), fmaf(gm.z, xv.z, ga.z), fmaf(gm.w, xv.w, ga.w)));
      st_stream_f4(a.g_x + row * kBD + my_chunk * 4,
                   make_float4(fmaf(gm.x, pv.x, ga.x), fmaf(gm.y, pv.y, ga.y), fmaf(gm.z, pv.z, ga.z), fmaf(gm.w, pv.w, ga.w)));
      if (a.am != nullptr) {
        st_stream_f4(a.am + row * kBN + my_chunk * 4, make_float4(pv.x + xv.x, pv.y + xv.y, pv.z + xv.z, pv.w + xv.w));
        st_stream_f4(a.am + row * kBN + kBD + my_chunk * 4, make_float4(pv.x * xv.x, pv.y * xv.y, pv.z * xv.z, pv.w * xv.w));
      }
    }
    // (measured: without this barrier rows of the last pass are intermittently wrong when a CTA runs >= 3 tiles — the
    // staging rows must not be re-entered by warps that are already a tile ahead; scripts/debug_tail_bwd.py)
    __syncthreads();
  };

  uint32_t it = 0;
  int64_t tile = blockIdx.x, prev_tile = -1;
  if (tile < n_tiles) load_tile(tile);
  for (; tile < n_tiles; tile += gridDim.x, ++it) {
    if (it > 0) {
      b_mbar_wait(b_smem_u32(mbar), (it - 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- row-local backward of normalise / dropout / LeakyReLU -> g_t: the A operand of this tile
    const int64_t r0 = tile * kBM;
#pragma unroll
    for (int i = 0; i < kBIters; ++i) {
      const int r = my_row0 + kBRowsPerPass * i;
      const float ks[4] = {kv[i].x ? a.keep_scale : 0.f, kv[i].y ? a.keep_scale : 0.f, kv[i].z ? a.keep_scale : 0.f,
                           kv[i].w ? a.keep_scale : 0.f};
      const float tt[4] = {tv[i].x, tv[i].y, tv[i].z, tv[i].w};
      const float gg[4] = {gv[i].x, gv[i].y, gv[i].z, gv[i].w};
      float z[4], act[4];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        act[k] = (tt[k] > 0.f ? 1.0f : a.slope) * ks[k];
        z[k] = tt[k] * act[k];
        s1 = fmaf(z[k], z[k], s1);
        s2 = fmaf(z[k], gg[k], s2);
      }
      float gt[4];
      if (a.normalize) {
        s1 = b_sum16(s1);
        s2 = b_sum16(s2);
        const float nrm = sqrtf(s1);
        if (nrm > 1e-12f) {
          const float inv = 1.0f / nrm, c = s2 / s1;
#pragma unroll
          for (int k = 0; k < 4; ++k) gt[k] = (gg[k] - z[k] * c) * inv * act[k];
        } else {   // below F.normalize's eps the forward is z / eps
#pragma unroll
          for (int k = 0; k < 4; ++k) gt[k] = gg[k] * 1e12f * act[k];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) gt[k] = gg[k] * act[k];
      }
      const int64_t row = r0 + r;
      const float4 g4 = make_float4(gt[0], gt[1], gt[2], gt[3]);
      if (row < a.n) st_stream_f4(a.g_t + row * kBD + my_chunk * 4, g4);
      const float4 hi = make_float4(b_tf32_hi(g4.x), b_tf32_hi(g4.y), b_tf32_hi(g4.z), b_tf32_hi(g4.w));
      const uint32_t off = my_chunk * kBLboA + (r >> 3) * kBSBO + (r & 7) * 16;
      *reinterpret_cast<float4*>(smem + kBOffAhi + off) = hi;
      *reinterpret_cast<float4*>(smem + kBOffAlo + off) = make_float4(g4.x - hi.x, g4.y - hi.y, g4.z - hi.z, g4.w - hi.w);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = b_smem_u32(smem);
      const uint32_t tmem_d = tmem_base + (it & 1u) * uint32_t(kBN);
      if (b_elect_one()) {
        uint32_t acc = 0;
#pragma unroll
        for (int term = 0; term < 4; ++term) {   // smallest terms first: lo.lo, lo.hi, hi.lo, hi.hi
          const uint32_t a_off = (term < 2) ? kBOffAlo : kBOffAhi;
          const uint32_t b_off = (term == 0 || term == 2) ? kBOffBlo : kBOffBhi;
#pragma unroll
          for (int ks = 0; ks < kBD / 8; ++ks) {
            b_umma(tmem_d, b_desc(base + a_off + 2 * ks * kBLboA, kBLboA), b_desc(base + b_off + 2 * ks * kBLboB, kBLboB), acc);
            acc = 1;
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b_smem_u32(mbar))
                     : "memory");
      }
      __syncwarp();
    }
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) load_tile(next);
    if (it > 0) epilogue(prev_tile, (it - 1) & 1u);
    prev_tile = tile;
  }
  if (it > 0) {
    b_mbar_wait(b_smem_u32(mbar), (it - 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue(prev_tile, (it - 1) & 1u);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kBTmemCols) : "memory");
  }
}

}  // namespace
}  // namespace b200gcn

using namespace b200gcn;

extern "C" int b200gcn_bignn_tail_backward(const float* p, int64_t ldp, const float* x, int64_t ldx, const float* w1,
                                           const float* w2, const float* t, int64_t ld_t, const uint8_t* keep,
                                           float drop_p, float slope, int normalize, const float* g_out,
                                           int64_t ld_g, int64_t n, int32_t d_in, int32_t d_out, float* g_p,
                                           float* g_x, float* g_t, float* am, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(d_in == kBD && d_out == kBD, "the fused tail backward takes d_in = d_out = 64 (got %d, %d)", d_in, d_out);
  B200_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return B200GCN_OK;
  B200_CHECK_ARG(p && x && w1 && w2 && t && g_out && g_p && g_x && g_t, "NULL input / output");
  B200_CHECK_ARG(aligned16(p) && aligned16(x) && aligned16(w1) && aligned16(w2) && aligned16(t) && aligned16(g_out) &&
                     aligned16(g_p) && aligned16(g_x) && aligned16(g_t) && (!am || aligned16(am)),
                 "16-byte alignment");
  B200_CHECK_ARG(ldp % 4 == 0 && ldx % 4 == 0 && ld_t % 4 == 0 && ld_g % 4 == 0 && ldp >= kBD && ldx >= kBD &&
                     ld_t >= kBD && ld_g >= kBD,
                 "leading dimensions");
  B200_CHECK_ARG(!keep || (reinterpret_cast<uintptr_t>(keep) & 3u) == 0, "keep must be 4-byte aligned");
  B200_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "drop_p outside [0,1)");
  BwdArgs a{p, ldp, x, ldx, w1, w2, t, ld_t, keep, 1.0f / (1.0f - drop_p), slope, normalize, g_out, ld_g, n,
            g_p, g_x, g_t, am};
  int dev = 0, sms = 148;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (n + kBM - 1) / kBM;
  const int grid = int(n_tiles < int64_t(sms) ? n_tiles : int64_t(sms));
  B200_CHECK_CUDA(cudaFuncSetAttribute(bignn_tail_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kBSmem)));
  bignn_tail_bwd_kernel<<<grid, kBThreads, kBSmem, st>>>(a);
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}
