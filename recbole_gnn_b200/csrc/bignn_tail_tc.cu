// NGCF layer tail on the 5th-generation tensor cores (tcgen05 + TMEM), d_in = d_out = 64:
//   t = (p + x) W1^T + b1 + (p * x) W2^T + b2      BiGNNConv.forward after propagate(), layers.py:56-58
//   LeakyReLU -> dropout mask -> row L2-normalise        NGCF.forward, ngcf.py:96-98
// as ONE contraction  t = [p+x | p*x] (128 rows x K=128) . [W1 | W2]^T (K=128 x N=64)  per 128-row tile, issued by one
// elected thread as tcgen05.mma.kind::tf32 with the accumulator in TMEM, and made fp32-accurate by the split
//   a = a_hi + a_lo,  a_hi = a with the low 13 mantissa bits cleared (exactly what the TF32 datapath reads),
//   a_lo = a - a_hi (exact in fp32);   a.b = a_lo.b_lo + a_lo.b_hi + a_hi.b_lo + a_hi.b_hi
// (4 x 16 MMAs of 128x64x8 per tile; the only rounding left is the 2^-22 truncation of the lo parts and the fp32
// accumulation in TMEM).  The round-1 CUDA-core tail ran at 23 TFLOP/s fp32 and 0.16-0.19 of its HBM floor; this
// kernel is bound by the three N x 64 x 4-byte streams it has to move (p, x in; out (+out2) back).
//
// Pipeline of one CTA (256 threads, 1 CTA per SM, persistent over tiles; no warp specialisation).  Iteration i:
//   wait(MMAs of tile i-1)                                       [mbarrier; they ran under epilogue(i-2)]
//   regs(p,x of tile i) -> a,m -> hi/lo -> st.shared             (canonical K-major no-swizzle UMMA layout)
//   fence.proxy.async ; bar ; [thread 0] 64 x tcgen05.mma -> TMEM accumulator (i & 1) ; tcgen05.commit -> mbarrier
//   ld.global p,x of tile i+1 and the dropout mask of tile i into registers (in flight during everything below)
//   epilogue(i-1) UNDER the MMAs of tile i:  tcgen05.ld accumulator ((i-1) & 1) -> +bias, LeakyReLU, mask, partial
//       row sums of squares -> XOR-swizzled staging ; bar ; 16 lanes per row: scale by 1/max(norm, eps), out / out2
// (first build, profiles/r2_tail_tc_v1.txt: everything serialised per tile, shuffle reductions and four IEEE
// divisions per element group in the coalesced phase -> 54 % of the samples in the epilogue, 0.69 ms; this build
// moves the maths to the thread-per-row phase and overlaps it with the tensor core.)
#include "common.cuh"

namespace b200gcn {
namespace {

constexpr int kTM = 128;                  // rows per tile = UMMA M
constexpr int kD = 64;                    // d_in = d_out
constexpr int kK = 2 * kD;                // concatenated K: [p+x | p*x]
constexpr int kChunks = kK / 4;           // 16-byte K chunks per row (4 tf32 each)
constexpr int kThreadsTc = 512;           // 16 warps: the fill / epilogue phases are latency-bound, not issue-bound
constexpr int kRowsPerPass = kThreadsTc / 16;    // rows covered by one pass of the 16-lanes-per-row mappings
constexpr int kLoadIters = kTM / kRowsPerPass;   // passes per tile
constexpr int kColParts = kThreadsTc / 128;      // column groups of the accumulator read (warps per TMEM lane quarter)
constexpr int kColsPerThread = kD / kColParts;   // 16
// canonical K-major SWIZZLE_NONE layout: core matrix = 8 rows x 16 bytes, contiguous (128 B);
//   byte offset of (row r, chunk c) = c * LBO + (r / 8) * SBO + (r % 8) * 16,  SBO = 128.
// LBO gets one extra 16-byte slot so that the 16 lanes of a half-warp that hold the 16 chunks of ONE row hit 16
// different bank groups when they store (coalesced global loads want consecutive lanes on consecutive chunks).
constexpr uint32_t kSBO = 128;
constexpr uint32_t kLboA = kTM * 16 + 16;           // 2064
constexpr uint32_t kLboB = kD * 16;                 // 1024 (written once: bank conflicts there do not matter)
constexpr uint32_t kBytesA = kChunks * kLboA;       // 66,048 per hi / lo part
constexpr uint32_t kBytesB = kChunks * kLboB;       // 33,280 per hi / lo part
constexpr uint32_t kOffAhi = 0, kOffAlo = kBytesA, kOffBhi = 2 * kBytesA, kOffBlo = 2 * kBytesA + kBytesB;
constexpr uint32_t kOffStage = 2 * kBytesA + 2 * kBytesB;  // [128][64] fp32, 16-byte slots XOR-swizzled by row & 7
constexpr uint32_t kOffMisc = kOffStage + kTM * kD * 4;        // bias[64] floats, mbarrier, tmem address
constexpr uint32_t kSmemTc = kOffMisc + 64 * 4 + 16 + 16;
static_assert(kSmemTc <= 227 * 1024, "shared memory budget");
constexpr uint32_t kTmemCols = 128;                        // two 128 x 64 fp32 accumulators

struct TcArgs {
  const float* p; int64_t ldp;
  const float* x; int64_t ldx;
  const float* w1; const float* b1; const float* w2; const float* b2;
  int64_t n;
  float slope; const uint8_t* keep; float keep_scale; int normalize;
  float* out; int64_t ldo; float* out2; int64_t ldo2; float* pre; int64_t ld_pre;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout SWIZZLE_NONE (0) << 61
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  return uint64_t((smem_addr & 0x3ffffu) >> 4) | (uint64_t(lbo >> 4) << 16) | (uint64_t(sbo >> 4) << 32) |
         (uint64_t(1) << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N = 64, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kD >> 3) << 17) | (uint32_t(kTM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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

// B operand: W = [W1 | W2] as N = 64 rows (output feature j) x K = 128, K-major, split into hi / lo
__device__ __forceinline__ void stage_weights(const TcArgs& a, char* smem) {
  for (int idx = threadIdx.x; idx < kD * kChunks; idx += kThreadsTc) {
    const int j = idx / kChunks, c = idx % kChunks;           // chunk c: k = 4c .. 4c+3 of the concatenated K
    const float* src = (c < kChunks / 2) ? a.w1 + j * kD + c * 4 : a.w2 + j * kD + (c - kChunks / 2) * 4;
    const float4 w = *reinterpret_cast<const float4*>(src);
    const float4 h = make_float4(tf32_hi(w.x), tf32_hi(w.y), tf32_hi(w.z), tf32_hi(w.w));
    const float4 l = make_float4(w.x - h.x, w.y - h.y, w.z - h.z, w.w - h.w);
    const uint32_t off = c * kLboB + (j >> 3) * kSBO + (j & 7) * 16;
    *reinterpret_cast<float4*>(smem + kOffBhi + off) = h;
    *reinterpret_cast<float4*>(smem + kOffBlo + off) = l;
  }
}

// One lane of a converged warp (elect.sync).  The MMA loop below is executed by the WHOLE warp with warp-uniform
// values and only the tcgen05 instructions are predicated on the elected lane: issued from a divergent `tid == 0`
// branch instead, every descriptor goes through a per-thread -> uniform-register "waterfall" (ELECT / R2UR /
// BRA.U.ANY per operand, ~20 instructions per MMA) and the issue loop alone costs ~2 us per tile
// (profiles/r2_tail_tc_v2.txt: 32 % of warp 0's samples).
__device__ __forceinline__ bool elect_one() {
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

__global__ void __launch_bounds__(kThreadsTc, 1) bignn_tail_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(1024) char smem[];
  float* bias_s = reinterpret_cast<float*>(smem + kOffMisc);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + kOffMisc + 64 * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffMisc + 64 * 4 + 16);
  char* stage = smem + kOffStage;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (a.n + kTM - 1) / kTM;

  // ---- one-time setup: TMEM columns, mbarrier, weights and bias in shared memory
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < kD) bias_s[tid] = a.b1[tid] + a.b2[tid];   // x_trans + x_inter = (.. + b1) + (.. + b2)
  stage_weights(a, smem);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;


  // thread -> (row, chunk) for the global loads: 16 consecutive lanes read the 256 contiguous bytes of one row
  const int my_chunk = tid & 15;
  const int my_row0 = tid >> 4;                               // rows my_row0 + kRowsPerPass i
  // thread -> (row, column group) for the accumulator: warp w reads TMEM lanes 32 (w % 4) .. +31, columns 16 (w / 4) .. +15
  const int q = warp & 3, h = warp >> 2;
  const int my_acc_row = q * 32 + lane;
  float4 pv[kLoadIters], xv[kLoadIters];
  uint4 kp_new, kp_old;
  kp_new = kp_old = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);

  auto load_tile = [&](int64_t tile) {
    const int64_t r0 = tile * kTM;
#pragma unroll
    for (int i = 0; i < kLoadIters; ++i) {
      const int64_t row = r0 + my_row0 + kRowsPerPass * i;
      if (row < a.n) {
        pv[i] = ld_gather_f4(a.p + row * a.ldp + my_chunk * 4);
        xv[i] = ld_gather_f4(a.x + row * a.ldx + my_chunk * 4);
      } else {
        pv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        xv[i] = pv[i];
      }
    }
  };
  auto load_mask = [&](int64_t tile) {   // the 16 keep flags of (my_acc_row, columns 16 h ..): one 16-byte load
    const int64_t row = tile * kTM + my_acc_row;
    if (a.keep != nullptr && row < a.n)
      kp_new = __ldg(reinterpret_cast<const uint4*>(a.keep + row * int64_t(kD) + h * kColsPerThread));
  };

  // Everything that happens to the finished accumulator of `tile` (TMEM buffer `buf`).
  auto epilogue = [&](int64_t tile, uint32_t buf) {
    const int64_t r0 = tile * kTM;
    {
      static_assert(kColsPerThread == 16, "the accumulator read below is the .x16 shape");
      uint32_t v[16];
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + buf * uint32_t(kD) + uint32_t(h * kColsPerThread);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
            "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int64_t row = r0 + my_acc_row;
      const bool live = row < a.n;
      const uint32_t kw[4] = {kp_old.x, kp_old.y, kp_old.z, kp_old.w};
#pragma unroll
      for (int j = 0; j < kColsPerThread; j += 4) {
        float t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = __uint_as_float(v[j + k]) + bias_s[h * kColsPerThread + j + k];
        if (a.pre != nullptr && live)   // training only: the pre-activation rows, 64 contiguous bytes per thread
          *reinterpret_cast<float4*>(a.pre + row * a.ld_pre + h * kColsPerThread + j) = make_float4(t[0], t[1], t[2], t[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          t[k] = t[k] > 0.f ? t[k] : t[k] * a.slope;
          if (a.keep != nullptr) t[k] *= ((kw[j >> 2] >> (8 * k)) & 0xffu) ? a.keep_scale : 0.f;
        }
        // staging slot of (row, 16-byte column group c4): c4 ^ (row & 7) — conflict-free for this thread-per-row
        // store and for the 16-lanes-per-row load below
        const int c4 = (h * kColsPerThread + j) >> 2;
        *reinterpret_cast<float4*>(stage + my_acc_row * (kD * 4) + ((c4 ^ (my_acc_row & 7)) << 4)) =
            make_float4(t[0], t[1], t[2], t[3]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (a.out != nullptr) {   // 16 lanes per row, coalesced; one pass covers kRowsPerPass consecutive rows of the tile
      const int lig = lane & 15;
      float4 t[kLoadIters];
      float inv[kLoadIters];
#pragma unroll
      for (int i = 0; i < kLoadIters; ++i) {
        const int r = my_row0 + kRowsPerPass * i;
        t[i] = *reinterpret_cast<const float4*>(stage + r * (kD * 4) + ((lig ^ (r & 7)) << 4));
        inv[i] = t[i].x * t[i].x + t[i].y * t[i].y + t[i].z * t[i].z + t[i].w * t[i].w;
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
        for (int i = 0; i < kLoadIters; ++i) inv[i] += __shfl_xor_sync(0xffffffffu, inv[i], o, 16);
      }
#pragma unroll
      for (int i = 0; i < kLoadIters; ++i)
        inv[i] = a.normalize ? 1.0f / fmaxf(sqrtf(inv[i]), 1e-12f) : 1.0f;   // F.normalize eps
#pragma unroll
      for (int i = 0; i < kLoadIters; ++i) {
        const int64_t row = r0 + my_row0 + kRowsPerPass * i;
        if (row < a.n) {
          const float4 o = make_float4(t[i].x * inv[i], t[i].y * inv[i], t[i].z * inv[i], t[i].w * inv[i]);
          st_stream_f4(a.out + row * a.ldo + lig * 4, o);
          if (a.out2 != nullptr) st_stream_f4(a.out2 + row * a.ldo2 + lig * 4, o);
        }
      }
    }
    __syncthreads();   // no warp re-enters the staging rows a tile ahead (see bignn_tail_bwd_tc.cu)
  };

  uint32_t it = 0;
  int64_t tile = blockIdx.x, prev_tile = -1;
  if (tile < n_tiles) load_tile(tile);
  for (; tile < n_tiles; tile += gridDim.x, ++it) {
    if (it > 0) {   // the MMAs of the previous tile are done: its accumulator is complete, the A operand is free
      mbar_wait(smem_u32(mbar), (it - 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // ---- A operand of this tile: [p + x | p * x], hi and lo parts
#pragma unroll
    for (int i = 0; i < kLoadIters; ++i) {
      const int r = my_row0 + kRowsPerPass * i;
      const float4 s = make_float4(pv[i].x + xv[i].x, pv[i].y + xv[i].y, pv[i].z + xv[i].z, pv[i].w + xv[i].w);
      const float4 m = make_float4(pv[i].x * xv[i].x, pv[i].y * xv[i].y, pv[i].z * xv[i].z, pv[i].w * xv[i].w);
      const float4 sh = make_float4(tf32_hi(s.x), tf32_hi(s.y), tf32_hi(s.z), tf32_hi(s.w));
      const float4 mh = make_float4(tf32_hi(m.x), tf32_hi(m.y), tf32_hi(m.z), tf32_hi(m.w));
      const uint32_t row_off = (r >> 3) * kSBO + (r & 7) * 16;
      const uint32_t off_s = my_chunk * kLboA + row_off;
      const uint32_t off_m = (my_chunk + kChunks / 2) * kLboA + row_off;
      *reinterpret_cast<float4*>(smem + kOffAhi + off_s) = sh;
      *reinterpret_cast<float4*>(smem + kOffAhi + off_m) = mh;
      *reinterpret_cast<float4*>(smem + kOffAlo + off_s) = make_float4(s.x - sh.x, s.y - sh.y, s.z - sh.z, s.w - sh.w);
      *reinterpret_cast<float4*>(smem + kOffAlo + off_m) = make_float4(m.x - mh.x, m.y - mh.y, m.z - mh.z, m.w - mh.w);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // also: the previous epilogue's staging reads and TMEM loads are all done
    if (warp == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = smem_u32(smem);
      const uint32_t tmem_d = tmem_base + (it & 1u) * uint32_t(kD);
      if (elect_one()) {
        uint32_t acc = 0;
        // smallest terms first: lo.lo, lo.hi, hi.lo, hi.hi
#pragma unroll
        for (int term = 0; term < 4; ++term) {
          const uint32_t a_off = (term < 2) ? kOffAlo : kOffAhi;
          const uint32_t b_off = (term == 0 || term == 2) ? kOffBlo : kOffBhi;
#pragma unroll
          for (int ks = 0; ks < kK / 8; ++ks) {          // one MMA = K 8 = two 16-byte chunks
            const uint64_t da = umma_desc(base + a_off + 2 * ks * kLboA, kLboA, kSBO);
            const uint64_t db = umma_desc(base + b_off + 2 * ks * kLboB, kLboB, kSBO);
            umma_tf32(tmem_d, da, db, acc);
            acc = 1;
          }
        }
        // arrives on the mbarrier when every MMA above has finished reading shared memory and writing TMEM
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar))
                     : "memory");
      }
      __syncwarp();
    }
    // ---- rows of the next tile and the mask of this one leave HBM while the tensor core works
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles) load_tile(next);
    load_mask(tile);
    // ---- epilogue of the PREVIOUS tile, under this tile's MMAs
    if (it > 0) epilogue(prev_tile, (it - 1) & 1u);
    kp_old = kp_new;
    prev_tile = tile;
  }
  if (it > 0) {
    mbar_wait(smem_u32(mbar), (it - 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    epilogue(prev_tile, (it - 1) & 1u);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace
}  // namespace b200gcn

using namespace b200gcn;

// Internal entry used by b200gcn_bignn_tail (bignn_tail.cu) when d_in == d_out == 64.
int b200gcn_bignn_tail_tc_launch(const float* p, int64_t ldp, const float* x, int64_t ldx, const float* w1,
                                 const float* b1, const float* w2, const float* b2, int64_t n, float slope,
                                 const uint8_t* keep, float keep_scale, int normalize, float* out, int64_t ldo,
                                 float* out2, int64_t ldo2, float* pre_out, int64_t ld_pre, cudaStream_t st) {
  TcArgs a{p, ldp, x, ldx, w1, b1, w2, b2, n, slope, keep, keep_scale, normalize, out, ldo, out2, ldo2, pre_out, ld_pre};
  int dev = 0, sms = 148;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (n + kTM - 1) / kTM;
  const int grid = int(n_tiles < int64_t(sms) ? n_tiles : int64_t(sms));
  B200_CHECK_CUDA(cudaFuncSetAttribute(bignn_tail_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemTc)));
  bignn_tail_tc_kernel<<<grid, kThreadsTc, kSmemTc, st>>>(a);
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}
