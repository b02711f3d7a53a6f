// NGCF layer tail: BiGNNConv.forward after propagate() (recbole_gnn/model/layers.py:56-58) fused with
// the per-layer ops of NGCF.forward (ngcf.py:96-98):
//   t = (p + x) W1^T + b1 + (p * x) W2^T + b2 ; leaky_relu ; dropout mask ; row L2-normalise,
// written straight into a column slice of the [N, D*(L+1)] concat buffer (ngcf.py:100).
//
// 2 * 2 * N * d_in * d_out flops (33 GFLOP per layer at N = 2 M, 64x64) against ~3 * N * D * 4 bytes:
// ~21 flop/byte, under the fp32 CUDA-core ridge but far too small next to the SpMM (51 GB/layer) to
// be worth a tensor-core path whose tf32 inputs would not hold the 1e-4 contract after normalise.
// Classic register-tiled SGEMM: persistent CTAs, 64-row x 64-col output tile per pass, 4x4 outputs
// per thread, A = [p+x | p*x] built on the fly in shared memory, weights resident in shared memory.
#include "common.cuh"

#include <stdlib.h>

namespace b200gcn {
namespace {

constexpr int kT = 64;        // tile edge (rows, cols, k)
constexpr int kPad = kT + 4;  // row stride in floats: keeps float4 alignment, spreads banks
constexpr int kThreadsT = 256;

struct TailArgs {
  const float* p; int64_t ldp;
  const float* x; int64_t ldx;
  const float* w1; const float* b1; const float* w2; const float* b2;
  int64_t n; int d_in; int d_out;
  float slope; const uint8_t* keep; float keep_scale; int normalize;
  float* out; int64_t ldo; float* out2; int64_t ldo2; float* pre; int64_t ld_pre;
};

template <int NT>
__global__ void __launch_bounds__(kThreadsT) bignn_tail_kernel(const TailArgs a) {
  extern __shared__ __align__(16) float smem[];
  float (*As)[kPad] = reinterpret_cast<float (*)[kPad]>(smem);                   // (p + x)^T : [k][row]
  float (*Ms)[kPad] = reinterpret_cast<float (*)[kPad]>(smem + kT * kPad);       // (p * x)^T
  float (*W1s)[kPad] = reinterpret_cast<float (*)[kPad]>(smem + 2 * kT * kPad);  // W1^T tile : [k][j]
  float (*W2s)[kPad] = reinterpret_cast<float (*)[kPad]>(smem + 3 * kT * kPad);

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t n_tiles = (a.n + kT - 1) / kT;
  const int n_kc = (a.d_in + kT - 1) / kT;
  int loaded_nt = -1, loaded_kc = -1;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t r0 = tile * kT;
    float acc[NT][4][4];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[t][i][j] = 0.f;

    for (int kc = 0; kc < n_kc; ++kc) {
      __syncthreads();  // previous consumers of As/Ms/W done
      // stage A/M: 64 rows x 64 k, float4 along k, stored transposed
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * kThreadsT;
        const int r = idx >> 4, kq = idx & 15;
        const int64_t row = r0 + r;
        const int k = kc * kT + kq * 4;
        float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), xv = pv;
        if (row < a.n && k < a.d_in) {
          pv = *reinterpret_cast<const float4*>(a.p + row * a.ldp + k);
          xv = *reinterpret_cast<const float4*>(a.x + row * a.ldx + k);
        }
        As[kq * 4 + 0][r] = pv.x + xv.x; Ms[kq * 4 + 0][r] = pv.x * xv.x;
        As[kq * 4 + 1][r] = pv.y + xv.y; Ms[kq * 4 + 1][r] = pv.y * xv.y;
        As[kq * 4 + 2][r] = pv.z + xv.z; Ms[kq * 4 + 2][r] = pv.z * xv.z;
        As[kq * 4 + 3][r] = pv.w + xv.w; Ms[kq * 4 + 3][r] = pv.w * xv.w;
      }
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        if (loaded_nt != t || loaded_kc != kc) {
          if (t > 0) __syncthreads();  // earlier out-tile finished with the weight tiles
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * kThreadsT;
            const int j = idx >> 4, kq = idx & 15;
            const int jj = t * kT + j, k = kc * kT + kq * 4;
            float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
            if (jj < a.d_out && k < a.d_in) {
              u = *reinterpret_cast<const float4*>(a.w1 + int64_t(jj) * a.d_in + k);
              v = *reinterpret_cast<const float4*>(a.w2 + int64_t(jj) * a.d_in + k);
            }
            W1s[kq * 4 + 0][j] = u.x; W2s[kq * 4 + 0][j] = v.x;
            W1s[kq * 4 + 1][j] = u.y; W2s[kq * 4 + 1][j] = v.y;
            W1s[kq * 4 + 2][j] = u.z; W2s[kq * 4 + 2][j] = v.z;
            W1s[kq * 4 + 3][j] = u.w; W2s[kq * 4 + 3][j] = v.w;
          }
          loaded_nt = t;
          loaded_kc = kc;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kT; ++k) {
          const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          const float4 mv = *reinterpret_cast<const float4*>(&Ms[k][ty * 4]);
          const float4 w1v = *reinterpret_cast<const float4*>(&W1s[k][tx * 4]);
          const float4 w2v = *reinterpret_cast<const float4*>(&W2s[k][tx * 4]);
          const float ar[4] = {av.x, av.y, av.z, av.w};
          const float mr[4] = {mv.x, mv.y, mv.z, mv.w};
          const float w1r[4] = {w1v.x, w1v.y, w1v.z, w1v.w};
          const float w2r[4] = {w2v.x, w2v.y, w2v.z, w2v.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[t][i][j] = fmaf(ar[i], w1r[j], acc[t][i][j]);
              acc[t][i][j] = fmaf(mr[i], w2r[j], acc[t][i][j]);
            }
        }
      }
    }

    // epilogue: bias, optional pre-activation output, LeakyReLU, dropout mask, L2 normalise
    float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int j0 = t * kT + tx * 4;
      float bs[4] = {0.f, 0.f, 0.f, 0.f};
      if (j0 < a.d_out) {
        const float4 u = *reinterpret_cast<const float4*>(a.b1 + j0);
        const float4 v = *reinterpret_cast<const float4*>(a.b2 + j0);
        // x_trans + x_inter = (.. + b1) + (.. + b2); the two GEMVs were accumulated interleaved
        bs[0] = u.x + v.x; bs[1] = u.y + v.y; bs[2] = u.z + v.z; bs[3] = u.w + v.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t row = r0 + ty * 4 + i;
        float tv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) tv[j] = acc[t][i][j] + bs[j];
        if (a.pre != nullptr && row < a.n && j0 < a.d_out)
          *reinterpret_cast<float4*>(a.pre + row * a.ld_pre + j0) = make_float4(tv[0], tv[1], tv[2], tv[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) tv[j] = tv[j] > 0.f ? tv[j] : tv[j] * a.slope;
        if (a.keep != nullptr && row < a.n && j0 < a.d_out) {
          const uchar4 kp = *reinterpret_cast<const uchar4*>(a.keep + row * int64_t(a.d_out) + j0);
          tv[0] *= kp.x ? a.keep_scale : 0.f;
          tv[1] *= kp.y ? a.keep_scale : 0.f;
          tv[2] *= kp.z ? a.keep_scale : 0.f;
          tv[3] *= kp.w ? a.keep_scale : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[t][i][j] = tv[j];
          ss[i] += tv[j] * tv[j];
        }
      }
    }
    if (a.normalize) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) ss[i] += __shfl_xor_sync(0xffffffffu, ss[i], o, 16);
        ss[i] = fmaxf(sqrtf(ss[i]), 1e-12f);  // F.normalize eps
      }
    }
    if (a.out != nullptr) {
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int j0 = t * kT + tx * 4;
        if (j0 >= a.d_out) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t row = r0 + ty * 4 + i;
          if (row >= a.n) continue;
          float4 o = make_float4(acc[t][i][0], acc[t][i][1], acc[t][i][2], acc[t][i][3]);
          if (a.normalize) { o.x /= ss[i]; o.y /= ss[i]; o.z /= ss[i]; o.w /= ss[i]; }
          *reinterpret_cast<float4*>(a.out + row * a.ldo + j0) = o;
          if (a.out2 != nullptr) *reinterpret_cast<float4*>(a.out2 + row * a.ldo2 + j0) = o;
        }
      }
    }
  }
}

}  // namespace
}  // namespace b200gcn

using namespace b200gcn;

// tcgen05 path for d_in == d_out == 64 (bignn_tail_tc.cu)
int b200gcn_bignn_tail_tc_launch(const float* p, int64_t ldp, const float* x, int64_t ldx, const float* w1,
                                 const float* b1, const float* w2, const float* b2, int64_t n, float slope,
                                 const uint8_t* keep, float keep_scale, int normalize, float* out, int64_t ldo,
                                 float* out2, int64_t ldo2, float* pre_out, int64_t ld_pre, cudaStream_t st);

static bool tail_use_tensor_cores() {
  static const bool on = [] {
    const char* e = getenv("B200GCN_TAIL");
    return !(e && (e[0] == 'c' || e[0] == 'C'));   // B200GCN_TAIL=cuda selects the fp32 CUDA-core kernel (A/B runs)
  }();
  return on;
}

extern "C" int b200gcn_bignn_tail(const float* p, int64_t ldp, const float* x, int64_t ldx,
                                  const float* w1, const float* b1, const float* w2, const float* b2,
                                  int64_t n, int32_t d_in, int32_t d_out, float slope,
                                  const uint8_t* keep, float drop_p, int normalize, float* out,
                                  int64_t ldo, float* out2, int64_t ldo2, float* pre_out, int64_t ld_pre,
                                  void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(n >= 0, "n < 0");
  B200_CHECK_ARG(d_in > 0 && d_in % 4 == 0 && d_in <= 256, "d_in=%d must be a multiple of 4 in [4,256]", d_in);
  B200_CHECK_ARG(d_out > 0 && d_out % 4 == 0 && d_out <= 256, "d_out=%d must be a multiple of 4 in [4,256]", d_out);
  if (n == 0) return B200GCN_OK;
  B200_CHECK_ARG(p && x && w1 && b1 && w2 && b2, "NULL input");
  B200_CHECK_ARG(out || pre_out, "both out and pre_out are NULL");
  B200_CHECK_ARG(aligned16(p) && aligned16(x) && aligned16(w1) && aligned16(w2) && aligned16(b1) && aligned16(b2),
                 "inputs must be 16-byte aligned");
  B200_CHECK_ARG(ldp % 4 == 0 && ldx % 4 == 0 && ldp >= d_in && ldx >= d_in, "ldp/ldx");
  B200_CHECK_ARG(!out || (aligned16(out) && ldo % 4 == 0 && ldo >= d_out), "out alignment / ldo");
  B200_CHECK_ARG(!out2 || (out && aligned16(out2) && ldo2 % 4 == 0 && ldo2 >= d_out), "out2 needs out, alignment / ldo2");
  B200_CHECK_ARG(!pre_out || (aligned16(pre_out) && ld_pre % 4 == 0 && ld_pre >= d_out), "pre_out alignment / ld_pre");
  B200_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "drop_p outside [0,1)");
  B200_CHECK_ARG(!keep || (reinterpret_cast<uintptr_t>(keep) & 3u) == 0, "keep must be 4-byte aligned");
  if (d_in == 64 && d_out == 64 && (!keep || aligned16(keep)) && tail_use_tensor_cores())
    return b200gcn_bignn_tail_tc_launch(p, ldp, x, ldx, w1, b1, w2, b2, n, slope, keep, 1.0f / (1.0f - drop_p), normalize,
                                        out, ldo, out2, ldo2, pre_out, ld_pre, st);
  TailArgs a{p, ldp, x, ldx, w1, b1, w2, b2, n, d_in, d_out, slope, keep,
             1.0f / (1.0f - drop_p), normalize, out, ldo, out2, ldo2, pre_out, ld_pre};
  int dev = 0, sms = 148;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = (n + kT - 1) / kT;
  const int grid = int(n_tiles < int64_t(sms) * 2 ? n_tiles : int64_t(sms) * 2);
  constexpr size_t kSmem = size_t(4) * kT * kPad * sizeof(float);  // 69,632 B
#define B200_TAIL(NT)                                                                                \
  do {                                                                                               \
    B200_CHECK_CUDA(cudaFuncSetAttribute(bignn_tail_kernel<NT>,                                      \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmem)));  \
    bignn_tail_kernel<NT><<<grid, kThreadsT, kSmem, st>>>(a);                                        \
  } while (0)
  if (d_out <= 64) B200_TAIL(1);
  else if (d_out <= 128) B200_TAIL(2);
  else B200_TAIL(4);
#undef B200_TAIL
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}
