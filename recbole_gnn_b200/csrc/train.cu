// The training step around the propagation: what `LightGCN.calculate_loss` (lightgcn.py:83-110) and
// `NGCF.calculate_loss` (ngcf.py:106-123) do with the propagated tables, and the optimiser update that follows
// (`recbole.trainer.Trainer._train_epoch`: loss.backward(); optimizer.step() with torch.optim.Adam).
//
//   bpr_rows_kernel    one 16-lane group per (user, pos, neg) sample: gathers the three propagated rows, forms the two
//                      scores, the BPR term -log(gamma + sigmoid(s+ - s-)) and ITS gradient rows, which are scattered
//                      straight into the (zero-initialised) gradient tables of the propagation output with fp32
//                      atomics; gathers the three EmbLoss rows and reduces their squared norms.  Per-CTA partial sums
//                      go to a small buffer (fixed order -> deterministic loss value).
//   bpr_finish_kernel  reduces the partials in order, forms loss = mf + reg_weight * EmbLoss (both recbole 1.1.1
//                      variants: require_pow False = plain norms / B, True = squared norms / B / 2).
//   bpr_reg_grad_kernel  EmbLoss gradient rows (needs the batch norms, hence a second pass) scattered with atomics.
//   adam_kernel        torch.optim.Adam (no amsgrad, optional L2 weight decay) over a [n, D] table in one pass:
//                      7 streams of n*D*4 bytes, HBM-bound.
#include "common.cuh"

namespace b200gcn {
namespace {

constexpr int kBprCta = 256;
constexpr int kBprG = 16;       // lanes per sample (float4 per lane per pass)

struct BprArgs {
  const float* u_all; int64_t ld_u;       // propagated user rows  [n_users, D]
  const float* i_all; int64_t ld_i;       // propagated item rows  [n_items, D]
  const float* reg_u; int64_t ld_ru;      // EmbLoss tables (ego tables for LightGCN, the propagated ones for NGCF)
  const float* reg_i; int64_t ld_ri;
  const int64_t* user; const int64_t* pos; const int64_t* neg;
  int64_t batch; int32_t dim;
  float gamma;
  float* g_u_all; int64_t ld_gu;          // += dL/d u_all rows (atomic)
  float* g_i_all; int64_t ld_gi;
  float* partial;                          // [grid, 4]: sum of -log terms, sum sq of the three reg row groups
};

// Sum over the 16 lanes of the caller's half-warp.  The mask names only that half: the two halves of a warp walk
// different samples and may leave the sample loop at different times.
__device__ __forceinline__ float group16_sum(float v) {
  const unsigned mask = 0xffffu << (threadIdx.x & 16u);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, 16);
  return v;
}

__global__ void __launch_bounds__(kBprCta) bpr_rows_kernel(const BprArgs a) {
  __shared__ float red[kBprCta / kBprG][4];
  const int lig = threadIdx.x & (kBprG - 1);
  const int grp = threadIdx.x / kBprG;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float inv_b = 1.0f / float(a.batch);
  for (int64_t k = int64_t(blockIdx.x) * (kBprCta / kBprG) + grp; k < a.batch; k += int64_t(gridDim.x) * (kBprCta / kBprG)) {
    const int64_t u = a.user[k], p = a.pos[k], n = a.neg[k];
    const float* ur = a.u_all + u * a.ld_u;
    const float* pr = a.i_all + p * a.ld_i;
    const float* nr = a.i_all + n * a.ld_i;
    float sp = 0.f, sn = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f;
    for (int c = lig * 4; c < a.dim; c += kBprG * 4) {
      const float4 uv = *reinterpret_cast<const float4*>(ur + c);
      const float4 pv = *reinterpret_cast<const float4*>(pr + c);
      const float4 nv = *reinterpret_cast<const float4*>(nr + c);
      sp += uv.x * pv.x + uv.y * pv.y + uv.z * pv.z + uv.w * pv.w;
      sn += uv.x * nv.x + uv.y * nv.y + uv.z * nv.z + uv.w * nv.w;
      const float4 a0 = *reinterpret_cast<const float4*>(a.reg_u + u * a.ld_ru + c);
      const float4 a1 = *reinterpret_cast<const float4*>(a.reg_i + p * a.ld_ri + c);
      const float4 a2 = *reinterpret_cast<const float4*>(a.reg_i + n * a.ld_ri + c);
      q0 += a0.x * a0.x + a0.y * a0.y + a0.z * a0.z + a0.w * a0.w;
      q1 += a1.x * a1.x + a1.y * a1.y + a1.z * a1.z + a1.w * a1.w;
      q2 += a2.x * a2.x + a2.y * a2.y + a2.z * a2.z + a2.w * a2.w;
    }
    sp = group16_sum(sp);
    sn = group16_sum(sn);
    const float d = sp - sn;
    const float sig = 1.0f / (1.0f + expf(-d));
    // d/dd of -log(gamma + sigmoid(d)) / B
    const float coef = -inv_b * sig * (1.0f - sig) / (a.gamma + sig);
    if (lig == 0) acc[0] += -logf(a.gamma + sig);
    acc[1] += q0; acc[2] += q1; acc[3] += q2;
    if (a.g_u_all != nullptr) {
      float* gu = a.g_u_all + u * a.ld_gu;
      float* gp = a.g_i_all + p * a.ld_gi;
      float* gn = a.g_i_all + n * a.ld_gi;
      for (int c = lig * 4; c < a.dim; c += kBprG * 4) {
        const float4 uv = *reinterpret_cast<const float4*>(ur + c);
        const float4 pv = *reinterpret_cast<const float4*>(pr + c);
        const float4 nv = *reinterpret_cast<const float4*>(nr + c);
        atomicAdd(gu + c + 0, coef * (pv.x - nv.x)); atomicAdd(gu + c + 1, coef * (pv.y - nv.y));
        atomicAdd(gu + c + 2, coef * (pv.z - nv.z)); atomicAdd(gu + c + 3, coef * (pv.w - nv.w));
        atomicAdd(gp + c + 0, coef * uv.x); atomicAdd(gp + c + 1, coef * uv.y);
        atomicAdd(gp + c + 2, coef * uv.z); atomicAdd(gp + c + 3, coef * uv.w);
        atomicAdd(gn + c + 0, -coef * uv.x); atomicAdd(gn + c + 1, -coef * uv.y);
        atomicAdd(gn + c + 2, -coef * uv.z); atomicAdd(gn + c + 3, -coef * uv.w);
      }
    }
  }
#pragma unroll
  for (int j = 1; j < 4; ++j) acc[j] = group16_sum(acc[j]);
  if (lig == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) red[grp][j] = acc[j];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
    for (int g = 0; g < kBprCta / kBprG; ++g) s += red[g][threadIdx.x];
    a.partial[blockIdx.x * 4 + threadIdx.x] = s;
  }
}

// out[0] = loss, out[1] = mf, out[2..4] = the three batch norms (sqrt of the squared sums)
__global__ void bpr_finish_kernel(const float* __restrict__ partial, int n_part, int64_t batch, float reg_weight,
                                  int require_pow, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  for (int b = 0; b < n_part; ++b)
    for (int j = 0; j < 4; ++j) s[j] += partial[b * 4 + j];
  const float mf = s[0] / float(batch);
  const float n0 = sqrtf(s[1]), n1 = sqrtf(s[2]), n2 = sqrtf(s[3]);
  float reg;
  if (require_pow) reg = (s[1] + s[2] + s[3]) / float(batch) / 2.0f;   // sum ||e||^2 / B / norm
  else reg = (n0 + n1 + n2) / float(batch);                             // sum ||e|| / B
  out[0] = mf + reg_weight * reg;
  out[1] = mf;
  out[2] = n0; out[3] = n1; out[4] = n2;
}

struct RegArgs {
  const float* reg_u; int64_t ld_ru; const float* reg_i; int64_t ld_ri;
  const int64_t* user; const int64_t* pos; const int64_t* neg;
  int64_t batch; int32_t dim;
  float reg_weight; int require_pow;
  const float* stats;                      // out[] of bpr_finish_kernel
  float* g_reg_u; int64_t ld_gru; float* g_reg_i; int64_t ld_gri;
};

__global__ void __launch_bounds__(kBprCta) bpr_reg_grad_kernel(const RegArgs a) {
  const int lig = threadIdx.x & (kBprG - 1);
  const int grp = threadIdx.x / kBprG;
  const float inv_b = 1.0f / float(a.batch);
  float c0, c1, c2;
  if (a.require_pow) {
    c0 = c1 = c2 = a.reg_weight * inv_b;                    // d/de of ||e||^2 / B / 2 = e / B
  } else {
    // d/de of ||E||_F / B = E / ||E||_F / B ; a zero norm has a zero (sub)gradient in torch
    const float n0 = a.stats[2], n1 = a.stats[3], n2 = a.stats[4];
    c0 = n0 > 0.f ? a.reg_weight * inv_b / n0 : 0.f;
    c1 = n1 > 0.f ? a.reg_weight * inv_b / n1 : 0.f;
    c2 = n2 > 0.f ? a.reg_weight * inv_b / n2 : 0.f;
  }
  for (int64_t k = int64_t(blockIdx.x) * (kBprCta / kBprG) + grp; k < a.batch; k += int64_t(gridDim.x) * (kBprCta / kBprG)) {
    const int64_t u = a.user[k], p = a.pos[k], n = a.neg[k];
    for (int c = lig * 4; c < a.dim; c += kBprG * 4) {
      const float4 a0 = *reinterpret_cast<const float4*>(a.reg_u + u * a.ld_ru + c);
      const float4 a1 = *reinterpret_cast<const float4*>(a.reg_i + p * a.ld_ri + c);
      const float4 a2 = *reinterpret_cast<const float4*>(a.reg_i + n * a.ld_ri + c);
      float* gu = a.g_reg_u + u * a.ld_gru + c;
      float* gp = a.g_reg_i + p * a.ld_gri + c;
      float* gn = a.g_reg_i + n * a.ld_gri + c;
      atomicAdd(gu + 0, c0 * a0.x); atomicAdd(gu + 1, c0 * a0.y); atomicAdd(gu + 2, c0 * a0.z); atomicAdd(gu + 3, c0 * a0.w);
      atomicAdd(gp + 0, c1 * a1.x); atomicAdd(gp + 1, c1 * a1.y); atomicAdd(gp + 2, c1 * a1.z); atomicAdd(gp + 3, c1 * a1.w);
      atomicAdd(gn + 0, c2 * a2.x); atomicAdd(gn + 1, c2 * a2.y); atomicAdd(gn + 2, c2 * a2.z); atomicAdd(gn + 3, c2 * a2.w);
    }
  }
}

// torch.optim.Adam, single tensor, amsgrad off:  exp_avg.lerp_(g, 1-b1); exp_avg_sq = b2 v + (1-b2) g g;
// denom = sqrt(v) / sqrt(bc2) + eps; p -= (lr / bc1) * m / denom   (weight_decay: g += wd * p first)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n4,
                                                   float b1, float b2, float eps, float step_size, float bc2_sqrt,
                                                   float weight_decay) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += int64_t(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = gp[k];
      if (weight_decay != 0.f) gk = fmaf(weight_decay, pp[k], gk);
      mp[k] = mp[k] + (gk - mp[k]) * (1.0f - b1);
      vp[k] = vp[k] * b2 + gk * gk * (1.0f - b2);
      const float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
      pp[k] = pp[k] - step_size * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
}

}  // namespace
}  // namespace b200gcn

using namespace b200gcn;

extern "C" int b200gcn_bpr_loss_workspace(int64_t batch, size_t* bytes) {
  B200_CHECK_ARG(bytes != nullptr && batch >= 0, "bad arguments");
  const int64_t per_cta = kBprCta / kBprG;
  int64_t grid = (batch + per_cta - 1) / per_cta;
  if (grid > 1024) grid = 1024;
  if (grid < 1) grid = 1;
  *bytes = align_up(size_t(grid) * 4 * sizeof(float)) + 256;
  return B200GCN_OK;
}

extern "C" int b200gcn_bpr_loss(const float* u_all, int64_t ld_u, const float* i_all, int64_t ld_i, const float* reg_u,
                                int64_t ld_ru, const float* reg_i, int64_t ld_ri, const int64_t* user,
                                const int64_t* pos, const int64_t* neg, int64_t batch, int32_t dim, float gamma,
                                float reg_weight, int require_pow, float* g_u_all, int64_t ld_gu, float* g_i_all,
                                int64_t ld_gi, float* g_reg_u, int64_t ld_gru, float* g_reg_i, int64_t ld_gri,
                                float* loss_out, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(batch > 0 && dim > 0 && dim % 4 == 0, "batch=%lld dim=%d", (long long)batch, dim);
  B200_CHECK_ARG(u_all && i_all && reg_u && reg_i && user && pos && neg && loss_out && workspace, "NULL input");
  B200_CHECK_ARG((g_u_all == nullptr) == (g_i_all == nullptr) && (g_reg_u == nullptr) == (g_reg_i == nullptr),
                 "gradient tables come in pairs");
  B200_CHECK_ARG(aligned16(u_all) && aligned16(i_all) && aligned16(reg_u) && aligned16(reg_i) && ld_u % 4 == 0 &&
                     ld_i % 4 == 0 && ld_ru % 4 == 0 && ld_ri % 4 == 0,
                 "tables must be 16-byte aligned with leading dimensions %% 4 == 0");
  size_t need = 0;
  b200gcn_bpr_loss_workspace(batch, &need);
  if (workspace_bytes < need) {
    set_error("workspace %zu < required %zu", workspace_bytes, need);
    return B200GCN_ERR_WORKSPACE;
  }
  const int64_t per_cta = kBprCta / kBprG;
  int64_t grid = (batch + per_cta - 1) / per_cta;
  if (grid > 1024) grid = 1024;
  float* partial = static_cast<float*>(workspace);
  BprArgs a{u_all, ld_u, i_all, ld_i, reg_u, ld_ru, reg_i, ld_ri, user, pos, neg, batch, dim, gamma,
            g_u_all, ld_gu, g_i_all, ld_gi, partial};
  bpr_rows_kernel<<<unsigned(grid), kBprCta, 0, st>>>(a);
  B200_CHECK_LAUNCH();
  bpr_finish_kernel<<<1, 32, 0, st>>>(partial, int(grid), batch, reg_weight, require_pow, loss_out);
  B200_CHECK_LAUNCH();
  if (g_reg_u != nullptr && reg_weight != 0.f) {
    RegArgs r{reg_u, ld_ru, reg_i, ld_ri, user, pos, neg, batch, dim, reg_weight, require_pow, loss_out,
              g_reg_u, ld_gru, g_reg_i, ld_gri};
    bpr_reg_grad_kernel<<<unsigned(grid), kBprCta, 0, st>>>(r);
    B200_CHECK_LAUNCH();
  }
  return B200GCN_OK;
}

extern "C" int b200gcn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel,
                                 float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                                 void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_CHECK_ARG(numel >= 0 && numel % 4 == 0 && step >= 1, "numel must be a multiple of 4, step >= 1");
  if (numel == 0) return B200GCN_OK;
  B200_CHECK_ARG(param && grad && exp_avg && exp_avg_sq, "NULL input");
  B200_CHECK_ARG(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq), "16-byte alignment");
  const double bc1 = 1.0 - pow(double(beta1), double(step));
  const double bc2 = 1.0 - pow(double(beta2), double(step));
  const float step_size = float(double(lr) / bc1);
  const float bc2_sqrt = float(sqrt(bc2));
  int dev = 0, sms = 148;
  B200_CHECK_CUDA(cudaGetDevice(&dev));
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n4 = numel / 4;
  int64_t grid = (n4 + 255) / 256;
  if (grid > int64_t(sms) * 8) grid = int64_t(sms) * 8;     // a multiple of the SM count, grid-stride loop
  adam_kernel<<<unsigned(grid), 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n4, beta1, beta2, eps, step_size,
                                              bc2_sqrt, weight_decay);
  B200_CHECK_LAUNCH();
  return B200GCN_OK;
}
