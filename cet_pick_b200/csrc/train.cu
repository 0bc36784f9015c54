// Device pieces of the refinement TRAINING step (BASELINE.json configs[4]) that sit around the
// network's forward / backward: the loss of cet_pick/trains/tomo_cr_semi_trainer.py:43-112 without `--contrastive`
// (= `_sigmoid` + PULoss, cet_pick/models/loss.py:255-325) with its gradient w.r.t. the heat-map logits, the
// consistency MSE (loss.py:701-715), and the optimiser step of main.py:55 (torch.optim.Adam defaults) fused over ONE
// flat parameter bucket -- the same bucket the gradient all-reduce travels in (cet_pick_b200/trains/step.py).
// The backward kernels of the U-Net itself are not built.
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace cetpick {
namespace {

constexpr int RED_BLOCKS = 256, RED_THREADS = 256, NSUM = 8;

// sigmoid + clamp exactly like models/utils.py:167-169
__device__ __forceinline__ float sigmoid_clamp(float x, bool& clamped) {
  const float y = 1.0f / (1.0f + expf(-x));
  clamped = (y < 1e-4f) || (y > 1.0f - 1e-4f);
  return fminf(fmaxf(y, 1e-4f), 1.0f - 1e-4f);
}

// partial sums of loss.py:262-291: [n_pos, n_soft, n_unl, S1, S2, S3, S4, S5]
//   S1 = sum log(p)(1-p)^2 pos        S2 = sum log(1-p) p^2 pos
//   S3 = sum log(1-p) p^2 (1-gt)^4 soft   S4 = sum log(p)(1-p)^2 gt^4 soft   S5 = sum p^2 log(1-p) unl
__global__ void __launch_bounds__(RED_THREADS) pu_partial_kernel(const float* __restrict__ logits, const float* __restrict__ gt,
                                                                 long long n, int apply_sigmoid, double* __restrict__ part) {
  double s[NSUM] = {};
  for (long long i = blockIdx.x * (long long)RED_THREADS + threadIdx.x; i < n; i += (long long)RED_BLOCKS * RED_THREADS) {
    bool cl;
    const float p = apply_sigmoid ? sigmoid_clamp(logits[i], cl) : logits[i];
    const float g = gt[i];
    const float lp = logf(p), ln = logf(1.0f - p);
    const float a = lp * (1.0f - p) * (1.0f - p), b = ln * p * p;
    if (g == 1.0f) { s[0] += 1.0; s[3] += a; s[4] += b; }
    else if (g == -1.0f) { s[2] += 1.0; s[7] += b; }
    else if (g > -1.0f && g < 1.0f) {
      const float w1 = (1.0f - g) * (1.0f - g) * (1.0f - g) * (1.0f - g), w2 = g * g * g * g;
      s[1] += 1.0; s[5] += b * w1; s[6] += a * w2;
    }
  }
  __shared__ double sh[NSUM][RED_THREADS];
  for (int k = 0; k < NSUM; ++k) sh[k][threadIdx.x] = s[k];
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < NSUM; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < NSUM) part[blockIdx.x * NSUM + threadIdx.x] = sh[threadIdx.x][0];
}

// out[0] = loss, out[1] = pos_risk, out[2] = neg_risk, out[3] = n_pos; coef[0..5] for the gradient kernel
__global__ void pu_final_kernel(const double* __restrict__ part, double tau, double beta, float* __restrict__ out,
                                float* __restrict__ coef) {
  double s[NSUM] = {};
  for (int b = 0; b < RED_BLOCKS; ++b)
    for (int k = 0; k < NSUM; ++k) s[k] += part[b * NSUM + k];      // fixed order: reproducible
  const double n_pos = s[0], n_soft = s[1], n_unl = s[2];
  double pos_term = n_pos > 0 ? -s[3] / n_pos : 0.0, neg_pos = n_pos > 0 ? -s[4] / n_pos : 0.0;
  if (n_soft > 0) { pos_term -= s[5] / n_soft; neg_pos -= s[6] / n_soft; }
  const double pos_risk = pos_term * tau;
  const double unl_risk = n_unl > 0 ? -s[7] / n_unl : 0.0;
  const double neg_risk = -tau * neg_pos + unl_risk;
  const bool use_neg = !(neg_risk < -beta);
  out[0] = (float)(use_neg ? pos_risk + neg_risk : pos_risk);
  out[1] = (float)pos_risk; out[2] = (float)neg_risk; out[3] = (float)n_pos;
  coef[0] = n_pos > 0 ? (float)(1.0 / n_pos) : 0.f;
  coef[1] = n_soft > 0 ? (float)(1.0 / n_soft) : 0.f;
  coef[2] = n_unl > 0 ? (float)(1.0 / n_unl) : 0.f;
  coef[3] = (float)tau;
  coef[4] = use_neg ? 1.f : 0.f;
}

// d loss / d logit (through `_sigmoid`: zero where the clamp is active) or d loss / d p when apply_sigmoid == 0
__global__ void __launch_bounds__(256) pu_grad_kernel(const float* __restrict__ logits, const float* __restrict__ gt, long long n,
                                                      int apply_sigmoid, const float* __restrict__ coef, float scale,
                                                      float* __restrict__ grad) {
  const float inv_pos = coef[0], inv_soft = coef[1], inv_unl = coef[2], tau = coef[3], use_neg = coef[4];
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    bool cl = false;
    const float p = apply_sigmoid ? sigmoid_clamp(logits[i], cl) : logits[i];
    const float g = gt[i];
    const float lp = logf(p), ln = logf(1.0f - p), q = 1.0f - p;
    const float da = q * q / p - 2.0f * q * lp;          // d/dp [log(p)(1-p)^2]
    const float db = -p * p / q + 2.0f * p * ln;         // d/dp [log(1-p) p^2]
    float dpos = 0.f, dnegpos = 0.f, dunl = 0.f;
    if (g == 1.0f) { dpos = -da * inv_pos; dnegpos = -db * inv_pos; }
    else if (g == -1.0f) dunl = -db * inv_unl;
    else if (g > -1.0f && g < 1.0f) {
      const float w1 = (1.0f - g) * (1.0f - g) * (1.0f - g) * (1.0f - g), w2 = g * g * g * g;
      dpos = -db * w1 * inv_soft; dnegpos = -da * w2 * inv_soft;
    }
    float d = tau * dpos + use_neg * (-tau * dnegpos + dunl);
    if (apply_sigmoid) d = cl ? 0.f : d * p * q;
    grad[i] = d * scale;
  }
}

__global__ void __launch_bounds__(RED_THREADS) mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                                  double* __restrict__ part) {
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)RED_THREADS + threadIdx.x; i < n; i += (long long)RED_BLOCKS * RED_THREADS) {
    const float d = a[i] - b[i];
    s += (double)d * d;
  }
  __shared__ double sh[RED_THREADS];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__global__ void mse_final_kernel(const double* __restrict__ part, long long n, float* __restrict__ out) {
  double s = 0.0;
  for (int b = 0; b < RED_BLOCKS; ++b) s += part[b];
  out[0] = (float)(s / (double)n);
}

__global__ void __launch_bounds__(256) mse_grad_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float scale,
                                                       float* __restrict__ grad) {
  const float c = 2.0f / (float)n * scale;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) grad[i] = (a[i] - b[i]) * c;
}

// torch.optim.Adam (amsgrad off, maximize off): one fused pass over the flat bucket
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                   float wd, float bc1, float bc2_sqrt, float gscale) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float gi = g[i] * gscale;
    if (wd != 0.f) gi = fmaf(wd, p[i], gi);
    const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);            // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

}  // namespace
}  // namespace cetpick

using namespace cetpick;

extern "C" int cetpick_train_workspace_bytes(size_t* bytes) {
  if (!bytes) return CETPICK_ERR_BAD_ARG;
  *bytes = (size_t)RED_BLOCKS * NSUM * sizeof(double) + 64 * sizeof(float) + 256;
  return CETPICK_OK;
}

// loss.py:255-325 PULoss(tau)(pred, gt) with pred = _sigmoid(logits) when apply_sigmoid (the trainer applies it in place,
// tomo_cr_semi_trainer.py:53).  out4 (device): loss, positive risk, negative risk, number of positives (the reference
// raises ValueError when that is zero: callers check out4[3]).  grad (nullable, device [n]): grad_scale * d loss / d logits.
extern "C" int cetpick_pu_loss_f32(const float* logits, const float* gt, int64_t n, int apply_sigmoid, double tau, double beta,
                                   float* out4, float* grad, float grad_scale, void* ws, size_t ws_bytes, void* stream) {
  g_launches = 0;
  size_t need = 0;
  cetpick_train_workspace_bytes(&need);
  if (!logits || !gt || !out4 || n <= 0) return CETPICK_ERR_BAD_ARG;
  if (!ws || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 255)) return CETPICK_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* part = static_cast<double*>(ws);
  float* coef = reinterpret_cast<float*>(part + RED_BLOCKS * NSUM);
  pu_partial_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(logits, gt, n, apply_sigmoid, part);
  CETPICK_LAUNCH_CHECK();
  pu_final_kernel<<<1, 1, 0, st>>>(part, tau, beta, out4, coef);
  CETPICK_LAUNCH_CHECK();
  if (grad) {
    const int grid = (int)std::min<long long>(ceil_div<long long>(n, 256), (long long)num_sms() * 16);
    pu_grad_kernel<<<grid, 256, 0, st>>>(logits, gt, n, apply_sigmoid, coef, grad_scale, grad);
    CETPICK_LAUNCH_CHECK();
  }
  return CETPICK_OK;
}

// loss.py:701-715 ConsistencyLoss = F.mse_loss(a, b): out1 (device) and, optionally, grad_a = grad_scale * d loss / d a
extern "C" int cetpick_mse_loss_f32(const float* a, const float* b, int64_t n, float* out1, float* grad_a, float grad_scale,
                                    void* ws, size_t ws_bytes, void* stream) {
  g_launches = 0;
  size_t need = 0;
  cetpick_train_workspace_bytes(&need);
  if (!a || !b || !out1 || n <= 0) return CETPICK_ERR_BAD_ARG;
  if (!ws || ws_bytes < need || (reinterpret_cast<uintptr_t>(ws) & 255)) return CETPICK_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* part = static_cast<double*>(ws);
  mse_partial_kernel<<<RED_BLOCKS, RED_THREADS, 0, st>>>(a, b, n, part);
  CETPICK_LAUNCH_CHECK();
  mse_final_kernel<<<1, 1, 0, st>>>(part, n, out1);
  CETPICK_LAUNCH_CHECK();
  if (grad_a) {
    const int grid = (int)std::min<long long>(ceil_div<long long>(n, 256), (long long)num_sms() * 16);
    mse_grad_kernel<<<grid, 256, 0, st>>>(a, b, n, grad_scale, grad_a);
    CETPICK_LAUNCH_CHECK();
  }
  return CETPICK_OK;
}

// One torch.optim.Adam step (main.py:55 defaults: betas (0.9, 0.999), eps 1e-8, weight_decay 0) over a flat fp32 bucket;
// step = 1 for the first update; grads are multiplied by grad_scale first (1 / world size after a summed all-reduce).
extern "C" int cetpick_adam_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                     double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                                     double grad_scale, void* stream) {
  g_launches = 0;
  if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return CETPICK_ERR_BAD_ARG;
  const double bc1 = 1.0 - std::pow(beta1, (double)step), bc2 = 1.0 - std::pow(beta2, (double)step);
  const int grid = (int)std::min<long long>(ceil_div<long long>(n, 256), (long long)num_sms() * 16);
  adam_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, (float)lr, (float)beta1,
                                                                    (float)beta2, (float)eps, (float)weight_decay, (float)bc1,
                                                                    (float)std::sqrt(bc2), (float)grad_scale);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}
