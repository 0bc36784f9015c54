// Training-mode layers of the detector (BASELINE.json configs[4]): fp32 forward with batch-statistics
// BatchNorm and the backward of every layer kind the network has -- what `loss.backward()` of
// cet_pick/trains/base_trainer.py:135-155,484-489 runs through for cet_pick/models/networks/unet_small.py:63-97 and
// unet.py:198-249 (DownConv), :319-399 (UpConv), :861-886 (UNet.forward).
//
//   conv (2-D trunk 3x3 / 1x1 / stem 7x7 stride 2, 3-D head 3x3x3 dilated (1,4,4) and 3x1x1)   conv_f32_kernel
//   data gradient of a stride-1 conv = the same kernel on flipped, transposed weights          flip_weights_kernel
//   weight gradient (all of the above and the transposed conv)                                 wgrad_f32_kernel
//   ConvTranspose2d(k=2, s=2) forward; its data gradient is a stride-2 2x2 conv                upconv_f32_kernel
//   BatchNorm2d (batch statistics, running-stat update) [+ ReLU] forward / backward            bn_*_kernel
//   MaxPool2d(2, ceil_mode) forward / backward (first maximum wins, like ATen)                 pool_*_kernel
//   ReLU backward, per-channel sums (bias gradients)
//
// Tensors are fp32 with DENSE rows and explicit (slice, channel) strides, so one buffer serves the 2-D trunk
// ((D,C,h,w): slice = z) and the 3-D head ((C,D,h,w) read through the same strides): the reference's permutes
// (unet_small.py:71,83-84) and the channel concat (unet.py:390) are views here, not copies.
// This is the FIRST correct training path: CUDA-core fp32 convolutions, TF32 mma.sync weight gradient (switchable).  The
// tcgen05 kernels of the inference path are not used here (see DESIGN.md section 4.7 for accuracy and timing).
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace cetpick {
namespace {

using Geom = cetpick_conv_geom;

constexpr int CV_THREADS = 128;
constexpr int CV_CO = 16;      // output channels per thread
constexpr int CV_CI = 8;       // input channels per weight stage

// y[n][co][oy][ox] (+)= bias[co] + sum_{ci,tz,ty,tx} x[n + tz*dz - pz][ci][oy*s + ty*dy - py][ox*s + tx*dx - px] * w[co][ci][tz][ty][tx]
// (z taps stay inside the crop of `zdepth` slices that n belongs to)
__global__ void __launch_bounds__(CV_THREADS) conv_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ y,
                                                              const Geom g, const int flags) {
  extern __shared__ float s_w[];                       // [CV_CI][taps][CV_CO]
  const bool accumulate = flags & 1, relu = flags & 2;
  const int taps = g.kz * g.ky * g.kx;
  const int p = blockIdx.x * CV_THREADS + threadIdx.x;
  const int n = blockIdx.y, co0 = blockIdx.z * CV_CO;
  const bool valid = p < g.Ho * g.Wo;
  const int oy = valid ? p / g.Wo : 0, ox = valid ? p - oy * g.Wo : 0;
  const int crop0 = (n / g.zdepth) * g.zdepth;
  float acc[CV_CO];
#pragma unroll
  for (int c = 0; c < CV_CO; ++c) acc[c] = 0.f;
  for (int ci0 = 0; ci0 < g.Cin; ci0 += CV_CI) {
    __syncthreads();
    for (int i = threadIdx.x; i < CV_CI * taps * CV_CO; i += CV_THREADS) {
      const int co = i % CV_CO, t = (i / CV_CO) % taps, ci = i / (CV_CO * taps);
      float v = 0.f;
      if (ci0 + ci < g.Cin && co0 + co < g.Cout) v = w[((size_t)(co0 + co) * g.Cin + ci0 + ci) * taps + t];
      s_w[i] = v;
    }
    __syncthreads();
    if (!valid) continue;
    const int nci = min(CV_CI, g.Cin - ci0);
    for (int ci = 0; ci < nci; ++ci) {
      const float* xc = x + (size_t)(ci0 + ci) * g.xs_c;
      const float* wc = s_w + (size_t)ci * taps * CV_CO;
      for (int tz = 0; tz < g.kz; ++tz) {
        const int zin = n + tz * g.dz - g.pz;
        if (zin < crop0 || zin >= crop0 + g.zdepth) continue;
        for (int ty = 0; ty < g.ky; ++ty) {
          const int iy = oy * g.stride + ty * g.dy - g.py;
          if (iy < 0 || iy >= g.H) continue;
          const float* xr = xc + (size_t)zin * g.xs_n + (size_t)iy * g.W;
          const float* wr = wc + (size_t)((tz * g.ky + ty) * g.kx) * CV_CO;
          for (int tx = 0; tx < g.kx; ++tx) {
            const int ix = ox * g.stride + tx * g.dx - g.px;
            if (ix < 0 || ix >= g.W) continue;
            const float v = __ldg(xr + ix);
            const float4* w4 = reinterpret_cast<const float4*>(wr + tx * CV_CO);
#pragma unroll
            for (int q = 0; q < CV_CO / 4; ++q) {
              const float4 ww = w4[q];
              acc[4 * q + 0] = fmaf(v, ww.x, acc[4 * q + 0]);
              acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
              acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
              acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
            }
          }
        }
      }
    }
  }
  if (!valid) return;
#pragma unroll
  for (int c = 0; c < CV_CO; ++c) {
    if (co0 + c >= g.Cout) break;
    float* o = y + (size_t)n * g.ys_n + (size_t)(co0 + c) * g.ys_c + p;
    float v = acc[c] + (bias ? bias[co0 + c] : 0.f);
    if (accumulate) v += *o;
    *o = relu ? fmaxf(v, 0.f) : v;
  }
}


// The network's hot shapes -- (1,3,3) dilation 1 (trunk) and (3,3,3) dilation (1,4,4) (head), stride 1 -- with the taps
// unrolled: each thread owns TWO horizontally adjacent output pixels x CV_CO output channels, loads each input row
// segment once for both pixels and all three x taps, and reads each weight vector once for both pixels.
template <int KZ, int DXY>
__global__ void __launch_bounds__(CV_THREADS) conv3x3_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, float* __restrict__ y,
                                                                 const Geom g, const int flags) {
  constexpr int TAPS = KZ * 9;
  constexpr int NCOL = 2 * DXY + 2;                    // input columns the two pixels touch: ox - DXY .. ox + 1 + DXY
  __shared__ __align__(16) float s_w[CV_CI * TAPS * CV_CO];
  const bool accumulate = flags & 1, relu = flags & 2;
  const int pairs_per_row = (g.Wo + 1) >> 1;
  const int q = blockIdx.x * CV_THREADS + threadIdx.x;
  const int n = blockIdx.y, co0 = blockIdx.z * CV_CO;
  const bool valid = q < g.Ho * pairs_per_row;
  const int oy = valid ? q / pairs_per_row : 0, ox = valid ? 2 * (q - oy * pairs_per_row) : 0;
  const bool valid1 = valid && (ox + 1 < g.Wo);
  const int crop0 = (n / g.zdepth) * g.zdepth;
  float acc0[CV_CO], acc1[CV_CO];
#pragma unroll
  for (int c = 0; c < CV_CO; ++c) acc0[c] = acc1[c] = 0.f;
  // validity of the rows / columns this thread reads (zero padding = skipped loads)
  bool zok[KZ], yok[3], xok[NCOL];
#pragma unroll
  for (int tz = 0; tz < KZ; ++tz) { const int zin = n + tz - g.pz; zok[tz] = (zin >= crop0) && (zin < crop0 + g.zdepth); }
#pragma unroll
  for (int ty = 0; ty < 3; ++ty) { const int iy = oy + ty * DXY - g.py; yok[ty] = (iy >= 0) && (iy < g.H); }
#pragma unroll
  for (int c = 0; c < NCOL; ++c) { const int ix = ox - g.px + c; xok[c] = valid && (ix >= 0) && (ix < g.W); }
  for (int ci0 = 0; ci0 < g.Cin; ci0 += CV_CI) {
    __syncthreads();
    for (int i = threadIdx.x; i < CV_CI * TAPS * CV_CO; i += CV_THREADS) {
      const int co = i % CV_CO, t = (i / CV_CO) % TAPS, ci = i / (CV_CO * TAPS);
      float v = 0.f;
      if (ci0 + ci < g.Cin && co0 + co < g.Cout) v = w[((size_t)(co0 + co) * g.Cin + ci0 + ci) * TAPS + t];
      s_w[i] = v;
    }
    __syncthreads();
    if (!valid) continue;
    const int nci = min(CV_CI, g.Cin - ci0);
    for (int ci = 0; ci < nci; ++ci) {
      const float* xc = x + (size_t)(ci0 + ci) * g.xs_c;
      const float* wc = s_w + (size_t)ci * TAPS * CV_CO;
#pragma unroll
      for (int tz = 0; tz < KZ; ++tz) {
        if (!zok[tz]) continue;
        const float* xz = xc + (size_t)(n + tz - g.pz) * g.xs_n;
#pragma unroll
        for (int ty = 0; ty < 3; ++ty) {
          if (!yok[ty]) continue;
          const float* xr = xz + (size_t)(oy + ty * DXY - g.py) * g.W + (ox - g.px);
          float v[NCOL];
#pragma unroll
          for (int c = 0; c < NCOL; ++c) v[c] = 0.f;
#pragma unroll
          for (int tx = 0; tx < 3; ++tx) {                 // only the columns a tap lands on (dilation 4: 6 of 10)
            if (xok[tx * DXY]) v[tx * DXY] = __ldg(xr + tx * DXY);
            if (xok[tx * DXY + 1]) v[tx * DXY + 1] = __ldg(xr + tx * DXY + 1);
          }
#pragma unroll
          for (int tx = 0; tx < 3; ++tx) {
            const float a = v[tx * DXY], b = v[tx * DXY + 1];
            const float4* w4 = reinterpret_cast<const float4*>(wc + ((tz * 3 + ty) * 3 + tx) * CV_CO);
#pragma unroll
            for (int k = 0; k < CV_CO / 4; ++k) {
              const float4 ww = w4[k];
              acc0[4 * k + 0] = fmaf(a, ww.x, acc0[4 * k + 0]); acc1[4 * k + 0] = fmaf(b, ww.x, acc1[4 * k + 0]);
              acc0[4 * k + 1] = fmaf(a, ww.y, acc0[4 * k + 1]); acc1[4 * k + 1] = fmaf(b, ww.y, acc1[4 * k + 1]);
              acc0[4 * k + 2] = fmaf(a, ww.z, acc0[4 * k + 2]); acc1[4 * k + 2] = fmaf(b, ww.z, acc1[4 * k + 2]);
              acc0[4 * k + 3] = fmaf(a, ww.w, acc0[4 * k + 3]); acc1[4 * k + 3] = fmaf(b, ww.w, acc1[4 * k + 3]);
            }
          }
        }
      }
    }
  }
  if (!valid) return;
#pragma unroll
  for (int c = 0; c < CV_CO; ++c) {
    if (co0 + c >= g.Cout) break;
    float* o = y + (size_t)n * g.ys_n + (size_t)(co0 + c) * g.ys_c + (size_t)oy * g.Wo + ox;
    const float bv = bias ? bias[co0 + c] : 0.f;
    float v0 = acc0[c] + bv, v1 = acc1[c] + bv;
    if (accumulate) { v0 += o[0]; if (valid1) v1 += o[1]; }
    o[0] = relu ? fmaxf(v0, 0.f) : v0;
    if (valid1) o[1] = relu ? fmaxf(v1, 0.f) : v1;
  }
}

// wt[ci][co][taps-1-t] = w[co][ci][t]: the weights with which a stride-1 conv of dy gives dx
__global__ void flip_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int Cin, int taps) {
  const size_t total = (size_t)Cout * Cin * taps;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % taps), ci = (int)((i / taps) % Cin), co = (int)(i / ((size_t)taps * Cin));
    wt[((size_t)ci * Cout + co) * taps + (taps - 1 - t)] = w[i];
  }
}

// dw[co][ci][t] += sum_{n,oy,ox} dy[n][co][oy][ox] * x[n + tz*dz - pz][ci][oy*s + ty*dy - py][ox*s + tx*dx - px]
// One CTA: CO_T output channels x CI_T input channels, all taps, in registers, over a chunk of the (n, pixel) range;
// every loaded x value feeds CO_T accumulators, every dy value CI_T * TAPS; CTA reduction, atomics.
constexpr int WG_THREADS = 256;
template <int KZ, int KY, int KX, int CO_T, int CI_T>
__global__ void __launch_bounds__(WG_THREADS) wgrad_f32_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dw, const Geom g, const long long chunk) {
  constexpr int TAPS = KZ * KY * KX;
  constexpr int NACC = CO_T * CI_T * TAPS;
  const int ncig = ceil_div(g.Cin, CI_T);
  const int co0 = (blockIdx.x / ncig) * CO_T, ci0 = (blockIdx.x % ncig) * CI_T;
  const long long hw = (long long)g.Ho * g.Wo, total = (long long)g.N * hw;
  const long long i0 = (long long)blockIdx.y * chunk, i1 = min(total, i0 + chunk);
  float acc[CO_T][CI_T][TAPS];
#pragma unroll
  for (int o = 0; o < CO_T; ++o)
#pragma unroll
    for (int c = 0; c < CI_T; ++c)
#pragma unroll
      for (int t = 0; t < TAPS; ++t) acc[o][c][t] = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += WG_THREADS) {
    const int n = (int)(i / hw), p = (int)(i - (long long)n * hw);
    const int oy = p / g.Wo, ox = p - oy * g.Wo;
    float d[CO_T];
    bool any = false;
#pragma unroll
    for (int o = 0; o < CO_T; ++o) {
      d[o] = (co0 + o < g.Cout) ? __ldg(dy + (size_t)n * g.ys_n + (size_t)(co0 + o) * g.ys_c + p) : 0.f;
      any |= (d[o] != 0.f);
    }
    if (!any) continue;                                     // ReLU-masked gradients are exact zeros
    const int crop0 = (n / g.zdepth) * g.zdepth;
#pragma unroll
    for (int tz = 0; tz < KZ; ++tz) {
      const int zin = n + tz * g.dz - g.pz;
      if (zin < crop0 || zin >= crop0 + g.zdepth) continue;
#pragma unroll
      for (int ty = 0; ty < KY; ++ty) {
        const int iy = oy * g.stride + ty * g.dy - g.py;
        if (iy < 0 || iy >= g.H) continue;
        const float* xr = x + (size_t)zin * g.xs_n + (size_t)iy * g.W;
#pragma unroll
        for (int tx = 0; tx < KX; ++tx) {
          const int ix = ox * g.stride + tx * g.dx - g.px;
          if (ix < 0 || ix >= g.W) continue;
#pragma unroll
          for (int c = 0; c < CI_T; ++c) {
            if (ci0 + c >= g.Cin) continue;
            const float xv = __ldg(xr + (size_t)(ci0 + c) * g.xs_c + ix);
#pragma unroll
            for (int o = 0; o < CO_T; ++o) acc[o][c][(tz * KY + ty) * KX + tx] = fmaf(d[o], xv, acc[o][c][(tz * KY + ty) * KX + tx]);
          }
        }
      }
    }
  }
  __shared__ float s_red[WG_THREADS / 32][NACC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 0; o < CO_T; ++o)
#pragma unroll
    for (int c = 0; c < CI_T; ++c)
#pragma unroll
      for (int t = 0; t < TAPS; ++t) {
        float v = acc[o][c][t];
#pragma unroll
        for (int k = 16; k > 0; k >>= 1) v += __shfl_xor_sync(0xffffffffu, v, k);
        if (lane == 0) s_red[warp][(o * CI_T + c) * TAPS + t] = v;
      }
  __syncthreads();
  for (int i = threadIdx.x; i < NACC; i += WG_THREADS) {
    const int t = i % TAPS, c = (i / TAPS) % CI_T, o = i / (TAPS * CI_T);
    if (ci0 + c >= g.Cin || co0 + o >= g.Cout) continue;
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < WG_THREADS / 32; ++k) v += s_red[k][i];
    if (v != 0.f) atomicAdd(dw + ((size_t)(co0 + o) * g.Cin + ci0 + c) * TAPS + t, v);
  }
}

// ConvTranspose2d(k=2, s=2) + bias, cropped to (Ho, Wo) <= (2H, 2W) (unet.py:285-292): weights [Cin][Cout][2][2]
__global__ void __launch_bounds__(CV_THREADS) upconv_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                const float* __restrict__ bias, float* __restrict__ y,
                                                                const Geom g) {
  extern __shared__ float s_w[];                       // [CV_CI][4][CV_CO]
  const int p = blockIdx.x * CV_THREADS + threadIdx.x;
  const int n = blockIdx.y, co0 = blockIdx.z * CV_CO;
  const bool valid = p < g.Ho * g.Wo;
  const int oy = valid ? p / g.Wo : 0, ox = valid ? p - oy * g.Wo : 0;
  const int iy = oy >> 1, ix = ox >> 1, tap = (oy & 1) * 2 + (ox & 1);
  float acc[CV_CO];
#pragma unroll
  for (int c = 0; c < CV_CO; ++c) acc[c] = 0.f;
  for (int ci0 = 0; ci0 < g.Cin; ci0 += CV_CI) {
    __syncthreads();
    for (int i = threadIdx.x; i < CV_CI * 4 * CV_CO; i += CV_THREADS) {
      const int co = i % CV_CO, t = (i / CV_CO) % 4, ci = i / (CV_CO * 4);
      float v = 0.f;
      if (ci0 + ci < g.Cin && co0 + co < g.Cout) v = w[((size_t)(ci0 + ci) * g.Cout + co0 + co) * 4 + t];
      s_w[i] = v;
    }
    __syncthreads();
    if (!valid) continue;
    const int nci = min(CV_CI, g.Cin - ci0);
    for (int ci = 0; ci < nci; ++ci) {
      const float v = __ldg(x + (size_t)n * g.xs_n + (size_t)(ci0 + ci) * g.xs_c + (size_t)iy * g.W + ix);
      const float4* w4 = reinterpret_cast<const float4*>(s_w + (size_t)(ci * 4 + tap) * CV_CO);
#pragma unroll
      for (int q = 0; q < CV_CO / 4; ++q) {
        const float4 ww = w4[q];
        acc[4 * q + 0] = fmaf(v, ww.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(v, ww.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(v, ww.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(v, ww.w, acc[4 * q + 3]);
      }
    }
  }
  if (!valid) return;
#pragma unroll
  for (int c = 0; c < CV_CO; ++c) {
    if (co0 + c >= g.Cout) break;
    y[(size_t)n * g.ys_n + (size_t)(co0 + c) * g.ys_c + p] = acc[c] + (bias ? bias[co0 + c] : 0.f);
  }
}

// ------------------------------------------------------------------------------------------------ per-channel sums
// part[c][split][k]: k = 0: sum a, 1: sum a*b (b optional).  Rows are dense (HW contiguous), (slice, channel) strided.
constexpr int RED_THREADS = 256, RED_SPLITS = 32;

// MODE 0: a = x, b = x (sum, sum of squares)           -- BatchNorm statistics
// MODE 1: a = g = dy * (relu ? y > 0 : 1), b = xhat    -- BatchNorm backward sums (sum g, sum g * xhat)
// MODE 2: a = x                                        -- bias gradient
template <int MODE>
__global__ void __launch_bounds__(RED_THREADS) chan_reduce_kernel(const float* __restrict__ x, long long xs_n, long long xs_c,
                                                                  const float* __restrict__ y, const float* __restrict__ dy,
                                                                  long long ys_n, long long ys_c,
                                                                  const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                  int N, int HW, int relu, double* __restrict__ part) {
  const int c = blockIdx.x, sp = blockIdx.y;
  const long long total = (long long)N * HW;
  const long long chunk = ceil_div<long long>(total, RED_SPLITS);
  const long long i0 = sp * chunk, i1 = min(total, i0 + chunk);
  double s0 = 0.0, s1 = 0.0;
  const float mu = (MODE == 1) ? mean[c] : 0.f, is = (MODE == 1) ? invstd[c] : 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += RED_THREADS) {
    const int n = (int)(i / HW), p = (int)(i - (long long)n * HW);
    if (MODE == 0) {
      const float v = x[(size_t)n * xs_n + (size_t)c * xs_c + p];
      s0 += v; s1 += (double)v * v;
    } else if (MODE == 1) {
      const size_t o = (size_t)n * ys_n + (size_t)c * ys_c + p;
      float gv = dy[o];
      if (relu && !(y[o] > 0.f)) gv = 0.f;
      const float xh = (x[(size_t)n * xs_n + (size_t)c * xs_c + p] - mu) * is;
      s0 += gv; s1 += (double)gv * xh;
    } else {
      s0 += x[(size_t)n * xs_n + (size_t)c * xs_c + p];
    }
  }
  __shared__ double sh[2][RED_THREADS];
  sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1;
  __syncthreads();
  for (int o = RED_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[((size_t)c * RED_SPLITS + sp) * 2] = sh[0][0]; part[((size_t)c * RED_SPLITS + sp) * 2 + 1] = sh[1][0]; }
}

// BatchNorm2d training statistics (F.batch_norm(training=True)): biased variance for the normalisation, unbiased for
// the running estimate, running <- (1 - momentum) * running + momentum * batch
__global__ void bn_stats_final_kernel(const double* __restrict__ part, int C, double count, float eps, float momentum,
                                      float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                      float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int k = 0; k < RED_SPLITS; ++k) { s0 += part[((size_t)c * RED_SPLITS + k) * 2]; s1 += part[((size_t)c * RED_SPLITS + k) * 2 + 1]; }
  const double mu = s0 / count;
  const double var = fmax(s1 / count - mu * mu, 0.0);
  save_mean[c] = (float)mu;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mu);
  if (running_var) running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * var * (count > 1 ? count / (count - 1.0) : 1.0));
}

// sums[c] = {sum g, sum g*xhat}; dgamma += sum g*xhat, dbeta += sum g
__global__ void bn_bwd_final_kernel(const double* __restrict__ part, int C, float* __restrict__ sums,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int k = 0; k < RED_SPLITS; ++k) { s0 += part[((size_t)c * RED_SPLITS + k) * 2]; s1 += part[((size_t)c * RED_SPLITS + k) * 2 + 1]; }
  sums[2 * c] = (float)s0; sums[2 * c + 1] = (float)s1;
  if (dgamma) dgamma[c] += (float)s1;
  if (dbeta) dbeta[c] += (float)s0;
}

__global__ void chan_sum_final_kernel(const double* __restrict__ part, int C, float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0;
  for (int k = 0; k < RED_SPLITS; ++k) s0 += part[((size_t)c * RED_SPLITS + k) * 2];
  out[c] = accumulate ? out[c] + (float)s0 : (float)s0;
}

// y = [relu](gamma * (x - mean) * invstd + beta)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, long long xs_n, long long xs_c,
                                                       float* __restrict__ y, long long ys_n, long long ys_c,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ mean, const float* __restrict__ invstd,
                                                       int C, int HW, int relu) {
  const int n = blockIdx.z, c = blockIdx.y;
  const float a = gamma[c] * invstd[c], b = beta[c] - mean[c] * a;
  const float* xp = x + (size_t)n * xs_n + (size_t)c * xs_c;
  float* yp = y + (size_t)n * ys_n + (size_t)c * ys_c;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
    float v = fmaf(xp[p], a, b);
    if (relu) v = fmaxf(v, 0.f);
    yp[p] = v;
  }
  (void)C;
}

// dx = gamma * invstd * (g - sum_g / M - xhat * sum_gx / M),  g = dy * (relu ? y > 0 : 1)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const float* __restrict__ x, long long xs_n, long long xs_c,
                                                           const float* __restrict__ y, const float* __restrict__ dy,
                                                           long long ys_n, long long ys_c, float* __restrict__ dx,
                                                           long long dxs_n, long long dxs_c, const float* __restrict__ gamma,
                                                           const float* __restrict__ mean, const float* __restrict__ invstd,
                                                           const float* __restrict__ sums, float inv_count, int HW, int relu) {
  const int n = blockIdx.z, c = blockIdx.y;
  const float mu = mean[c], is = invstd[c], k = gamma[c] * is;
  const float m0 = sums[2 * c] * inv_count, m1 = sums[2 * c + 1] * inv_count;
  const float* xp = x + (size_t)n * xs_n + (size_t)c * xs_c;
  const float* yp = y + (size_t)n * ys_n + (size_t)c * ys_c;
  const float* gp = dy + (size_t)n * ys_n + (size_t)c * ys_c;
  float* dp = dx + (size_t)n * dxs_n + (size_t)c * dxs_c;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
    float gv = gp[p];
    if (relu && !(yp[p] > 0.f)) gv = 0.f;
    const float xh = (xp[p] - mu) * is;
    dp[p] = k * (gv - m0 - xh * m1);
  }
}

// MaxPool2d(2, ceil_mode=True) (unet.py:225): windows clipped at the border
__global__ void __launch_bounds__(256) pool_f32_kernel(const float* __restrict__ x, long long xs_n, long long xs_c,
                                                       float* __restrict__ y, long long ys_n, long long ys_c, int H, int W,
                                                       int Ho, int Wo) {
  const int n = blockIdx.z, c = blockIdx.y;
  const float* xp = x + (size_t)n * xs_n + (size_t)c * xs_c;
  float* yp = y + (size_t)n * ys_n + (size_t)c * ys_c;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < Ho * Wo; p += gridDim.x * 256) {
    const int oy = p / Wo, ox = p - oy * Wo;
    float m = -INFINITY;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const int iy = 2 * oy + a, ix = 2 * ox + b;
        if (iy < H && ix < W) { const float v = xp[(size_t)iy * W + ix]; if (v > m || v != v) m = v; }
      }
    yp[p] = m;
  }
}

// dx[iy][ix] (+)= dy[iy/2][ix/2] if (iy, ix) is the FIRST maximum of its window in scan order (ATen's argmax rule)
__global__ void __launch_bounds__(256) pool_bwd_f32_kernel(const float* __restrict__ x, long long xs_n, long long xs_c,
                                                           const float* __restrict__ dy, long long dys_n, long long dys_c,
                                                           float* __restrict__ dx, long long dxs_n, long long dxs_c, int H, int W,
                                                           int Ho, int Wo, int accumulate) {
  const int n = blockIdx.z, c = blockIdx.y;
  const float* xp = x + (size_t)n * xs_n + (size_t)c * xs_c;
  const float* gp = dy + (size_t)n * dys_n + (size_t)c * dys_c;
  float* dp = dx + (size_t)n * dxs_n + (size_t)c * dxs_c;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < H * W; p += gridDim.x * 256) {
    const int iy = p / W, ix = p - iy * W;
    const int oy = iy >> 1, ox = ix >> 1;
    float m = -INFINITY;
    int arg = -1;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const int yy = 2 * oy + a, xx = 2 * ox + b;
        if (yy < H && xx < W) { const float v = xp[(size_t)yy * W + xx]; if (v > m || v != v) { m = v; arg = yy * W + xx; } }
      }
    const float gv = (arg == p) ? gp[(size_t)oy * Wo + ox] : 0.f;
    dp[p] = accumulate ? dp[p] + gv : gv;
  }
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                       float* __restrict__ dx, size_t n) {
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) dx[i] = (y[i] > 0.f) ? dy[i] : 0.f;
}


// Weight gradient as a tiled SGEMM with implicit im2col: dW[co][(tap, ci)] = sum_i dy[co][i] * X[(tap, ci)][i], i running
// over the (slice, pixel) positions.  CTA tile BM output channels x BN (tap, ci) columns, K chunks of 32 positions staged in
// shared memory (position-major rows, so a thread reads its 4 + 4 operands of a k step with two 16-byte loads), 4 x 4
// accumulators per thread, split over the positions (blockIdx.z), fp32 atomics at the end.
constexpr int GK = 32, G_THREADS = 256;
template <int BM, int BN>
__global__ void __launch_bounds__(G_THREADS) wgrad_gemm_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dw, const Geom g, const long long chunk) {
  static_assert(BM * BN == 4096, "256 threads x (4 x 4)");
  constexpr int A_PER = GK * BM / G_THREADS, B_PER = GK * BN / G_THREADS;
  __shared__ __align__(16) float As[GK][BM + 4];
  __shared__ __align__(16) float Bs[GK][BN + 4];
  const int taps = g.kz * g.ky * g.kx, ntot = taps * g.Cin;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const long long hw = (long long)g.Ho * g.Wo, total = (long long)g.N * hw;
  const long long i0 = (long long)blockIdx.z * chunk, i1 = min(total, i0 + chunk);
  const int kk = threadIdx.x & 31, r0 = threadIdx.x >> 5;          // this thread stages position kk of every chunk
  const int tm = threadIdx.x % (BM / 4), tn = threadIdx.x / (BM / 4);
  // the B columns this thread stages: column -> (tap, ci) -> channel offset and tap displacement (fixed for the kernel)
  long long coff[B_PER];
  int cdz[B_PER], cdy[B_PER], cdx[B_PER];
#pragma unroll
  for (int j = 0; j < B_PER; ++j) {
    const int col = n0 + r0 + 8 * j;
    if (col < ntot) {
      const int tap = col / g.Cin, ci = col - tap * g.Cin;
      const int tz = tap / (g.ky * g.kx), ty = (tap / g.kx) % g.ky, tx = tap % g.kx;
      coff[j] = (long long)ci * g.xs_c;
      cdz[j] = tz * g.dz - g.pz; cdy[j] = ty * g.dy - g.py; cdx[j] = tx * g.dx - g.px;
    } else {
      coff[j] = -1; cdz[j] = cdy[j] = cdx[j] = 0;
    }
  }
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (long long k0 = i0; k0 < i1; k0 += GK) {
    const long long i = k0 + kk;
    const bool live = i < i1;
    const int n = live ? (int)(i / hw) : 0, p = live ? (int)(i - (long long)n * hw) : 0;
    const int oy = p / g.Wo, ox = p - oy * g.Wo;
    const int crop0 = (n / g.zdepth) * g.zdepth;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      const int mm = r0 + 8 * j, co = m0 + mm;
      As[kk][mm] = (live && co < g.Cout) ? __ldg(dy + (size_t)n * g.ys_n + (size_t)co * g.ys_c + p) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
      float v = 0.f;
      if (live && coff[j] >= 0) {
        const int zin = n + cdz[j], iy = oy * g.stride + cdy[j], ix = ox * g.stride + cdx[j];
        if (zin >= crop0 && zin < crop0 + g.zdepth && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
          v = __ldg(x + (size_t)zin * g.xs_n + coff[j] + (size_t)iy * g.W + ix);
      }
      Bs[kk][r0 + 8 * j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
      acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]); acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
      acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]); acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
      acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]); acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
      acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]); acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int co = m0 + tm * 4 + a;
    if (co >= g.Cout) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int col = n0 + tn * 4 + b;
      if (col >= ntot || acc[a][b] == 0.f) continue;
      const int tap = col / g.Cin, ci = col - tap * g.Cin;
      atomicAdd(dw + ((size_t)co * g.Cin + ci) * taps + tap, acc[a][b]);
    }
  }
}

template <int BM, int BN>
int launch_wgrad_gemm(const float* x, const float* dy, float* dw, const Geom& g, cudaStream_t s) {
  const long long total = (long long)g.N * g.Ho * g.Wo;
  const int ntot = g.kz * g.ky * g.kx * g.Cin;
  const int tiles = ceil_div(ntot, BN) * ceil_div(g.Cout, BM);
  // a few waves of CTAs; chunks of at least 2048 positions (64 k steps of 32) per CTA
  long long splits = std::max<long long>(1, std::min<long long>(ceil_div<long long>(total, 2048), ceil_div<long long>(6LL * num_sms(), tiles)));
  splits = std::min<long long>(splits, 65535);
  long long chunk = ceil_div<long long>(total, splits);
  chunk = ceil_div<long long>(chunk, GK) * GK;
  splits = ceil_div<long long>(total, chunk);
  wgrad_gemm_kernel<BM, BN><<<dim3(ceil_div(ntot, BN), ceil_div(g.Cout, BM), (unsigned)splits), G_THREADS, 0, s>>>(x, dy, dw, g, chunk);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

int g_train_tf32 = 1;      // 1: tensor-core (TF32) contraction in the wide layers (3x3 trunk convs, weight gradients); 0: fp32 FMA everywhere

// ---------------------------------------------------------------------------------------------- TF32 tensor-core variant
// Same tiling and implicit-im2col staging as wgrad_gemm_kernel, but the 32-position chunk is contracted with
// mma.sync.m16n8k8 TF32 (fp32 accumulate): operands are rounded to TF32 (cvt.rna) when they are staged -- the arithmetic
// class of the reference's own training (PyTorch's default cudnn.allow_tf32 = True).  8 warps: 2 x 4 (BM = 64) or 1 x 8
// (BM = 32) warp tiles of 32 x 16; shared-memory rows are padded to a stride of 8 mod 32 words, so the fragment loads
// (lane = 4 g + t reads row t / t + 4, column g / g + 8) hit 32 different banks.
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int BM, int BN>
__global__ void __launch_bounds__(G_THREADS) wgrad_mma_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                              float* __restrict__ dw, const Geom g, const long long chunk) {
  static_assert(BM * BN == 4096 && (BM == 32 || BM == 64), "8 warps x (32 x 16)");
  constexpr int A_PER = GK * BM / G_THREADS, B_PER = GK * BN / G_THREADS;
  constexpr int LDA = BM + 8, LDB = BN + 8;
  __shared__ __align__(16) uint32_t As[GK][LDA];
  __shared__ __align__(16) uint32_t Bs[GK][LDB];
  const int taps = g.kz * g.ky * g.kx, ntot = taps * g.Cin;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const long long hw = (long long)g.Ho * g.Wo, total = (long long)g.N * hw;
  const long long i0 = (long long)blockIdx.z * chunk, i1 = min(total, i0 + chunk);
  const int kk = threadIdx.x & 31, r0 = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  constexpr int WM = BM / 32;                          // warps along m
  const int mw = (warp % WM) * 32, nw = (warp / WM) * 16;
  // per staged column: channel offset (elements; < 0 = no such column) and the tap displacement packed as three biased bytes
  int coff[B_PER], dpk[B_PER];
#pragma unroll
  for (int j = 0; j < B_PER; ++j) {
    const int col = n0 + r0 + 8 * j;
    if (col < ntot) {
      const int tap = col / g.Cin, ci = col - tap * g.Cin;
      const int tz = tap / (g.ky * g.kx), ty = (tap / g.kx) % g.ky, tx = tap % g.kx;
      coff[j] = (int)((long long)ci * g.xs_c);
      dpk[j] = (tz * g.dz - g.pz + 64) | ((ty * g.dy - g.py + 64) << 8) | ((tx * g.dx - g.px + 64) << 16);
    } else {
      coff[j] = -1; dpk[j] = 0;
    }
  }
  float acc[2][2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
  // software pipeline: the global loads of chunk k + 1 are in flight while chunk k is contracted
  float ra[A_PER], rb[B_PER];
  auto fetch = [&](long long k0) {
    const long long i = k0 + kk;
    const bool live = i < i1;
    const int n = live ? (int)(i / hw) : 0, p = live ? (int)(i - (long long)n * hw) : 0;
    const int oy = p / g.Wo, ox = p - oy * g.Wo;
    const int crop0 = (n / g.zdepth) * g.zdepth;
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      const int co = m0 + r0 + 8 * j;
      ra[j] = (live && co < g.Cout) ? __ldg(dy + (size_t)n * g.ys_n + (size_t)co * g.ys_c + p) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
      float v = 0.f;
      if (live && coff[j] >= 0) {
        const int zin = n + (dpk[j] & 0xff) - 64, iy = oy * g.stride + ((dpk[j] >> 8) & 0xff) - 64;
        const int ix = ox * g.stride + ((dpk[j] >> 16) & 0xff) - 64;
        if (zin >= crop0 && zin < crop0 + g.zdepth && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
          v = __ldg(x + (size_t)zin * g.xs_n + (size_t)coff[j] + (size_t)iy * g.W + ix);
      }
      rb[j] = v;
    }
  };
  if (i0 < i1) fetch(i0);
  for (long long k0 = i0; k0 < i1; k0 += GK) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < A_PER; ++j) As[kk][r0 + 8 * j] = to_tf32(ra[j]);
#pragma unroll
    for (int j = 0; j < B_PER; ++j) Bs[kk][r0 + 8 * j] = to_tf32(rb[j]);
    __syncthreads();
    if (k0 + GK < i1) fetch(k0 + GK);
#pragma unroll
    for (int k8 = 0; k8 < GK; k8 += 8) {
      uint32_t af[2][4], bf[2][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int m = mw + mi * 16 + gq;
        af[mi][0] = As[k8 + tq][m];     af[mi][1] = As[k8 + tq][m + 8];
        af[mi][2] = As[k8 + tq + 4][m]; af[mi][3] = As[k8 + tq + 4][m + 8];
      }
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) {
        const int nn = nw + ni * 8 + gq;
        bf[ni][0] = Bs[k8 + tq][nn]; bf[ni][1] = Bs[k8 + tq + 4][nn];
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) mma_tf32(acc[mi][ni], af[mi], bf[ni]);
    }
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int co = m0 + mw + mi * 16 + gq + ((c & 2) ? 8 : 0);
        const int col = n0 + nw + ni * 8 + 2 * tq + (c & 1);
        const float v = acc[mi][ni][c];
        if (co >= g.Cout || col >= ntot || v == 0.f) continue;
        const int tap = col / g.Cin, ci = col - tap * g.Cin;
        atomicAdd(dw + ((size_t)co * g.Cin + ci) * taps + tap, v);
      }
}

// 3 x 3 'same' convolution (2-D trunk layers; also their data gradient on flipped weights) on the tensor cores:
// CTA = 64 output channels x one 8 x 16 pixel tile of one slice, K loop over chunks of 8 input channels.  Per chunk the
// 10 x 18 input patch (halo included, zero padding applied while staging) and the 9 x 8 x 64 weights are staged as TF32;
// each of the 9 taps is then 2 x 4 mma.sync.m16n8k8 per warp (warp tile 32 channels x 32 pixels = 2 tile rows), the B
// fragments being the patch shifted by the tap -- no im2col copy.  Next chunk's global loads are in flight meanwhile.
constexpr int CM_TW = 16, CM_TH = 8, CM_CI = 8, CM_CO = 64;
constexpr int CM_PLANE = 184;                      // 10 x 18 patch, padded: 184 = 24 mod 32 words -> conflict-free B fragments
constexpr int CM_LDW = CM_CO + 8;                  // 72 = 8 mod 32 words -> conflict-free A fragments
__global__ void __launch_bounds__(256, 2) conv3x3_mma_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, float* __restrict__ y,
                                                          const Geom g, const int flags) {
  __shared__ __align__(16) uint32_t s_p[CM_CI][CM_PLANE];
  __shared__ __align__(16) uint32_t s_w[9][CM_CI][CM_LDW];
  const bool accumulate = flags & 1, relu = flags & 2;
  const int tiles_x = ceil_div(g.W, CM_TW);
  const int tx0 = (blockIdx.x % tiles_x) * CM_TW, ty0 = (blockIdx.x / tiles_x) * CM_TH;
  const int n = blockIdx.y, co0 = blockIdx.z * CM_CO;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gq = lane >> 2, tq = lane & 3;
  const int mw = (warp & 1) * 32, wn = warp >> 1;          // warp: channels [mw, mw + 32), tile rows 2 wn, 2 wn + 1
  constexpr int P_PER = (CM_CI * 180 + 255) / 256;         // 6 patch elements per thread
  constexpr int W_PER = 9 * CM_CI * CM_CO / 256;           // 18 weights per thread
  float rp[P_PER], rw[W_PER];
  auto fetch = [&](int ci0) {
#pragma unroll
    for (int j = 0; j < P_PER; ++j) {
      const int e = threadIdx.x + 256 * j;
      float v = 0.f;
      if (e < CM_CI * 180) {
        const int ci = e / 180, r = e - ci * 180, py = r / 18, px = r - py * 18;
        const int iy = ty0 + py - 1, ix = tx0 + px - 1;
        if (ci0 + ci < g.Cin && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
          v = __ldg(x + (size_t)n * g.xs_n + (size_t)(ci0 + ci) * g.xs_c + (size_t)iy * g.W + ix);
      }
      rp[j] = v;
    }
#pragma unroll
    for (int j = 0; j < W_PER; ++j) {
      const int e = threadIdx.x + 256 * j;                 // co-major, then the 72 contiguous (ci, tap) weights of the chunk
      const int co = e / 72, r = e - co * 72, ci = r / 9;
      rw[j] = (co0 + co < g.Cout && ci0 + ci < g.Cin) ? __ldg(w + ((size_t)(co0 + co) * g.Cin + ci0) * 9 + r) : 0.f;
    }
  };
  float acc[2][4][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
  fetch(0);
  for (int ci0 = 0; ci0 < g.Cin; ci0 += CM_CI) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < P_PER; ++j) {
      const int e = threadIdx.x + 256 * j;
      if (e < CM_CI * 180) s_p[e / 180][e % 180] = to_tf32(rp[j]);
    }
#pragma unroll
    for (int j = 0; j < W_PER; ++j) {
      const int e = threadIdx.x + 256 * j;
      const int co = e / 72, r = e - co * 72, ci = r / 9, tap = r - ci * 9;
      s_w[tap][ci][co] = to_tf32(rw[j]);
    }
    __syncthreads();
    if (ci0 + CM_CI < g.Cin) fetch(ci0 + CM_CI);
#pragma unroll 1
    for (int ty = 0; ty < 3; ++ty)
#pragma unroll
    for (int tx = 0; tx < 3; ++tx) {
      const int tap = ty * 3 + tx;
      uint32_t af[2][4], bf[4][2];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        const int m = mw + mi * 16 + gq;
        af[mi][0] = s_w[tap][tq][m];     af[mi][1] = s_w[tap][tq][m + 8];
        af[mi][2] = s_w[tap][tq + 4][m]; af[mi][3] = s_w[tap][tq + 4][m + 8];
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int prow = 2 * wn + (ni >> 1) + ty, pcol = (ni & 1) * 8 + gq + tx;
        bf[ni][0] = s_p[tq][prow * 18 + pcol]; bf[ni][1] = s_p[tq + 4][prow * 18 + pcol];
      }
#pragma unroll
      for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) mma_tf32(acc[mi][ni], af[mi], bf[ni]);
    }
  }
#pragma unroll
  for (int mi = 0; mi < 2; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int co = co0 + mw + mi * 16 + gq + ((c & 2) ? 8 : 0);
        const int oy = ty0 + 2 * wn + (ni >> 1), ox = tx0 + (ni & 1) * 8 + 2 * tq + (c & 1);
        if (co >= g.Cout || oy >= g.H || ox >= g.W) continue;
        float* o = y + (size_t)n * g.ys_n + (size_t)co * g.ys_c + (size_t)oy * g.W + ox;
        float v = acc[mi][ni][c] + (bias ? bias[co] : 0.f);
        if (accumulate) v += *o;
        *o = relu ? fmaxf(v, 0.f) : v;
      }
}

// Weight gradient of a 3 x 3 'same' convolution (the 2-D trunk) on tiles: CTA = 32 output channels x 32 input channels x
// all 9 taps, looping over its share of the 128-pixel tiles (8 x 16 pixels of one slice, or 8 x 8 of two slices for the
// 8-pixel-wide bottom level).  Per tile the dy tile [32 co][128 px] and the input patch with its halo [32 ci][10 x 18]
// are staged once as TF32 and serve all 9 taps: dW[co][ci][tap] += sum_px dy[co][px] * x[ci][px shifted by the tap]
// = 9 x mma.sync.m16n8k8 per warp and K step (A = dy, rows = co; B = the patch shifted by the tap, columns = ci), with
// the 9 x 16 x 8 accumulators in registers for the whole tile range and one atomicAdd per (co, ci, tap) at the end.
// The im2col form (wgrad_mma_kernel) re-gathers every input element once per tap column with its bounds arithmetic.
constexpr int WT_PLANE = 228;                      // >= 200 (two 10 x 10 patches), = 4 mod 32 words: conflict-free B fragments
constexpr int WT_LDY = 132;                        // 128 + 4: conflict-free A fragments
template <bool TW8>
__global__ void __launch_bounds__(256, 2) wgrad3x3_tile_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dw, const Geom g, const int ci_blocks,
                                                               const long long ntiles, const long long tiles_per_cta) {
  __shared__ __align__(16) uint32_t s_x[32][WT_PLANE];
  __shared__ __align__(16) uint32_t s_dy[32][WT_LDY];
  constexpr int PW = TW8 ? 10 : 18, PSZ = TW8 ? 200 : 180;
  const int co0 = (blockIdx.x / ci_blocks) * 32, ci0 = (blockIdx.x % ci_blocks) * 32;
  const long long t_begin = (long long)blockIdx.y * tiles_per_cta, t_end = min(ntiles, t_begin + tiles_per_cta);
  const int tiles_x = TW8 ? 1 : ceil_div(g.W, 16), tiles_y = ceil_div(g.H, 8);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gq = lane >> 2, tq = lane & 3;
  const int mi = warp & 1, nj = warp >> 1;             // warp: co rows [16 mi, 16 mi + 16), ci columns [8 nj, 8 nj + 8)
  float acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[t][c] = 0.f;
  for (long long T = t_begin; T < t_end; ++T) {
    const int tx = (int)(T % tiles_x), ty = (int)((T / tiles_x) % tiles_y);
    const int nn = (int)(T / ((long long)tiles_x * tiles_y)) * (TW8 ? 2 : 1);      // first slice of the tile
    const int y0 = ty * 8, x0 = tx * 16;
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * 128; e += 256) {
      const int co = e >> 7, pos = e & 127;
      const int zz = TW8 ? pos >> 6 : 0, row = TW8 ? (pos >> 3) & 7 : pos >> 4, col = TW8 ? pos & 7 : pos & 15;
      const int oy = y0 + row, ox = x0 + col, n = nn + zz;
      float v = 0.f;
      if (co0 + co < g.Cout && n < g.N && oy < g.H && ox < g.W)
        v = __ldg(dy + (size_t)n * g.ys_n + (size_t)(co0 + co) * g.ys_c + (size_t)oy * g.W + ox);
      s_dy[co][pos] = to_tf32(v);
    }
    for (int e = threadIdx.x; e < 32 * PSZ; e += 256) {
      const int ci = e / PSZ, r = e - ci * PSZ;
      const int zz = TW8 ? r / 100 : 0, r2 = TW8 ? r - zz * 100 : r, py = r2 / PW, px = r2 - py * PW;
      const int iy = y0 + py - 1, ix = x0 + px - 1, n = nn + zz;
      float v = 0.f;
      if (ci0 + ci < g.Cin && n < g.N && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W)
        v = __ldg(x + (size_t)n * g.xs_n + (size_t)(ci0 + ci) * g.xs_c + (size_t)iy * g.W + ix);
      s_x[ci][r] = to_tf32(v);
    }
    __syncthreads();
#pragma unroll 2
    for (int ks = 0; ks < 16; ++ks) {
      // the 8 pixels of this K step lie in one tile row: patch index of the first one at tap (0, 0)
      const int pb = TW8 ? (ks >> 3) * 100 + (ks & 7) * 10 : (ks >> 1) * 18 + (ks & 1) * 8;
      uint32_t af[4];
      af[0] = s_dy[mi * 16 + gq][ks * 8 + tq];     af[1] = s_dy[mi * 16 + gq + 8][ks * 8 + tq];
      af[2] = s_dy[mi * 16 + gq][ks * 8 + tq + 4]; af[3] = s_dy[mi * 16 + gq + 8][ks * 8 + tq + 4];
      const uint32_t* xr = &s_x[nj * 8 + gq][pb + tq];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int off = (t / 3) * PW + (t % 3);
        uint32_t bf[2] = {xr[off], xr[off + 4]};
        mma_tf32(acc[t], af, bf);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int co = co0 + mi * 16 + gq + ((c & 2) ? 8 : 0), ci = ci0 + nj * 8 + 2 * tq + (c & 1);
      const float v = acc[t][c];
      if (co < g.Cout && ci < g.Cin && v != 0.f) atomicAdd(dw + ((size_t)co * g.Cin + ci) * 9 + t, v);
    }
}

int launch_wgrad3x3_tile(const float* x, const float* dy, float* dw, const Geom& g, cudaStream_t s) {
  const bool tw8 = g.W <= 8;
  const int co_blocks = ceil_div(g.Cout, 32), ci_blocks = ceil_div(g.Cin, 32);
  const long long ntiles = tw8 ? (long long)ceil_div(g.N, 2) * ceil_div(g.H, 8)
                               : (long long)g.N * ceil_div(g.H, 8) * ceil_div(g.W, 16);
  // enough CTAs to fill the machine a few times over, at least 4 tiles each (the 288 atomics per thread are paid once per CTA)
  long long splits = std::max<long long>(1, std::min<long long>(ceil_div<long long>(ntiles, 4),
                                                                 ceil_div<long long>(6LL * num_sms(), co_blocks * ci_blocks)));
  splits = std::min<long long>(splits, 65535);
  const long long per = ceil_div<long long>(ntiles, splits);
  splits = ceil_div<long long>(ntiles, per);
  const dim3 grid(co_blocks * ci_blocks, (unsigned)splits);
  if (tw8) wgrad3x3_tile_kernel<true><<<grid, 256, 0, s>>>(x, dy, dw, g, ci_blocks, ntiles, per);
  else wgrad3x3_tile_kernel<false><<<grid, 256, 0, s>>>(x, dy, dw, g, ci_blocks, ntiles, per);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

template <int BM, int BN>
int launch_wgrad_mma(const float* x, const float* dy, float* dw, const Geom& g, cudaStream_t s) {
  const long long total = (long long)g.N * g.Ho * g.Wo;
  const int ntot = g.kz * g.ky * g.kx * g.Cin;
  const int tiles = ceil_div(ntot, BN) * ceil_div(g.Cout, BM);
  long long splits = std::max<long long>(1, std::min<long long>(ceil_div<long long>(total, 2048), ceil_div<long long>(6LL * num_sms(), tiles)));
  splits = std::min<long long>(splits, 65535);
  long long chunk = ceil_div<long long>(total, splits);
  chunk = ceil_div<long long>(chunk, GK) * GK;
  splits = ceil_div<long long>(total, chunk);
  wgrad_mma_kernel<BM, BN><<<dim3(ceil_div(ntot, BN), ceil_div(g.Cout, BM), (unsigned)splits), G_THREADS, 0, s>>>(x, dy, dw, g, chunk);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}


bool geom_ok(const Geom* g) {
  return g && g->N > 0 && g->Cin > 0 && g->Cout > 0 && g->H > 0 && g->W > 0 && g->Ho > 0 && g->Wo > 0 && g->kz > 0 && g->ky > 0 &&
         g->kx > 0 && g->dz > 0 && g->dy > 0 && g->dx > 0 && g->stride > 0 && g->zdepth > 0 && g->N % g->zdepth == 0 &&
         g->pz >= 0 && g->py >= 0 && g->px >= 0;
}

template <int KZ, int KY, int KX, int CO_T, int CI_T>
int launch_wgrad(const float* x, const float* dy, float* dw, const Geom& g, cudaStream_t s) {
  const long long total = (long long)g.N * g.Ho * g.Wo;
  const int groups = ceil_div(g.Cout, CO_T) * ceil_div(g.Cin, CI_T);
  // enough CTAs to fill the machine a few times over; chunks of at least 16384 (n, pixel) positions so that the CTA's
  // reduction of its CO_T * CI_T * TAPS partial sums stays a small share of its work
  long long splits = std::max<long long>(1, std::min<long long>(ceil_div<long long>(total, 16384), ceil_div<long long>(8LL * num_sms(), groups)));
  splits = std::min<long long>(splits, 65535);
  const long long chunk = ceil_div<long long>(total, splits);
  splits = ceil_div<long long>(total, chunk);
  wgrad_f32_kernel<KZ, KY, KX, CO_T, CI_T><<<dim3(groups, (unsigned)splits), WG_THREADS, 0, s>>>(x, dy, dw, g, chunk);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

}  // namespace
}  // namespace cetpick

using namespace cetpick;

extern "C" int cetpick_train_conv_f32(const float* x, const float* w, const float* bias, float* y, const cetpick_conv_geom* g,
                                      int flags, void* stream) {
  g_launches = 0;
  if (!x || !w || !y || !geom_ok(g)) return CETPICK_ERR_BAD_ARG;
  const int taps = g->kz * g->ky * g->kx;
  const size_t smem = (size_t)CV_CI * taps * CV_CO * sizeof(float);
  if (smem > 48 * 1024) return CETPICK_ERR_UNSUPPORTED;
  if (g->N > 65535) return CETPICK_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (g->ky == 3 && g->kx == 3 && g->stride == 1 && g->dz == 1 && g->dy == g->dx && g->Ho == g->H && g->Wo == g->W) {
    // the hot shapes ('same' 3x3 in-plane taps): two pixels per thread, taps unrolled
    const dim3 grid2(ceil_div(g->Ho * ((g->Wo + 1) / 2), CV_THREADS), g->N, ceil_div(g->Cout, CV_CO));
    if (g->kz == 1 && g->dy == 1 && g_train_tf32 && g->Cout >= 32 && g->Cin >= 8 && g->py == 1 && g->px == 1) {
      const dim3 gm(ceil_div(g->W, CM_TW) * ceil_div(g->H, CM_TH), g->N, ceil_div(g->Cout, CM_CO));
      conv3x3_mma_kernel<<<gm, 256, 0, s>>>(x, w, bias, y, *g, flags);
      CETPICK_LAUNCH_CHECK();
      return CETPICK_OK;
    }
    if (g->kz == 1 && g->dy == 1) {
      conv3x3_f32_kernel<1, 1><<<grid2, CV_THREADS, 0, s>>>(x, w, bias, y, *g, flags);
      CETPICK_LAUNCH_CHECK();
      return CETPICK_OK;
    }
    if (g->kz == 3 && g->dy == 4) {
      conv3x3_f32_kernel<3, 4><<<grid2, CV_THREADS, 0, s>>>(x, w, bias, y, *g, flags);
      CETPICK_LAUNCH_CHECK();
      return CETPICK_OK;
    }
  }
  const dim3 grid(ceil_div(g->Ho * g->Wo, CV_THREADS), g->N, ceil_div(g->Cout, CV_CO));
  conv_f32_kernel<<<grid, CV_THREADS, smem, s>>>(x, w, bias, y, *g, flags);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

extern "C" int cetpick_train_flip_weights_f32(const float* w, float* wt, int Cout, int Cin, int taps, void* stream) {
  g_launches = 0;
  if (!w || !wt || Cout <= 0 || Cin <= 0 || taps <= 0) return CETPICK_ERR_BAD_ARG;
  const size_t total = (size_t)Cout * Cin * taps;
  flip_weights_kernel<<<(int)std::min<size_t>(ceil_div<size_t>(total, 256), 1024), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, wt, Cout, Cin, taps);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

extern "C" int cetpick_train_conv_wgrad_f32(const float* x, const float* dy, float* dw, const cetpick_conv_geom* g, void* stream) {
  g_launches = 0;
  if (!x || !dy || !dw || !geom_ok(g)) return CETPICK_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int kz = g->kz, ky = g->ky, kx = g->kx;
  // wide layers: tiled SGEMM with implicit im2col; the few-channel ends of the network (stem, hm) keep the direct kernel
  const bool mma_ok = g_train_tf32 && (long long)g->Cin * g->xs_c < 0x7fffffffLL && g->pz < 64 && g->py < 64 && g->px < 64 &&
                      (kz - 1) * g->dz < 64 && (ky - 1) * g->dy < 64 && (kx - 1) * g->dx < 64;
  static const bool tile_off = getenv("CETPICK_WGRAD_NO_TILE") != nullptr;            // A/B switch for the profiles
  if (g_train_tf32 && !tile_off && kz == 1 && ky == 3 && kx == 3 && g->dy == 1 && g->dx == 1 && g->py == 1 && g->px == 1 &&
      g->stride == 1 && g->Ho == g->H && g->Wo == g->W && g->Cout >= 32 && g->Cin >= 16)
    return launch_wgrad3x3_tile(x, dy, dw, *g, s);
  if (g->Cout >= 64 && kz * ky * kx * g->Cin >= 64)
    return mma_ok ? launch_wgrad_mma<64, 64>(x, dy, dw, *g, s) : launch_wgrad_gemm<64, 64>(x, dy, dw, *g, s);
  if (g->Cout >= 32 && kz * ky * kx * g->Cin >= 128)
    return mma_ok ? launch_wgrad_mma<32, 128>(x, dy, dw, *g, s) : launch_wgrad_gemm<32, 128>(x, dy, dw, *g, s);
  if (kz == 1 && ky == 3 && kx == 3) return launch_wgrad<1, 3, 3, 4, 2>(x, dy, dw, *g, s);
  if (kz == 3 && ky == 3 && kx == 3) return launch_wgrad<3, 3, 3, 2, 1>(x, dy, dw, *g, s);
  if (kz == 1 && ky == 7 && kx == 7) return launch_wgrad<1, 7, 7, 1, 1>(x, dy, dw, *g, s);
  if (kz == 1 && ky == 1 && kx == 1) return launch_wgrad<1, 1, 1, 4, 8>(x, dy, dw, *g, s);
  if (kz == 3 && ky == 1 && kx == 1) return launch_wgrad<3, 1, 1, 1, 8>(x, dy, dw, *g, s);
  if (kz == 1 && ky == 2 && kx == 2) return launch_wgrad<1, 2, 2, 4, 4>(x, dy, dw, *g, s);
  return CETPICK_ERR_UNSUPPORTED;
}

extern "C" int cetpick_train_set_tf32(int on) {
  g_train_tf32 = on ? 1 : 0;
  return CETPICK_OK;
}

extern "C" int cetpick_train_upconv_f32(const float* x, const float* w, const float* bias, float* y, const cetpick_conv_geom* g,
                                        void* stream) {
  g_launches = 0;
  if (!x || !w || !y || !geom_ok(g)) return CETPICK_ERR_BAD_ARG;
  if (g->Ho > 2 * g->H || g->Wo > 2 * g->W || g->N > 65535) return CETPICK_ERR_BAD_ARG;
  const dim3 grid(ceil_div(g->Ho * g->Wo, CV_THREADS), g->N, ceil_div(g->Cout, CV_CO));
  upconv_f32_kernel<<<grid, CV_THREADS, CV_CI * 4 * CV_CO * sizeof(float), static_cast<cudaStream_t>(stream)>>>(x, w, bias, y, *g);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

extern "C" int cetpick_train_net_workspace_bytes(int C_max, size_t* bytes) {
  if (!bytes || C_max <= 0) return CETPICK_ERR_BAD_ARG;
  *bytes = align_up((size_t)C_max * RED_SPLITS * 2 * sizeof(double), 256) + align_up((size_t)C_max * 2 * sizeof(float), 256);
  return CETPICK_OK;
}

namespace {
struct RedWs { double* part; float* sums; };
int red_ws(void* ws, size_t ws_bytes, int C, RedWs* out) {
  const size_t a = align_up((size_t)C * RED_SPLITS * 2 * sizeof(double), 256), b = align_up((size_t)C * 2 * sizeof(float), 256);
  if (!ws || ws_bytes < a + b || (reinterpret_cast<uintptr_t>(ws) & 255)) return CETPICK_ERR_WORKSPACE;
  out->part = static_cast<double*>(ws);
  out->sums = reinterpret_cast<float*>(static_cast<char*>(ws) + a);
  return CETPICK_OK;
}
dim3 ew_grid(int HW, int C, int N) { return dim3(std::max(1, std::min(ceil_div(HW, 256), 64)), C, N); }
}  // namespace

extern "C" int cetpick_train_bn_f32(const float* x, long long xs_n, long long xs_c, float* y, long long ys_n, long long ys_c,
                                    const float* gamma, const float* beta, float* running_mean, float* running_var,
                                    float* save_mean, float* save_invstd, int N, int C, int HW, float eps, float momentum,
                                    int relu, void* ws, size_t ws_bytes, void* stream) {
  g_launches = 0;
  if (!x || !y || !gamma || !beta || !save_mean || !save_invstd || N <= 0 || C <= 0 || HW <= 0 || N > 65535 || C > 65535)
    return CETPICK_ERR_BAD_ARG;
  RedWs r;
  if (int rc = red_ws(ws, ws_bytes, C, &r)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  chan_reduce_kernel<0><<<dim3(C, RED_SPLITS), RED_THREADS, 0, s>>>(x, xs_n, xs_c, nullptr, nullptr, 0, 0, nullptr, nullptr, N, HW, 0,
                                                                    r.part);
  CETPICK_LAUNCH_CHECK();
  bn_stats_final_kernel<<<ceil_div(C, 128), 128, 0, s>>>(r.part, C, (double)N * HW, eps, momentum, save_mean, save_invstd,
                                                        running_mean, running_var);
  CETPICK_LAUNCH_CHECK();
  bn_apply_kernel<<<ew_grid(HW, C, N), 256, 0, s>>>(x, xs_n, xs_c, y, ys_n, ys_c, gamma, beta, save_mean, save_invstd, C, HW, relu);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

extern "C" int cetpick_train_bn_bwd_f32(const float* x, long long xs_n, long long xs_c, const float* y, const float* dy,
                                        long long ys_n, long long ys_c, float* dx, long long dxs_n, long long dxs_c,
                                        const float* gamma, const float* save_mean, const float* save_invstd, float* dgamma,
                                        float* dbeta, int N, int C, int HW, int relu, void* ws, size_t ws_bytes, void* stream) {
  g_launches = 0;
  if (!x || !y || !dy || !dx || !gamma || !save_mean || !save_invstd || N <= 0 || C <= 0 || HW <= 0 || N > 65535 || C > 65535)
    return CETPICK_ERR_BAD_ARG;
  RedWs r;
  if (int rc = red_ws(ws, ws_bytes, C, &r)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  chan_reduce_kernel<1><<<dim3(C, RED_SPLITS), RED_THREADS, 0, s>>>(x, xs_n, xs_c, y, dy, ys_n, ys_c, save_mean, save_invstd, N, HW,
                                                                    relu, r.part);
  CETPICK_LAUNCH_CHECK();
  bn_bwd_final_kernel<<<ceil_div(C, 128), 128, 0, s>>>(r.part, C, r.sums, dgamma, dbeta);
  CETPICK_LAUNCH_CHECK();
  bn_bwd_apply_kernel<<<ew_grid(HW, C, N), 256, 0, s>>>(x, xs_n, xs_c, y, dy, ys_n, ys_c, dx, dxs_n, dxs_c, gamma, save_mean,
                                                        save_invstd, r.sums, (float)(1.0 / ((double)N * HW)), HW, relu);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

extern "C" int cetpick_train_channel_sum_f32(const float* x, long long xs_n, long long xs_c, float* out, int N, int C, int HW,
                                             int accumulate, void* ws, size_t ws_bytes, void* stream) {
  g_launches = 0;
  if (!x || !out || N <= 0 || C <= 0 || HW <= 0 || C > 65535) return CETPICK_ERR_BAD_ARG;
  RedWs r;
  if (int rc = red_ws(ws, ws_bytes, C, &r)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  chan_reduce_kernel<2><<<dim3(C, RED_SPLITS), RED_THREADS, 0, s>>>(x, xs_n, xs_c, nullptr, nullptr, 0, 0, nullptr, nullptr, N, HW, 0,
                                                                    r.part);
  CETPICK_LAUNCH_CHECK();
  chan_sum_final_kernel<<<ceil_div(C, 128), 128, 0, s>>>(r.part, C, out, accumulate);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

extern "C" int cetpick_train_pool_f32(const float* x, long long xs_n, long long xs_c, float* y, long long ys_n, long long ys_c,
                                      int N, int C, int H, int W, void* stream) {
  g_launches = 0;
  if (!x || !y || N <= 0 || C <= 0 || H <= 0 || W <= 0 || N > 65535 || C > 65535) return CETPICK_ERR_BAD_ARG;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  pool_f32_kernel<<<ew_grid(Ho * Wo, C, N), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, xs_n, xs_c, y, ys_n, ys_c, H, W, Ho, Wo);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

extern "C" int cetpick_train_pool_bwd_f32(const float* x, long long xs_n, long long xs_c, const float* dy, long long dys_n,
                                          long long dys_c, float* dx, long long dxs_n, long long dxs_c, int N, int C, int H, int W,
                                          int accumulate, void* stream) {
  g_launches = 0;
  if (!x || !dy || !dx || N <= 0 || C <= 0 || H <= 0 || W <= 0 || N > 65535 || C > 65535) return CETPICK_ERR_BAD_ARG;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  pool_bwd_f32_kernel<<<ew_grid(H * W, C, N), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, xs_n, xs_c, dy, dys_n, dys_c, dx, dxs_n,
                                                                                          dxs_c, H, W, Ho, Wo, accumulate);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

extern "C" int cetpick_train_relu_bwd_f32(const float* y, const float* dy, float* dx, size_t n, void* stream) {
  g_launches = 0;
  if (!y || !dy || !dx || n == 0) return CETPICK_ERR_BAD_ARG;
  relu_bwd_kernel<<<(int)std::min<size_t>(ceil_div<size_t>(n, 256), (size_t)num_sms() * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      y, dy, dx, n);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}
