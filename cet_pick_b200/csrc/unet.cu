// Detector forward plan: TomoConvUNet (cet_pick/models/networks/unet_small.py:30-97) with its 2-D
// U-Net trunk (models/networks/unet.py:198-249 DownConv, :319-399 UpConv, :861-886 UNet.forward).
//
// finalize(): folds every eval-mode BatchNorm into the preceding convolution
//   (w' = w * g / sqrt(var + 1e-5), b' = beta - mean * g / sqrt(var + 1e-5) [+ conv bias * scale]),
//   rounds the folded weights to bf16 and packs them [k-block][Cout][KC] for conv_tc.cu.
// forward(): activations are bf16 NHWC with the z axis as the batch axis (the reference folds z
//   into the batch the same way, unet_small.py:66-71), so NDHWC for the 3-D head is the same memory.
//   Layer by layer over the whole volume; torch.cat is never materialised (two TMA sources).
#include "common.cuh"
#include "conv_tc.cuh"
#include "conv_march.cuh"
#include "conv_stem.cuh"
#include "conv_up.cuh"
#include "conv_halo.cuh"
#include "conv_block.cuh"

#include <cuda_bf16.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace cetpick {

namespace {

constexpr double BN_EPS = 1e-5;

// ------------------------------------------------------------------------------------------
// CUDA-core kernels for the layers that are not GEMM-shaped
// ------------------------------------------------------------------------------------------

// Stem: Conv2d(1,16,7,stride 2,pad 3,bias=False) + BN + ReLU (unet_small.py:35-37,72-74).
// fp32 (D,H,W) in, bf16 NHWC16 (D,h,w,16) out.  One input channel: not GEMM-shaped enough for a
// TMA-fed UMMA (K = 49), so this is an FFMA kernel sized to be issue-bound on FMAs: CTA = 64 x 16
// output pixels, thread = 2 x 2 pixels x 16 channels (64 accumulators; one weight row of 16 floats is
// reused by 4 pixels).  The (133 x 37) input patch is split into even / odd columns in shared memory
// so that the stride-2 window reads of a warp hit consecutive banks.
constexpr int ST_TW = 64, ST_TH = 16, ST_IW = 2 * ST_TW + 5, ST_IH = 2 * ST_TH + 5, ST_HP = 68;

__device__ __forceinline__ float round_tf32_dev(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <bool F32OUT>
__global__ void __launch_bounds__(256, 2) stem_kernel(const float* __restrict__ in, int D, int H, int W,
                                                      int h, int w, const float* __restrict__ wgt /*[49][16]*/,
                                                      const float* __restrict__ bias /*[16]*/,
                                                      void* __restrict__ out_v) {
  __nv_bfloat16* out = static_cast<__nv_bfloat16*>(out_v);
  __shared__ float s_par[2][ST_IH][ST_HP];       // [column parity][patch row][patch column / 2]
  __shared__ __align__(16) float s_w[49 * 16];
  __shared__ float s_b[16];
  const int tid = threadIdx.x;
  for (int i = tid; i < 49 * 16; i += 256) s_w[i] = wgt[i];
  if (tid < 16) s_b[tid] = bias[tid];
  const int tiles_x = ceil_div(w, ST_TW), tiles_y = ceil_div(h, ST_TH);
  const long long total = (long long)tiles_x * tiles_y * D;
  const int lx = tid & 31, ty = tid >> 5;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int tx0 = (int)(t % tiles_x) * ST_TW;
    const int ty0 = (int)((t / tiles_x) % tiles_y) * ST_TH;
    const int z = (int)(t / ((long long)tiles_x * tiles_y));
    const float* plane = in + (size_t)z * H * W;
    const int ix0 = 2 * tx0 - 3, iy0 = 2 * ty0 - 3;
    __syncthreads();
    for (int i = tid; i < ST_IH * ST_IW; i += 256) {
      const int r = i / ST_IW, c = i - r * ST_IW;
      const int gy = iy0 + r, gx = ix0 + c;
      s_par[c & 1][r][c >> 1] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(plane + (size_t)gy * W + gx) : 0.f;
    }
    __syncthreads();
    float acc[2][2][16];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[a][b][c] = s_b[c];
#pragma unroll
    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        float wv[16];
        const float4* wr = reinterpret_cast<const float4*>(&s_w[(ky * 7 + kx) * 16]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 f = wr[q];
          wv[4 * q] = f.x; wv[4 * q + 1] = f.y; wv[4 * q + 2] = f.z; wv[4 * q + 3] = f.w;
        }
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            // output pixel (yy, xx) = (2*ty + a, lx + 32*b): patch row 2*yy + ky, patch column 2*xx + kx
            const float v = s_par[kx & 1][2 * (2 * ty + a) + ky][lx + 32 * b + (kx >> 1)];
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[a][b][c] = fmaf(v, wv[c], acc[a][b][c]);
          }
      }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ox = tx0 + lx + 32 * b, oy = ty0 + 2 * ty + a;
        if (F32OUT) {                       // TF32 mode: fp32 NHWC16, values rounded to TF32
          if (ox < w && oy < h) {
            float4* d4 = reinterpret_cast<float4*>(static_cast<float*>(out_v) + (((size_t)z * h + oy) * w + ox) * 16);
#pragma unroll
            for (int c = 0; c < 4; ++c)
              d4[c] = make_float4(round_tf32_dev(fmaxf(acc[a][b][4 * c], 0.f)), round_tf32_dev(fmaxf(acc[a][b][4 * c + 1], 0.f)),
                                  round_tf32_dev(fmaxf(acc[a][b][4 * c + 2], 0.f)), round_tf32_dev(fmaxf(acc[a][b][4 * c + 3], 0.f)));
          }
        } else if (ox < w && oy < h) {
          uint32_t pk[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            __nv_bfloat162 v2 = __floats2bfloat162_rn(fmaxf(acc[a][b][2 * c], 0.f), fmaxf(acc[a][b][2 * c + 1], 0.f));
            pk[c] = *reinterpret_cast<uint32_t*>(&v2);
          }
          uint4* dst = reinterpret_cast<uint4*>(out + (((size_t)z * h + oy) * w + ox) * 16);
          dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
  }
}

// MaxPool2d(2, ceil_mode=True) (unet.py:225) on bf16 NHWC; one thread = 8 channels of an out pixel.
__global__ void __launch_bounds__(256) pool2x2_kernel(const __nv_bfloat16* __restrict__ in, int N, int H,
                                                      int W, int C, __nv_bfloat16* __restrict__ out) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, C8 = C / 8;
  const size_t total = (size_t)N * Ho * Wo * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % C8);
    size_t r = i / C8;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const int y0 = 2 * oy, x0 = 2 * ox;
    const uint4* base = reinterpret_cast<const uint4*>(in);
    auto at = [&](int y, int x) { return base[(((size_t)n * H + y) * W + x) * C8 + c8]; };
    uint4 m = at(y0, x0);
    auto mx = [](uint4 a, uint4 b) {
      uint4 r;
      __nv_bfloat162* ra = reinterpret_cast<__nv_bfloat162*>(&a);
      __nv_bfloat162* rb = reinterpret_cast<__nv_bfloat162*>(&b);
      __nv_bfloat162* rr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
      for (int k = 0; k < 4; ++k) rr[k] = __hmax2(ra[k], rb[k]);
      return r;
    };
    const bool hx = x0 + 1 < W, hy = y0 + 1 < H;
    if (hx) m = mx(m, at(y0, x0 + 1));
    if (hy) m = mx(m, at(y0 + 1, x0));
    if (hx && hy) m = mx(m, at(y0 + 1, x0 + 1));
    reinterpret_cast<uint4*>(out)[i] = m;
  }
}

// TF32 mode twins (fp32 NHWC activations): MaxPool2d(2, ceil) with one thread per 4 channels, and the hm head
__global__ void __launch_bounds__(256) pool2x2_f32_kernel(const float* __restrict__ in, int N, int H, int W, int C,
                                                          float* __restrict__ out) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2, C4 = C / 4;
  const size_t total = (size_t)N * Ho * Wo * C4;
  const float4* base = reinterpret_cast<const float4*>(in);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % C4);
    size_t r = i / C4;
    const int ox = (int)(r % Wo); r /= Wo;
    const int oy = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const int y0 = 2 * oy, x0 = 2 * ox;
    auto at = [&](int y, int x) { return base[(((size_t)n * H + y) * W + x) * C4 + c4]; };
    auto mx = [](float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); };
    float4 m = at(y0, x0);
    const bool hx = x0 + 1 < W, hy = y0 + 1 < H;
    if (hx) m = mx(m, at(y0, x0 + 1));
    if (hy) m = mx(m, at(y0 + 1, x0));
    if (hx && hy) m = mx(m, at(y0 + 1, x0 + 1));
    reinterpret_cast<float4*>(out)[i] = m;
  }
}

__global__ void __launch_bounds__(256) hm_head_f32_kernel(const float* __restrict__ feat, int D, size_t plane,
                                                          const float* __restrict__ wgt /*[3][32]*/, int apply_sigmoid,
                                                          float* __restrict__ out) {
  __shared__ float s_w[96];
  if (threadIdx.x < 96) s_w[threadIdx.x] = wgt[threadIdx.x];
  __syncthreads();
  const size_t total = (size_t)D * plane;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int z = (int)(i / plane);
    float acc = 0.f;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
      const int zz = z + dz;
      if (zz < 0 || zz >= D) continue;
      const float4* src = reinterpret_cast<const float4*>(feat + (i + (ptrdiff_t)dz * (ptrdiff_t)plane) * 32);
      const float* wr = &s_w[(dz + 1) * 32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = src[q];
        acc = fmaf(v.x, wr[4 * q], acc); acc = fmaf(v.y, wr[4 * q + 1], acc);
        acc = fmaf(v.z, wr[4 * q + 2], acc); acc = fmaf(v.w, wr[4 * q + 3], acc);
      }
    }
    if (apply_sigmoid) {
      const float y = 1.0f / (1.0f + expf(-acc));
      acc = fminf(fmaxf(y, 1e-4f), 1.0f - 1e-4f);
    }
    out[i] = acc;
  }
}

// `hm` head: Conv3d(C,1,(3,1,1),pad (1,0,0),bias=False) (unet_small.py:53-61,89) + optional
// _sigmoid (models/utils.py:167-169).  bf16 (D,h,w,C) in, fp32 (D,h,w) out; C == 32.
__global__ void __launch_bounds__(256) hm_head_kernel(const __nv_bfloat16* __restrict__ feat, int D,
                                                      size_t plane, const float* __restrict__ wgt /*[3][32]*/,
                                                      int apply_sigmoid, float* __restrict__ out) {
  __shared__ float s_w[96];
  if (threadIdx.x < 96) s_w[threadIdx.x] = wgt[threadIdx.x];
  __syncthreads();
  const size_t total = (size_t)D * plane;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int z = (int)(i / plane);
    float acc = 0.f;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
      const int zz = z + dz;
      if (zz < 0 || zz >= D) continue;
      const uint4* src = reinterpret_cast<const uint4*>(feat + (i + (ptrdiff_t)dz * (ptrdiff_t)plane) * 32);
      const float* wr = &s_w[(dz + 1) * 32];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 v = src[q];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(h2[k]);
          acc = fmaf(f.x, wr[q * 8 + 2 * k], acc);
          acc = fmaf(f.y, wr[q * 8 + 2 * k + 1], acc);
        }
      }
    }
    if (apply_sigmoid) {
      const float y = 1.0f / (1.0f + expf(-acc));
      acc = fminf(fmaxf(y, 1e-4f), 1.0f - 1e-4f);
    }
    out[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------
struct PackedConv {
  int KC = 0, ntaps = 0, Ntot = 0, nsrc = 1, C[2] = {0, 0};
  int tap[27][3] = {};
  int relu = 0;
  size_t w_off = 0, b_off = 0;   // byte offsets into the device weight blob
  bool has_bias = false;
  double flops_per_pixel = 0;    // 2 * K * N
  bool upk = false;              // ConvTranspose weights packed for conv_up.cu
  bool halo = false;             // 3x3 weights packed for conv_halo.cu
  int march = -1;                // >= 0: MarchMode of conv_march.cu (w_off then holds ITS weight image)
};

// Optional per-launch timing (CUDA events on the launching stream) for bench.py's roofline.
struct Profiler {
  bool on = false;
  std::vector<cudaEvent_t> ev;
  std::vector<std::string> names;
  std::vector<double> flops;
  int n = 0;
  void begin() { n = 0; names.clear(); flops.clear(); }
  void mark(const char* name, double fl, cudaStream_t st) {
    if (!on) return;
    if ((int)ev.size() <= n) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
    cudaEventRecord(ev[n], st);
    names.push_back(name); flops.push_back(fl);
    ++n;
  }
};

}  // namespace
}  // namespace cetpick

using namespace cetpick;

struct cetpick_unet {
  int n_blocks, head_conv, proj_c;
  int precision = 0;             // 0: BF16 operands (the specialised kernels); 1: TF32 operands (generic kernel, fp32 maps)
  mutable Profiler prof;         // per plan: two plans (or two streams of two plans) never share timing state
  std::map<std::string, std::vector<float>> params;
  bool finalized = false;
  std::vector<uint8_t> blob;     // host staging of all packed weights
  void* d_blob = nullptr;
  // packed layers
  size_t stem_w = 0, stem_b = 0, stem_tc_w = 0, hm_w = 0;
  float stem_shift[16] = {};
  bool fold_cf = false;          // conv_final folded into feature_head.0 (weights + tap-validity bias table)
  size_t fh0_btab = 0;
  std::vector<PackedConv> down1, down2, upc, up1, up2;
  PackedConv conv_final, fh0, fh2, proj;

  const std::vector<float>* get(const std::string& k, size_t numel) const {
    auto it = params.find(k);
    if (it == params.end() || it->second.size() != numel) return nullptr;
    return &it->second;
  }
};

namespace {

struct Fold { std::vector<double> scale, shift; };

// eval-mode BatchNorm as y = x * scale + shift
bool bn_fold(const cetpick_unet* m, const std::string& p, int C, Fold& f) {
  auto w = m->get(p + ".weight", C), b = m->get(p + ".bias", C);
  auto mu = m->get(p + ".running_mean", C), var = m->get(p + ".running_var", C);
  if (!w || !b || !mu || !var) return false;
  f.scale.resize(C); f.shift.resize(C);
  for (int c = 0; c < C; ++c) {
    const double s = (double)(*w)[c] / std::sqrt((double)(*var)[c] + BN_EPS);
    f.scale[c] = s;
    f.shift[c] = (double)(*b)[c] - (double)(*mu)[c] * s;
  }
  return true;
}

// fp32 -> nearest TF32 (10-bit mantissa, ties away from zero like cvt.rna.tf32.f32)
float tf32_round_host(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return x;
  u = (u + 0x1000u) & 0xffffe000u;
  memcpy(&x, &u, 4);
  return x;
}

size_t blob_alloc(cetpick_unet* m, size_t bytes) {
  const size_t off = align_up(m->blob.size(), 256);
  m->blob.resize(off + bytes, 0);
  return off;
}

// Pack a PyTorch conv weight (Cout, Cin_total, *kernel) into [k-block][Cout][KC] bf16 with the
// per-output-channel scale folded in.  The K loop is (source, tap, channel chunk).
bool pack_conv(cetpick_unet* m, const std::string& wkey, int Cout, int nsrc, int Csrc, int ntaps,
               const Fold* fold, const std::vector<float>* conv_bias, int relu, PackedConv& pc,
               const std::vector<float>* w_override = nullptr) {
  const int Cin = nsrc * Csrc;
  auto w = w_override ? w_override : m->get(wkey, (size_t)Cout * Cin * ntaps);
  if (!w || w->size() != (size_t)Cout * Cin * ntaps) return false;
  const bool tf32 = m->precision == 1;
  pc.KC = std::min(tf32 ? 32 : 64, Csrc);
  if (Csrc % pc.KC) return false;
  pc.ntaps = ntaps; pc.Ntot = Cout; pc.nsrc = nsrc; pc.C[0] = Csrc; pc.C[1] = nsrc > 1 ? Csrc : 0;
  pc.relu = relu;
  const int chunks = Csrc / pc.KC;
  const size_t nkb = (size_t)nsrc * ntaps * chunks;
  const int mmode = ntaps == 9 ? MARCH_2D_ROWS : ntaps == 27 ? MARCH_3D_PLANES : -1;
  if (tf32) {
    // TF32 mode: every convolution runs through the generic implicit GEMM with fp32 weights rounded to TF32
    pc.w_off = blob_alloc(m, nkb * Cout * pc.KC * 4);
    float* dst = reinterpret_cast<float*>(m->blob.data() + pc.w_off);
    for (int s = 0; s < nsrc; ++s)
      for (int t = 0; t < ntaps; ++t)
        for (int ch = 0; ch < chunks; ++ch) {
          const size_t kb = ((size_t)s * ntaps + t) * chunks + ch;
          for (int n = 0; n < Cout; ++n)
            for (int k = 0; k < pc.KC; ++k) {
              const int ci = s * Csrc + ch * pc.KC + k;
              const double v = (double)(*w)[((size_t)n * Cin + ci) * ntaps + t] * (fold ? fold->scale[n] : 1.0);
              dst[(kb * Cout + n) * pc.KC + k] = tf32_round_host((float)v);
            }
        }
  } else if (mmode >= 0 && march_supported(mmode, Csrc, nsrc, Cout)) {
    // narrow layer: weight image of the marching kernel (conv_march.cu)
    const std::vector<uint16_t> pk = march_pack_weights(mmode, w->data(), Cout, nsrc, Csrc, fold ? fold->scale.data() : nullptr);
    pc.march = mmode;
    pc.w_off = blob_alloc(m, pk.size() * 2);
    memcpy(m->blob.data() + pc.w_off, pk.data(), pk.size() * 2);
  } else if (ntaps == 9 && halo_supported(Csrc, nsrc, Cout) && (fold || conv_bias)) {
    // wide level: weight image of the halo-tile kernel (conv_halo.cu)
    const std::vector<uint16_t> pk = halo_pack_weights(w->data(), Cout, nsrc, Csrc, fold ? fold->scale.data() : nullptr);
    pc.halo = true;
    pc.w_off = blob_alloc(m, pk.size() * 2);
    memcpy(m->blob.data() + pc.w_off, pk.data(), pk.size() * 2);
  } else {
  pc.w_off = blob_alloc(m, nkb * Cout * pc.KC * 2);
  uint16_t* dst = reinterpret_cast<uint16_t*>(m->blob.data() + pc.w_off);
  for (int s = 0; s < nsrc; ++s)
    for (int t = 0; t < ntaps; ++t)
      for (int ch = 0; ch < chunks; ++ch) {
        const size_t kb = ((size_t)s * ntaps + t) * chunks + ch;
        for (int n = 0; n < Cout; ++n)
          for (int k = 0; k < pc.KC; ++k) {
            const int ci = s * Csrc + ch * pc.KC + k;
            const double v = (double)(*w)[((size_t)n * Cin + ci) * ntaps + t] * (fold ? fold->scale[n] : 1.0);
            dst[(kb * Cout + n) * pc.KC + k] = f2bf_host((float)v);
          }
      }
  }
  pc.has_bias = fold || conv_bias;
  if (pc.has_bias) {
    pc.b_off = blob_alloc(m, (size_t)Cout * 4);
    float* b = reinterpret_cast<float*>(m->blob.data() + pc.b_off);
    for (int n = 0; n < Cout; ++n) {
      double v = conv_bias ? (double)(*conv_bias)[n] : 0.0;
      if (fold) v = v * fold->scale[n] + fold->shift[n];
      b[n] = (float)v;
    }
  }
  pc.flops_per_pixel = 2.0 * Cin * ntaps * Cout;
  return true;
}

void taps_3x3(PackedConv& pc) {
  int t = 0;
  for (int ky = 0; ky < 3; ++ky)
    for (int kx = 0; kx < 3; ++kx) { pc.tap[t][0] = 0; pc.tap[t][1] = ky - 1; pc.tap[t][2] = kx - 1; ++t; }
}
void taps_3x3x3_dil(PackedConv& pc, int dy, int dx) {
  int t = 0;
  for (int kz = 0; kz < 3; ++kz)
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        pc.tap[t][0] = kz - 1; pc.tap[t][1] = (ky - 1) * dy; pc.tap[t][2] = (kx - 1) * dx; ++t;
      }
}
void taps_3x1x1(PackedConv& pc) {
  for (int kz = 0; kz < 3; ++kz) { pc.tap[kz][0] = kz - 1; pc.tap[kz][1] = 0; pc.tap[kz][2] = 0; }
}

struct Dims { int h, w; };

std::vector<Dims> level_dims(int n_blocks, int64_t H, int64_t W) {
  std::vector<Dims> d(n_blocks);
  d[0] = {(int)((H - 1) / 2 + 1), (int)((W - 1) / 2 + 1)};   // 7x7 stride-2 pad-3 stem
  for (int i = 1; i < n_blocks; ++i) d[i] = {(d[i - 1].h + 1) / 2, (d[i - 1].w + 1) / 2};   // ceil-mode pool
  return d;
}

struct WsPlan {
  std::vector<size_t> off;     // 3 buffers per level
  std::vector<size_t> size;
  size_t total = 0;
};

WsPlan ws_plan(int n_blocks, int64_t D, int64_t H, int64_t W, int esz = 2) {
  WsPlan p;
  auto dims = level_dims(n_blocks, H, W);
  size_t o = 0;
  for (int i = 0; i < n_blocks; ++i) {
    const size_t C = (size_t)32 << i;
    const size_t sz = align_up((size_t)D * dims[i].h * dims[i].w * C * esz, 1024);
    p.size.push_back(sz);
    for (int b = 0; b < 3; ++b) { p.off.push_back(o); o += sz; }
  }
  p.total = o;
  return p;
}

struct HeadExtras {
  const float* bias_tab = nullptr;
  const float* bias_tab_host = nullptr;
  const float* hm_w = nullptr;
  const float* hm_w_host = nullptr;
  float* hm_out = nullptr;
  int hm_sigmoid = 0;
  int z_origin = 0;
};

int run_conv(const cetpick_unet* m, const std::string& name, const PackedConv& pc, const void* s0, const void* s1,
             int NIMG, int H, int W, int epi, void* out, int Ho, int Wo, int Cout, cudaStream_t st,
             const HeadExtras* ex = nullptr, void* pool_out = nullptr) {
  m->prof.mark((std::string("conv:") + name + (pc.march >= 0 ? ":march" : pc.halo ? ":halo" : ":tc")).c_str(),
              pc.flops_per_pixel * (double)NIMG * H * W, st);
  const float* bias = pc.has_bias ? reinterpret_cast<const float*>(static_cast<const uint8_t*>(m->d_blob) + pc.b_off) : nullptr;
  if (pc.march >= 0) {
    if (epi != EPI_BF16_NHWC) return CETPICK_ERR_STATE;
    MarchLaunch M;
    M.mode = pc.march; M.dil = 4; M.nsrc = pc.nsrc; M.src[0] = s0; M.src[1] = s1; M.C = pc.C[0];
    M.NIMG = NIMG; M.H = H; M.W = W;
    M.wpk = static_cast<const uint8_t*>(m->d_blob) + pc.w_off;
    M.Cout = pc.Ntot; M.bias = bias; M.relu = pc.relu; M.out = out; M.pool_out = pool_out;
    M.bias_host = pc.has_bias ? reinterpret_cast<const float*>(m->blob.data() + pc.b_off) : nullptr;
    if (ex) {
      M.bias_tab = ex->bias_tab; M.bias_tab_host = ex->bias_tab_host;
      M.hm_w = ex->hm_w; M.hm_w_host = ex->hm_w_host; M.hm_out = ex->hm_out; M.hm_sigmoid = ex->hm_sigmoid;
      M.z_origin = ex->z_origin;
    }
    return conv_march_launch(M, st);
  }
  if (pc.halo) {
    if (epi != EPI_BF16_NHWC || !bias) return CETPICK_ERR_STATE;
    HaloLaunch Hh;
    Hh.nsrc = pc.nsrc; Hh.src[0] = s0; Hh.src[1] = s1; Hh.C = pc.C[0]; Hh.NIMG = NIMG; Hh.H = H; Hh.W = W;
    Hh.wpk = static_cast<const uint8_t*>(m->d_blob) + pc.w_off; Hh.bias = bias; Hh.Cout = pc.Ntot;
    Hh.relu = pc.relu; Hh.out = out;
    return conv_halo_launch(Hh, st);
  }
  ConvLaunch L;
  L.nsrc = pc.nsrc; L.src[0] = s0; L.src[1] = s1; L.C[0] = pc.C[0]; L.C[1] = pc.C[1];
  L.NIMG = NIMG; L.H = H; L.W = W;
  L.wpk = static_cast<const uint8_t*>(m->d_blob) + pc.w_off;
  L.KC = pc.KC; L.ntaps = pc.ntaps; L.tf32 = m->precision == 1;
  memcpy(L.tap, pc.tap, sizeof(L.tap));
  L.Ntot = pc.Ntot;
  L.bias = bias;
  L.relu = pc.relu; L.epi = epi; L.out = out; L.out_cstride = pc.Ntot;
  L.Ho = Ho; L.Wo = Wo; L.Cout = Cout;
  return conv_tc_launch(L, st);
}

// conv1 + conv2 (+ pool) of a 32-channel full-resolution block as ONE kernel (conv_block.cu) when both layers were
// packed for the marching kernel and a thread-block cluster can span the row.  CETPICK_BLOCK=1 selects it (A/B measurements).
bool use_block(const PackedConv& c1, const PackedConv& c2, int W) {
  // Off unless CETPICK_BLOCK=1: the fused kernel is exact but, with one 128-pixel M-tile per CTA and a single group of
  // conv1 epilogue warps on the critical path of every row, it is slower than the two marching kernels it replaces
  // (profiles/r3o_block_*: tensor pipe 10 % active, latency-bound); see DESIGN.md.
  const char* e = getenv("CETPICK_BLOCK");
  const bool on = e && e[0] == '1';
  return on && c1.march == MARCH_2D_ROWS && c2.march == MARCH_2D_ROWS && c1.Ntot == 32 && c2.Ntot == 32 && c2.nsrc == 1 &&
         c2.C[0] == 32 && c1.relu && c2.relu && c1.has_bias && c2.has_bias && W > 128 && block_supported(c1.C[0], c1.nsrc, W);
}

int run_block(const cetpick_unet* m, const std::string& name, const PackedConv& c1, const PackedConv& c2, const void* s0,
              const void* s1, int NIMG, int H, int W, void* out, void* pool_out, cudaStream_t st) {
  m->prof.mark((std::string("conv:") + name + ".c1+c2:block").c_str(), (c1.flops_per_pixel + c2.flops_per_pixel) * (double)NIMG * H * W, st);
  BlockLaunch B;
  B.nsrc = c1.nsrc; B.src[0] = s0; B.src[1] = s1; B.C1 = c1.C[0]; B.NIMG = NIMG; B.H = H; B.W = W;
  B.w1pk = static_cast<const uint8_t*>(m->d_blob) + c1.w_off;
  B.w2pk = static_cast<const uint8_t*>(m->d_blob) + c2.w_off;
  B.bias1_host = reinterpret_cast<const float*>(m->blob.data() + c1.b_off);
  B.bias2_host = reinterpret_cast<const float*>(m->blob.data() + c2.b_off);
  B.out = out; B.pool_out = pool_out;
  return conv_block_launch(B, st);
}

}  // namespace

extern "C" int cetpick_unet_create(cetpick_unet** plan, int n_blocks, int head_conv, int proj_channels) {
  if (!plan || n_blocks < 2 || n_blocks > 6) return CETPICK_ERR_BAD_ARG;
  if (head_conv != 32) return CETPICK_ERR_UNSUPPORTED;       // opts.py:207-211 default for task semi
  if (proj_channels < 0 || proj_channels > 256 || (proj_channels % 16)) return CETPICK_ERR_UNSUPPORTED;
  cetpick_unet* m = new cetpick_unet();
  m->n_blocks = n_blocks; m->head_conv = head_conv; m->proj_c = proj_channels;
  *plan = m;
  return CETPICK_OK;
}

extern "C" void cetpick_unet_destroy(cetpick_unet* m) {
  if (!m) return;
  for (cudaEvent_t e : m->prof.ev) cudaEventDestroy(e);
  if (m->d_blob) cudaFree(m->d_blob);
  delete m;
}

extern "C" int cetpick_unet_set_param(cetpick_unet* m, const char* key, const float* data, int64_t numel) {
  if (!m || !key || (!data && numel > 0) || numel < 0) return CETPICK_ERR_BAD_ARG;
  const std::string k(key);
  if (k.size() >= 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return CETPICK_OK;
  m->params[k].assign(data, data + numel);
  m->finalized = false;
  return CETPICK_OK;
}

extern "C" int cetpick_unet_set_precision(cetpick_unet* m, int mode) {
  if (!m || (mode != 0 && mode != 1)) return CETPICK_ERR_BAD_ARG;
  if (m->precision != mode) m->finalized = false;
  m->precision = mode;
  return CETPICK_OK;
}

extern "C" int cetpick_unet_finalize(cetpick_unet* m) {
  if (!m) return CETPICK_ERR_BAD_ARG;
  m->blob.clear();
  m->down1.clear(); m->down2.clear(); m->upc.clear(); m->up1.clear(); m->up2.clear();
  const int nb = m->n_blocks;
  Fold f;
  // stem
  {
    auto w = m->get("conv1.weight", 16 * 49);
    if (!w || !bn_fold(m, "bn1", 16, f)) return CETPICK_ERR_STATE;
    m->stem_w = blob_alloc(m, 49 * 16 * 4);
    m->stem_b = blob_alloc(m, 16 * 4);
    float* sw = reinterpret_cast<float*>(m->blob.data() + m->stem_w);
    float* sb = reinterpret_cast<float*>(m->blob.data() + m->stem_b);
    for (int c = 0; c < 16; ++c) {
      for (int t = 0; t < 49; ++t) sw[t * 16 + c] = (float)((double)(*w)[c * 49 + t] * f.scale[c]);
      sb[c] = (float)f.shift[c];
      m->stem_shift[c] = (float)f.shift[c];
    }
    // tensor-core image of the same weights (conv_stem.cu)
    const std::vector<uint16_t> pk = stem_pack_weights(w->data(), f.scale.data());
    m->stem_tc_w = blob_alloc(m, pk.size() * 2);
    memcpy(m->blob.data() + m->stem_tc_w, pk.data(), pk.size() * 2);
  }
  int outs = 16;
  for (int i = 0; i < nb; ++i) {
    const int ins = (i == 0) ? 16 : outs;
    outs = 32 << i;
    const std::string p = "unet.down_convs." + std::to_string(i);
    PackedConv c1, c2;
    if (!bn_fold(m, p + ".norm0", outs, f) || !pack_conv(m, p + ".conv1.weight", outs, 1, ins, 9, &f, nullptr, 1, c1)) return CETPICK_ERR_STATE;
    if (!bn_fold(m, p + ".norm1", outs, f) || !pack_conv(m, p + ".conv2.weight", outs, 1, outs, 9, &f, nullptr, 1, c2)) return CETPICK_ERR_STATE;
    taps_3x3(c1); taps_3x3(c2);
    m->down1.push_back(c1); m->down2.push_back(c2);
  }
  for (int i = 0; i < nb - 1; ++i) {
    const int ins = outs;
    outs = ins / 2;
    const std::string p = "unet.up_convs." + std::to_string(i);
    // ConvTranspose2d(ins, outs, 2, 2) + bias, then norm0 + ReLU: a GEMM with N = 4*outs columns
    PackedConv u;
    {
      auto w = m->get(p + ".upconv.weight", (size_t)ins * outs * 4);
      auto b = m->get(p + ".upconv.bias", outs);
      if (!w || !b || !bn_fold(m, p + ".norm0", outs, f)) return CETPICK_ERR_STATE;
      const bool tf32 = m->precision == 1;
      u.KC = std::min(tf32 ? 32 : 64, ins); u.ntaps = 1; u.Ntot = 4 * outs; u.nsrc = 1; u.C[0] = ins; u.relu = 1;
      if (tf32) {
        const int chunks = ins / u.KC;
        u.w_off = blob_alloc(m, (size_t)chunks * u.Ntot * u.KC * 4);
        float* dst = reinterpret_cast<float*>(m->blob.data() + u.w_off);
        for (int ch = 0; ch < chunks; ++ch)
          for (int q = 0; q < 4; ++q)
            for (int co = 0; co < outs; ++co)
              for (int k = 0; k < u.KC; ++k) {
                const int ci = ch * u.KC + k;
                const double v = (double)(*w)[((size_t)ci * outs + co) * 4 + q] * f.scale[co];
                dst[((size_t)ch * u.Ntot + q * outs + co) * u.KC + k] = tf32_round_host((float)v);
              }
      } else if (upconv_supported(ins, outs)) {
        const std::vector<uint16_t> pk = upconv_pack_weights(w->data(), ins, outs, f.scale.data());
        u.upk = true;
        u.w_off = blob_alloc(m, pk.size() * 2);
        memcpy(m->blob.data() + u.w_off, pk.data(), pk.size() * 2);
      } else {
        const int chunks = ins / u.KC;
        u.w_off = blob_alloc(m, (size_t)chunks * u.Ntot * u.KC * 2);
        uint16_t* dst = reinterpret_cast<uint16_t*>(m->blob.data() + u.w_off);
        for (int ch = 0; ch < chunks; ++ch)
          for (int q = 0; q < 4; ++q)
            for (int co = 0; co < outs; ++co)
              for (int k = 0; k < u.KC; ++k) {
                const int ci = ch * u.KC + k;
                const double v = (double)(*w)[((size_t)ci * outs + co) * 4 + q] * f.scale[co];
                dst[((size_t)ch * u.Ntot + q * outs + co) * u.KC + k] = f2bf_host((float)v);
              }
      }
      u.has_bias = true;
      u.b_off = blob_alloc(m, (size_t)u.Ntot * 4);
      float* bb = reinterpret_cast<float*>(m->blob.data() + u.b_off);
      for (int q = 0; q < 4; ++q)
        for (int co = 0; co < outs; ++co) bb[q * outs + co] = (float)((double)(*b)[co] * f.scale[co] + f.shift[co]);
      u.flops_per_pixel = 2.0 * ins * 4 * outs;
    }
    PackedConv c1, c2;
    if (!bn_fold(m, p + ".norm1", outs, f) || !pack_conv(m, p + ".conv1.weight", outs, 2, outs, 9, &f, nullptr, 1, c1)) return CETPICK_ERR_STATE;
    if (!bn_fold(m, p + ".norm2", outs, f) || !pack_conv(m, p + ".conv2.weight", outs, 1, outs, 9, &f, nullptr, 1, c2)) return CETPICK_ERR_STATE;
    taps_3x3(c1); taps_3x3(c2);
    m->upc.push_back(u); m->up1.push_back(c1); m->up2.push_back(c2);
  }
  if (outs != 32) return CETPICK_ERR_STATE;
  {
    auto b = m->get("unet.conv_final.bias", 32);
    if (!b || !pack_conv(m, "unet.conv_final.weight", 32, 1, 32, 1, nullptr, b, 0, m->conv_final)) return CETPICK_ERR_STATE;
    if (!pack_conv(m, "feature_head.0.weight", 32, 1, 32, 27, nullptr, nullptr, 1, m->fh0)) return CETPICK_ERR_STATE;
    m->fold_cf = false;
    if (m->fh0.march >= 0) {
      // conv_final is a bias-carrying 1x1 conv with no activation between it and feature_head.0
      // (unet.py:882 -> unet_small.py:85): fh0(Wc u + bc) = (W_t Wc) * u + sum over in-volume taps of
      // W_t bc.  The second term depends only on which taps are inside the volume: a 64-row table.
      auto wf = m->get("feature_head.0.weight", (size_t)32 * 32 * 27);
      auto wc = m->get("unet.conv_final.weight", (size_t)32 * 32);
      std::vector<float> wfold((size_t)32 * 32 * 27);
      for (int co = 0; co < 32; ++co)
        for (int ci = 0; ci < 32; ++ci)
          for (int t = 0; t < 27; ++t) {
            double a = 0.0;
            for (int mid = 0; mid < 32; ++mid) a += (double)(*wf)[((size_t)co * 32 + mid) * 27 + t] * (double)(*wc)[mid * 32 + ci];
            wfold[((size_t)co * 32 + ci) * 27 + t] = (float)a;
          }
      PackedConv folded;
      if (!pack_conv(m, "", 32, 1, 32, 27, nullptr, nullptr, 1, folded, &wfold)) return CETPICK_ERR_STATE;
      m->fh0 = folded;
      m->fh0_btab = blob_alloc(m, (size_t)64 * 32 * 4);
      float* tab = reinterpret_cast<float*>(m->blob.data() + m->fh0_btab);
      for (int cz = 0; cz < 4; ++cz)
        for (int cy = 0; cy < 4; ++cy)
          for (int cx = 0; cx < 4; ++cx)
            for (int co = 0; co < 32; ++co) {
              double a = 0.0;
              for (int kz = 0; kz < 3; ++kz)
                for (int ky = 0; ky < 3; ++ky)
                  for (int kx = 0; kx < 3; ++kx) {
                    auto ok = [](int k, int c) { return k == 1 || (k == 0 && (c & 1)) || (k == 2 && (c & 2)); };
                    if (!ok(kz, cz) || !ok(ky, cy) || !ok(kx, cx)) continue;
                    for (int mid = 0; mid < 32; ++mid)
                      a += (double)(*wf)[((size_t)co * 32 + mid) * 27 + (kz * 3 + ky) * 3 + kx] * (double)(*b)[mid];
                  }
              tab[((cz * 4 + cy) * 4 + cx) * 32 + co] = (float)a;
            }
      m->fold_cf = true;
    }
    if (!pack_conv(m, "feature_head.2.weight", 32, 1, 32, 27, nullptr, nullptr, 1, m->fh2)) return CETPICK_ERR_STATE;
    taps_3x3x3_dil(m->fh0, 4, 4); taps_3x3x3_dil(m->fh2, 4, 4);
    auto hw = m->get("hm.weight", 32 * 3);
    if (!hw) return CETPICK_ERR_STATE;
    m->hm_w = blob_alloc(m, 96 * 4);
    float* d = reinterpret_cast<float*>(m->blob.data() + m->hm_w);
    for (int c = 0; c < 32; ++c)
      for (int kz = 0; kz < 3; ++kz) d[kz * 32 + c] = (*hw)[c * 3 + kz];
    if (m->proj_c > 0) {
      if (!pack_conv(m, "proj.weight", m->proj_c, 1, 32, 3, nullptr, nullptr, 0, m->proj)) return CETPICK_ERR_STATE;
      taps_3x1x1(m->proj);
    }
  }
  if (m->d_blob) { cudaFree(m->d_blob); m->d_blob = nullptr; }
  CETPICK_CUDA(cudaMalloc(&m->d_blob, m->blob.size()));
  CETPICK_CUDA(cudaMemcpy(m->d_blob, m->blob.data(), m->blob.size(), cudaMemcpyHostToDevice));
  m->finalized = true;
  return CETPICK_OK;
}

extern "C" int cetpick_unet_workspace_bytes(const cetpick_unet* m, int64_t D, int64_t H, int64_t W,
                                            int want_proj, size_t* bytes) {
  (void)want_proj;
  if (!m || !bytes || D <= 0 || H <= 0 || W <= 0) return CETPICK_ERR_BAD_ARG;
  auto dims = level_dims(m->n_blocks, H, W);
  if (dims.back().h < 1 || dims.back().w < 1) return CETPICK_ERR_BAD_ARG;
  *bytes = ws_plan(m->n_blocks, D, H, W, m->precision == 1 ? 4 : 2).total + 1024;
  return CETPICK_OK;
}

namespace {
int unet_forward_impl(cetpick_unet* m, const float* tomo, const uint8_t* tomo_u8, const float* lut_host, int64_t D64,
                      int64_t H64, int64_t W64, int64_t z_origin, float* hm, int apply_sigmoid, float* proj, void* ws,
                      size_t ws_bytes, void* stream) {
  g_launches = 0;
  if (!m || (!tomo && !tomo_u8) || !hm || D64 <= 0 || H64 <= 0 || W64 <= 0) return CETPICK_ERR_BAD_ARG;
  if (tomo_u8 && (!lut_host || lut_host[0] != 0.f)) return CETPICK_ERR_BAD_ARG;
  if (!m->finalized) return CETPICK_ERR_STATE;
  if (proj && m->proj_c == 0) return CETPICK_ERR_STATE;
  if (D64 > 32767 || H64 > (1 << 20) || W64 > (1 << 20)) return CETPICK_ERR_BAD_ARG;
  const int D = (int)D64, H = (int)H64, W = (int)W64, nb = m->n_blocks;
  const bool tf32 = m->precision == 1;
  const WsPlan wp = ws_plan(nb, D, H, W, tf32 ? 4 : 2);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  if (!ws || ws_bytes < wp.total + (size_t)(base - static_cast<uint8_t*>(ws))) return CETPICK_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto dims = level_dims(nb, H, W);
  auto buf = [&](int level, int b) { return reinterpret_cast<__nv_bfloat16*>(base + wp.off[level * 3 + b]); };
  const uint8_t* blob = static_cast<const uint8_t*>(m->d_blob);
  const int sms = num_sms();
  int rc;

  // stem -> X0 (16 channels)
  {
    m->prof.begin();
    m->prof.mark("stem", 2.0 * 49 * 16 * (double)D * dims[0].h * dims[0].w, st);
    if (tf32) {                                 // TF32 mode: CUDA-core fp32 stem (exact fp32 FMAs), fp32 NHWC16 out
      if (tomo_u8) return CETPICK_ERR_UNSUPPORTED;
      const long long tiles = (long long)ceil_div(dims[0].w, ST_TW) * ceil_div(dims[0].h, ST_TH) * D;
      const int grid = (int)std::min<long long>(tiles, (long long)sms * 2);
      stem_kernel<true><<<grid, 256, 0, st>>>(tomo, D, H, W, dims[0].h, dims[0].w, reinterpret_cast<const float*>(blob + m->stem_w),
                                              reinterpret_cast<const float*>(blob + m->stem_b), buf(0, 0));
      CETPICK_LAUNCH_CHECK();
    } else if (tomo_u8) {                       // quantised levels straight into the tensor-core march
      if (!stem_tc_supported_u8(tomo_u8, W)) return CETPICK_ERR_UNSUPPORTED;   // rows must be 16-byte aligned
      StemLaunch SL;
      SL.in_u8 = tomo_u8; SL.D = D; SL.H = H; SL.W = W; SL.wpk = blob + m->stem_tc_w; SL.out = buf(0, 0);
      for (int k = 0; k < 256; ++k) SL.lut[k] = f2bf_host(lut_host[k]);
      memcpy(SL.bias, m->stem_shift, sizeof(SL.bias));
      if ((rc = conv_stem_launch(SL, st))) return rc;
    } else if (stem_tc_supported(tomo, W)) {   // tensor-core march (conv_stem.cu)
      StemLaunch SL;
      SL.in = tomo; SL.D = D; SL.H = H; SL.W = W; SL.wpk = blob + m->stem_tc_w; SL.out = buf(0, 0);
      memcpy(SL.bias, m->stem_shift, sizeof(SL.bias));
      if ((rc = conv_stem_launch(SL, st))) return rc;
    } else {                                    // rows not 16-byte aligned: CUDA-core kernel
      const long long tiles = (long long)ceil_div(dims[0].w, ST_TW) * ceil_div(dims[0].h, ST_TH) * D;
      const int grid = (int)std::min<long long>(tiles, (long long)sms * 2);
      stem_kernel<false><<<grid, 256, 0, st>>>(tomo, D, H, W, dims[0].h, dims[0].w,
                                        reinterpret_cast<const float*>(blob + m->stem_w),
                                        reinterpret_cast<const float*>(blob + m->stem_b), buf(0, 0));
      CETPICK_LAUNCH_CHECK();
    }
  }
  // encoder: level i: in Y0 -> conv1 -> Y1 -> conv2 -> Y2 (skip) -> pool -> next level's Y0
  for (int i = 0; i < nb; ++i) {
    const int h = dims[i].h, w = dims[i].w;
    if (use_block(m->down1[i], m->down2[i], w)) {
      if ((rc = run_block(m, "down" + std::to_string(i), m->down1[i], m->down2[i], buf(i, 0), nullptr, D, h, w, buf(i, 2),
                          i < nb - 1 ? buf(i + 1, 0) : nullptr, st))) return rc;
      continue;
    }
    if ((rc = run_conv(m, "down" + std::to_string(i) + ".c1", m->down1[i], buf(i, 0), nullptr, D, h, w, EPI_BF16_NHWC, buf(i, 1), 0, 0, 0, st))) return rc;
    // MaxPool2d(2, ceil) (unet.py:225) comes out of the marching kernel's epilogue where that kernel runs
    const bool fuse_pool = (i < nb - 1) && m->down2[i].march == MARCH_2D_ROWS && m->down2[i].relu;
    if ((rc = run_conv(m, "down" + std::to_string(i) + ".c2", m->down2[i], buf(i, 1), nullptr, D, h, w, EPI_BF16_NHWC, buf(i, 2), 0, 0, 0, st,
                       nullptr, fuse_pool ? buf(i + 1, 0) : nullptr))) return rc;
    if (i < nb - 1 && !fuse_pool) {
      const int C = 32 << i;
      m->prof.mark("pool2x2", 0.0, st);
      const size_t total = (size_t)D * dims[i + 1].h * dims[i + 1].w * (C / (tf32 ? 4 : 8));
      const int grid = (int)std::min<size_t>(ceil_div<size_t>(total, 256), (size_t)sms * 16);
      if (tf32) pool2x2_f32_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(buf(i, 2)), D, h, w, C, reinterpret_cast<float*>(buf(i + 1, 0)));
      else pool2x2_kernel<<<grid, 256, 0, st>>>(buf(i, 2), D, h, w, C, buf(i + 1, 0));
      CETPICK_LAUNCH_CHECK();
    }
  }
  // decoder: up i works at level j = nb-2-i; input = output of the level below
  const __nv_bfloat16* below = buf(nb - 1, 2);
  for (int i = 0; i < nb - 1; ++i) {
    const int j = nb - 2 - i;
    const int h = dims[j].h, w = dims[j].w, Cout = 32 << j;
    if (m->upc[i].upk) {
      const PackedConv& u = m->upc[i];
      m->prof.mark(("conv:up" + std::to_string(i) + ".upconv:up").c_str(), u.flops_per_pixel * (double)D * dims[j + 1].h * dims[j + 1].w, st);
      UpLaunch U;
      U.src = below; U.Cin = u.C[0]; U.NIMG = D; U.h = dims[j + 1].h; U.w = dims[j + 1].w;
      U.wpk = blob + u.w_off; U.bias = reinterpret_cast<const float*>(blob + u.b_off);
      U.Cout = Cout; U.out = buf(j, 0); U.Ho = h; U.Wo = w;
      if ((rc = conv_up_launch(U, st))) return rc;
    } else if ((rc = run_conv(m, "up" + std::to_string(i) + ".upconv", m->upc[i], below, nullptr, D, dims[j + 1].h, dims[j + 1].w, EPI_UPCONV_2X2, buf(j, 0), h, w, Cout, st))) return rc;
    if (use_block(m->up1[i], m->up2[i], w)) {
      if ((rc = run_block(m, "up" + std::to_string(i), m->up1[i], m->up2[i], buf(j, 0), buf(j, 2), D, h, w, buf(j, 1), nullptr, st))) return rc;
      below = buf(j, 1);
      continue;
    }
    if ((rc = run_conv(m, "up" + std::to_string(i) + ".c1", m->up1[i], buf(j, 0), buf(j, 2), D, h, w, EPI_BF16_NHWC, buf(j, 1), 0, 0, 0, st))) return rc;
    if ((rc = run_conv(m, "up" + std::to_string(i) + ".c2", m->up2[i], buf(j, 1), nullptr, D, h, w, EPI_BF16_NHWC, buf(j, 0), 0, 0, 0, st))) return rc;
    below = buf(j, 0);
  }
  const int h0 = dims[0].h, w0 = dims[0].w;
  // conv_final (1x1 + bias) is folded into feature_head.0; feature_head: X0 -> X1 -> (X0 | hm).
  // Without a `proj` request the hm head (+ _sigmoid) is fused into feature_head.2's epilogue and the
  // 32-channel feature map is never written.
  {
    const float* hmw = reinterpret_cast<const float*>(blob + m->hm_w);
    // trunk output X (buffer 0 or 1 of level 0, depending on the fused block) and the other buffer Y
    __nv_bfloat16* X = const_cast<__nv_bfloat16*>(below);
    __nv_bfloat16* Y = (X == buf(0, 0)) ? buf(0, 1) : buf(0, 0);
    const __nv_bfloat16* f_in = X;
    __nv_bfloat16* f_out = Y;
    if (!m->fold_cf) {
      if ((rc = run_conv(m, "conv_final", m->conv_final, X, nullptr, D, h0, w0, EPI_BF16_NHWC, Y, 0, 0, 0, st))) return rc;
      HeadExtras ex0;
      ex0.z_origin = (int)(z_origin % 840);
      if ((rc = run_conv(m, "fhead0", m->fh0, Y, nullptr, D, h0, w0, EPI_BF16_NHWC, X, 0, 0, 0, st, &ex0))) return rc;
      f_in = X; f_out = Y;
    } else {
      HeadExtras ex;
      ex.z_origin = (int)(z_origin % 840);     // 840 = lcm(1..8): every ring length divides it
      ex.bias_tab = reinterpret_cast<const float*>(blob + m->fh0_btab);
      ex.bias_tab_host = reinterpret_cast<const float*>(m->blob.data() + m->fh0_btab);
      if ((rc = run_conv(m, "fhead0+conv_final", m->fh0, X, nullptr, D, h0, w0, EPI_BF16_NHWC, Y, 0, 0, 0, st, &ex))) return rc;
      f_in = Y; f_out = X;
    }
    if (!proj && m->fh2.march >= 0) {
      HeadExtras ex;
      ex.z_origin = (int)(z_origin % 840);
      ex.hm_w = hmw; ex.hm_w_host = reinterpret_cast<const float*>(m->blob.data() + m->hm_w);
      ex.hm_out = hm; ex.hm_sigmoid = apply_sigmoid;
      if ((rc = run_conv(m, "fhead2+hm", m->fh2, f_in, nullptr, D, h0, w0, EPI_BF16_NHWC, nullptr, 0, 0, 0, st, &ex))) return rc;
    } else {
      HeadExtras ex2;
      ex2.z_origin = (int)(z_origin % 840);
      if ((rc = run_conv(m, "fhead2", m->fh2, f_in, nullptr, D, h0, w0, EPI_BF16_NHWC, f_out, 0, 0, 0, st, &ex2))) return rc;
      const size_t plane = (size_t)h0 * w0, total = plane * D;
      m->prof.mark("hm_head", 2.0 * 96 * (double)total, st);
      const int grid = (int)std::min<size_t>(ceil_div<size_t>(total, 256), (size_t)sms * 16);
      if (tf32) hm_head_f32_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(f_out), D, plane, hmw, apply_sigmoid, hm);
      else hm_head_kernel<<<grid, 256, 0, st>>>(f_out, D, plane, hmw, apply_sigmoid, hm);
      CETPICK_LAUNCH_CHECK();
      if (proj) {
        if ((rc = run_conv(m, "proj", m->proj, f_out, nullptr, D, h0, w0, EPI_F32_L2NORM_NCDHW, proj, 0, 0, 0, st))) return rc;
      }
    }
  }
  m->prof.mark("end", 0.0, st);
  return CETPICK_OK;
}
}  // namespace

extern "C" int cetpick_unet_forward(cetpick_unet* m, const float* tomo, int64_t D, int64_t H, int64_t W,
                                    float* hm, int apply_sigmoid, float* proj, void* ws, size_t ws_bytes,
                                    void* stream) {
  if (!tomo) return CETPICK_ERR_BAD_ARG;
  return unet_forward_impl(m, tomo, nullptr, nullptr, D, H, W, 0, hm, apply_sigmoid, proj, ws, ws_bytes, stream);
}

extern "C" int cetpick_unet_forward_u8(cetpick_unet* m, const uint8_t* tomo_q, const float* level_values_host,
                                       int64_t D, int64_t H, int64_t W, float* hm, int apply_sigmoid, float* proj,
                                       void* ws, size_t ws_bytes, void* stream) {
  if (!tomo_q || !level_values_host) return CETPICK_ERR_BAD_ARG;
  return unet_forward_impl(m, nullptr, tomo_q, level_values_host, D, H, W, 0, hm, apply_sigmoid, proj, ws, ws_bytes, stream);
}

extern "C" int cetpick_unet_forward_slab(cetpick_unet* m, const float* tomo, const uint8_t* tomo_q,
                                         const float* level_values_host, int64_t D, int64_t H, int64_t W,
                                         int64_t z_origin, float* hm, int apply_sigmoid, float* proj, void* ws,
                                         size_t ws_bytes, void* stream) {
  if ((tomo == nullptr) == (tomo_q == nullptr) || z_origin < 0) return CETPICK_ERR_BAD_ARG;
  if (tomo_q && !level_values_host) return CETPICK_ERR_BAD_ARG;
  return unet_forward_impl(m, tomo, tomo_q, level_values_host, D, H, W, z_origin, hm, apply_sigmoid, proj, ws, ws_bytes,
                           stream);
}

extern "C" int cetpick_unet_profile_enable(cetpick_unet* m, int on) {
  if (!m) return CETPICK_ERR_BAD_ARG;
  m->prof.on = on != 0;
  return CETPICK_OK;
}

// Per-launch device times of the most recent profiled cetpick_unet_forward of this plan (synchronises).
extern "C" int cetpick_unet_profile_read(cetpick_unet* m, int max_entries, int* n, float* ms, double* flops, char* names32) {
  if (!m || !n) return CETPICK_ERR_BAD_ARG;
  Profiler& P = m->prof;
  const int cnt = std::max(0, P.n - 1);
  *n = cnt;
  if (cnt == 0) return CETPICK_OK;
  CETPICK_CUDA(cudaEventSynchronize(P.ev[P.n - 1]));
  for (int i = 0; i < cnt && i < max_entries; ++i) {
    float t = 0.f;
    CETPICK_CUDA(cudaEventElapsedTime(&t, P.ev[i], P.ev[i + 1]));
    if (ms) ms[i] = t;
    if (flops) flops[i] = P.flops[i];
    if (names32) { strncpy(names32 + (size_t)i * 32, P.names[i].c_str(), 31); names32[(size_t)i * 32 + 31] = 0; }
  }
  return CETPICK_OK;
}

#ifdef CETPICK_TEST_HOOKS   // test / tuning hooks: built into libcetpick_test_sm100a.so only (include/cetpick_test.h)
// Test hook: one convolution through conv_tc.cu with caller-packed weights (tests/test_gpu_conv.py).
extern "C" int cetpick_conv_bf16(int nsrc, const void* src0, int C0, const void* src1, int C1, int NIMG,
                                 int H, int W, const void* wpk, int KC, int ntaps, const int* taps,
                                 int Ntot, const float* bias, int relu, int epi, void* out, int out_cstride,
                                 int Ho, int Wo, int Cout, void* stream) {
  g_launches = 0;
  if (!taps || ntaps < 1 || ntaps > 27) return CETPICK_ERR_BAD_ARG;
  ConvLaunch L;
  L.nsrc = nsrc; L.src[0] = src0; L.src[1] = src1; L.C[0] = C0; L.C[1] = C1;
  L.NIMG = NIMG; L.H = H; L.W = W; L.wpk = wpk; L.KC = KC; L.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) for (int k = 0; k < 3; ++k) L.tap[t][k] = taps[t * 3 + k];
  L.Ntot = Ntot; L.bias = bias; L.relu = relu; L.epi = epi; L.out = out; L.out_cstride = out_cstride;
  L.Ho = Ho; L.Wo = Wo; L.Cout = Cout;
  return conv_tc_launch(L, static_cast<cudaStream_t>(stream));
}

// Test hook: counters of the tensor-map cache (conv_tc.cu)
extern "C" int cetpick_tmap_cache_stats(int64_t* hits, int64_t* misses) {
  tmap_cache_stats(hits, misses);
  return CETPICK_OK;
}

#endif  // CETPICK_TEST_HOOKS
