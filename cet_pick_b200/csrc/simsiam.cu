// Exploration-step embedding network on the device: TomoResClassifier.forward_test of
// cet_pick/models/networks/simsiam_model.py:325-366 (arch `simsiam3d_18` / `simsiam_18`; BasicBlock :44-73, stages
// :256-271, 3-D feature layer + heads :181-215), the network simsiam_test_hm_3d.py:136-195 runs over every candidate
// sub-volume (BASELINE.json configs[3]: 8192 sub-volumes of 32^3).
//
// finalize(): every eval-mode BatchNorm (2d / 3d / 1d, affine or not) is folded into the preceding conv / Linear in
//   double precision, weights rounded to bf16 once and packed for conv_small.cu.
// forward(): all slices of all sub-volumes go through the 2-D trunk as one batch of small maps (the reference reshapes
//   (B,D,H,W) -> (B*D,1,H,W), :333-337), bf16 NHWC activations:
//     stem_pool_kernel  conv 7x7 s2 p3 (1 -> 64) + BN + ReLU + MaxPool2d(3, 2, 1)   (CUDA cores: one input channel)
//     conv_small_kernel every 3x3 / 1x1 conv of the BasicBlocks (stride 1 / 2, residual add + ReLU in the epilogue),
//                       the Conv3d(256,256,3) feature layer (one 2x2xD sub-volume per M-tile) and the Linear layers
//     avgpool_kernel    AdaptiveAvgPool3d((1,1,1))
//
// 2-D variant (cetpick_simsiam_create_2d): TomoResClassifier2D.forward_test of
// cet_pick/models/networks/simsiam_model_2d.py:617-774 (arch `simsiam2d_18`): conv1 is 3x3 stride 1 without a max-pool,
// the three stages run on the full / half / quarter resolution of a 2-D patch (32x32 -> 16x16 -> 8x8), then
// AdaptiveAvgPool2d, fc: 256 -> head_conv, and the proj / pred MLPs of width head_conv.
//     stem2d_kernel     conv 3x3 s1 p1 (1 -> 64) + BN + ReLU (CUDA cores)
//     conv_small_kernel the BasicBlocks in row bands of 128 pixels, the Linear layers
#include "common.cuh"
#include "conv_small.cuh"

#include <cuda_bf16.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace cetpick {
namespace {

constexpr double BN_EPS = 1e-5;

// conv1 7x7 s2 p3 (1->64) + BN + ReLU + MaxPool2d(3,2,1): one CTA per slice.  H = W = 32 (conv map 16x16, pooled 8x8)
// or 16 (8x8 -> 4x4).  Thread = one conv-map row segment of 16 (or 8) pixels x 4 channels; the ReLU'd conv map goes to
// shared memory as fp32 and is pooled from there (every window holds at least one in-map pixel and values are >= 0).
template <int HW>
__global__ void __launch_bounds__(256) stem_pool_kernel(const float* __restrict__ in, long long nslices,
                                                        const float* __restrict__ wgt /*[49][64], BN scale folded*/,
                                                        const float* __restrict__ shift /*[64]*/,
                                                        __nv_bfloat16* __restrict__ out /*[nslices][HW/4][HW/4][64]*/) {
  constexpr int CW = HW / 2, PW = HW / 4, IP = HW + 6;
  extern __shared__ __align__(16) uint8_t stem_smem[];
  float* s_w = reinterpret_cast<float*>(stem_smem);                                   // [49][64]
  float (*s_in)[IP + 1] = reinterpret_cast<float (*)[IP + 1]>(stem_smem + 49 * 64 * 4);   // [IP][IP+1]
  // ReLU'd conv map as bf16 (max-pooling commutes with the monotone rounding); 66-element rows: odd word stride
  __nv_bfloat16 (*s_map)[66] = reinterpret_cast<__nv_bfloat16 (*)[66]>(stem_smem + 49 * 64 * 4 + ((IP * (IP + 1) * 4 + 15) & ~15));
  const int tid = threadIdx.x;
  for (int i = tid; i < 49 * 64; i += 256) s_w[i] = wgt[i];
  const int cg = tid & 15, rowseg = tid >> 4;           // 16 channel groups x 16 row segments
  float sh[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) sh[c] = shift[cg * 4 + c];
  for (long long s = blockIdx.x; s < nslices; s += gridDim.x) {
    __syncthreads();
    const float* src = in + (size_t)s * HW * HW;
    for (int i = tid; i < IP * IP; i += 256) {
      const int r = i / IP, c = i - r * IP;
      const int y = r - 3, x = c - 3;
      s_in[r][c] = (y >= 0 && y < HW && x >= 0 && x < HW) ? __ldg(src + y * HW + x) : 0.f;
    }
    __syncthreads();
    // conv rows: CW rows of CW pixels; a thread takes SEG pixels of one row
    constexpr int SEG = CW * CW / 16;                   // pixels per thread: 16 (HW 32) or 4 (HW 16)
    const int p0 = rowseg * SEG, oy = p0 / CW, ox0 = p0 % CW;
    float acc[SEG][4];
#pragma unroll
    for (int i = 0; i < SEG; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[i][c] = sh[c];
#pragma unroll 1
    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const float4 w4 = *reinterpret_cast<const float4*>(&s_w[(ky * 7 + kx) * 64 + cg * 4]);
#pragma unroll
        for (int i = 0; i < SEG; ++i) {
          const float v = s_in[2 * oy + ky][2 * (ox0 + i) + kx];
          acc[i][0] = fmaf(v, w4.x, acc[i][0]); acc[i][1] = fmaf(v, w4.y, acc[i][1]);
          acc[i][2] = fmaf(v, w4.z, acc[i][2]); acc[i][3] = fmaf(v, w4.w, acc[i][3]);
        }
      }
#pragma unroll
    for (int i = 0; i < SEG; ++i)
#pragma unroll
      for (int c = 0; c < 4; ++c) s_map[p0 + i][cg * 4 + c] = __float2bfloat16_rn(fmaxf(acc[i][c], 0.f));
    __syncthreads();
    // MaxPool2d(3, stride 2, pad 1) on the CW x CW map -> PW x PW
    for (int o = tid; o < PW * PW * 32; o += 256) {
      const int c2 = o & 31, pp = o >> 5, py = pp / PW, px = pp % PW;
      float m0 = 0.f, m1 = 0.f;                         // post-ReLU values: 0 is a safe identity
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = 2 * py + dy;
        if (y < 0 || y >= CW) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = 2 * px + dx;
          if (x < 0 || x >= CW) continue;
          m0 = fmaxf(m0, __bfloat162float(s_map[y * CW + x][2 * c2]));
          m1 = fmaxf(m1, __bfloat162float(s_map[y * CW + x][2 * c2 + 1]));
        }
      }
      reinterpret_cast<__nv_bfloat162*>(out + ((size_t)s * PW * PW + pp) * 64)[c2] = __floats2bfloat162_rn(m0, m1);
    }
  }
}

// The same stem on the legacy tensor path (mma.sync m16n8k8, TF32 operands, fp32 accumulation) for H = W = 32: per slice
// an implicit GEMM  [256 conv pixels] x [49 taps, padded to 56] x [64 channels].  The A fragments are gathered straight
// from the zero-framed input patch in shared memory (element (pixel, tap) = patch[2*oy + ky][2*ox + kx]), the folded
// weights sit in shared memory as TF32 (72-word rows: conflict-free B fragments); warp w owns conv rows 2w, 2w+1 and
// all 64 channels.  Shift + ReLU + bf16 into the shared conv map (72-element rows: conflict-free C stores), pooled as
// in stem_pool_kernel.  The fp32 CUDA-core version above spent 11.7 ms of the 39 ms configs[3] batch here.
__device__ __forceinline__ uint32_t f2tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int SM_IP = 38, SM_IPP = 39, SM_KP = 56, SM_LDW = 72, SM_LDM = 72;
constexpr int SM_SMEM = SM_KP * SM_LDW * 4 + ((SM_IP * SM_IPP * 4 + 15) & ~15) + 256 * SM_LDM * 2;

__global__ void __launch_bounds__(256) stem_pool_mma_kernel(const float* __restrict__ in, long long nslices,
                                                            const float* __restrict__ wgt /*[49][64], BN scale folded*/,
                                                            const float* __restrict__ shift /*[64]*/,
                                                            __nv_bfloat16* __restrict__ out /*[nslices][8][8][64]*/) {
  constexpr int HW = 32, CW = 16, PW = 8;
  extern __shared__ __align__(16) uint8_t stem_smem[];
  uint32_t (*s_w)[SM_LDW] = reinterpret_cast<uint32_t (*)[SM_LDW]>(stem_smem);                          // [56][72] tf32
  uint32_t* s_in = reinterpret_cast<uint32_t*>(stem_smem + SM_KP * SM_LDW * 4);                        // [38][39] tf32
  __nv_bfloat16 (*s_map)[SM_LDM] =
      reinterpret_cast<__nv_bfloat16 (*)[SM_LDM]>(stem_smem + SM_KP * SM_LDW * 4 + ((SM_IP * SM_IPP * 4 + 15) & ~15));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  // rows 49 / 50 of the padded K axis carry the BatchNorm shift (split into two TF32 terms, so it stays exact to ~2^-22);
  // the A operand has 1.0 in those two columns: the accumulator comes out with the shift already added
  for (int i = tid; i < SM_KP * 64; i += 256) {
    const int k = i >> 6, c = i & 63;
    uint32_t v = 0u;
    if (k < 49) v = f2tf32(__ldg(wgt + k * 64 + c));
    else if (k == 49) v = f2tf32(__ldg(shift + c));
    else if (k == 50) { const float sv = __ldg(shift + c); v = f2tf32(sv - __uint_as_float(f2tf32(sv))); }
    s_w[k][c] = v;
  }
  // this thread's elements of the zero-framed 38 x 38 patch: (patch offset, source offset or -1 for the frame)
  constexpr int NPRE = (SM_IP * SM_IP + 255) / 256;
  float pre[NPRE];                                      // the NEXT slice's values, in flight during this slice's MMAs
  auto fetch = [&](long long s) {
    const float* src = in + (size_t)s * HW * HW;
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
      const int i = tid + 256 * j;
      const int r = i / SM_IP, c = i - r * SM_IP, y = r - 3, x = c - 3;
      pre[j] = (i < SM_IP * SM_IP && y >= 0 && y < HW && x >= 0 && x < HW) ? __ldg(src + y * HW + x) : 0.f;
    }
  };
  if ((long long)blockIdx.x < nslices) fetch(blockIdx.x);
  for (long long s = blockIdx.x; s < nslices; s += gridDim.x) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NPRE; ++j) {
      const int i = tid + 256 * j;
      const int r = i / SM_IP, c = i - r * SM_IP;
      if (i < SM_IP * SM_IP) s_in[r * SM_IPP + c] = f2tf32(pre[j]);
    }
    __syncthreads();
    if (s + gridDim.x < nslices) fetch(s + gridDim.x);
    float acc[2][8][4];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
      for (int ni = 0; ni < 8; ++ni)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[mi][ni][c] = 0.f;
#pragma unroll
    for (int k8 = 0; k8 < 7; ++k8) {
      // patch offsets of this thread's two taps of the step (tap t -> row t / 7, column t % 7); columns 49 / 50 = 1.0
      const int t0 = k8 * 8 + tq, t1 = t0 + 4;
      const int o0 = (t0 / 7) * SM_IPP + (t0 % 7), o1 = (t1 / 7) * SM_IPP + (t1 % 7);
      uint32_t a[2][4];
#pragma unroll
      for (int mi = 0; mi < 2; ++mi) {
        // conv row oy = 2 * warp + mi, pixels ox = gq and gq + 8: patch origin (2 * oy, 2 * ox)
        const int base = (2 * (2 * warp + mi)) * SM_IPP + 2 * gq;
        if (k8 < 6) {
          a[mi][0] = s_in[base + o0];      a[mi][1] = s_in[base + 16 + o0];
          a[mi][2] = s_in[base + o1];      a[mi][3] = s_in[base + 16 + o1];
        } else {
          const uint32_t one = 0x3f800000u;
          a[mi][0] = t0 == 48 ? s_in[base + o0] : (t0 <= 50 ? one : 0u);
          a[mi][1] = t0 == 48 ? s_in[base + 16 + o0] : (t0 <= 50 ? one : 0u);
          a[mi][2] = 0u; a[mi][3] = 0u;
        }
      }
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        uint32_t b[2];
        b[0] = s_w[k8 * 8 + tq][ni * 8 + gq];
        b[1] = s_w[k8 * 8 + tq + 4][ni * 8 + gq];
        mma_m16n8k8_tf32(acc[0][ni], a[0], b);
        mma_m16n8k8_tf32(acc[1][ni], a[1], b);
      }
    }
#pragma unroll
    for (int mi = 0; mi < 2; ++mi) {
      const int p = (2 * warp + mi) * CW + gq;
#pragma unroll
      for (int ni = 0; ni < 8; ++ni) {
        const int c = ni * 8 + 2 * tq;
        *reinterpret_cast<__nv_bfloat162*>(&s_map[p][c]) =
            __floats2bfloat162_rn(fmaxf(acc[mi][ni][0], 0.f), fmaxf(acc[mi][ni][1], 0.f));
        *reinterpret_cast<__nv_bfloat162*>(&s_map[p + 8][c]) =
            __floats2bfloat162_rn(fmaxf(acc[mi][ni][2], 0.f), fmaxf(acc[mi][ni][3], 0.f));
      }
    }
    __syncthreads();
    // MaxPool2d(3, stride 2, pad 1) on the 16 x 16 map -> 8 x 8
    for (int o = tid; o < PW * PW * 32; o += 256) {
      const int c2 = o & 31, pp = o >> 5, py = pp / PW, px = pp % PW;
      float m0 = 0.f, m1 = 0.f;                         // post-ReLU values: 0 is a safe identity
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = 2 * py + dy;
        if (y < 0 || y >= CW) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = 2 * px + dx;
          if (x < 0 || x >= CW) continue;
          const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&s_map[y * CW + x][2 * c2]));
          m0 = fmaxf(m0, v.x);
          m1 = fmaxf(m1, v.y);
        }
      }
      reinterpret_cast<__nv_bfloat162*>(out + ((size_t)s * PW * PW + pp) * 64)[c2] = __floats2bfloat162_rn(m0, m1);
    }
  }
}

// 2-D variant: conv1 3x3 s1 p1 (1 -> 64) + BN + ReLU (simsiam_model_2d.py:625-627, 755-757), fp32 patch in, bf16 NHWC out.
// One CTA per patch; thread = (pixel lane, 8-channel group): its 9 x 8 folded weights live in registers, the patch with
// a zero frame in shared memory; 8 neighbouring threads write the 128 B of one pixel.
__global__ void __launch_bounds__(256) stem2d_kernel(const float* __restrict__ in, long long npatch, int HW,
                                                     const float* __restrict__ wgt /*[9][64], BN scale folded*/,
                                                     const float* __restrict__ shift /*[64]*/,
                                                     __nv_bfloat16* __restrict__ out /*[npatch][HW][HW][64]*/) {
  extern __shared__ __align__(16) uint8_t stem_smem[];
  float* s_in = reinterpret_cast<float*>(stem_smem);          // [(HW+2)][(HW+2)]
  const int IP = HW + 2, tid = threadIdx.x, cg = tid & 7, lane_px = tid >> 3;
  float w[9][8], sh[8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) w[t][c] = __ldg(wgt + t * 64 + cg * 8 + c);
#pragma unroll
  for (int c = 0; c < 8; ++c) sh[c] = __ldg(shift + cg * 8 + c);
  for (long long s = blockIdx.x; s < npatch; s += gridDim.x) {
    __syncthreads();
    const float* src = in + (size_t)s * HW * HW;
    for (int i = tid; i < IP * IP; i += 256) {
      const int r = i / IP, c = i - r * IP, y = r - 1, x = c - 1;
      s_in[i] = (y >= 0 && y < HW && x >= 0 && x < HW) ? __ldg(src + y * HW + x) : 0.f;
    }
    __syncthreads();
    for (int px = lane_px; px < HW * HW; px += 32) {
      const int y = px / HW, x = px - y * HW;
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = sh[c];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float v = s_in[(y + ky) * IP + x + kx];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = fmaf(v, w[ky * 3 + kx][c], acc[c]);
        }
      uint4 o;
      __nv_bfloat162 h;
      h = __floats2bfloat162_rn(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f)); o.x = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f)); o.y = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f)); o.z = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f)); o.w = *reinterpret_cast<uint32_t*>(&h);
      *reinterpret_cast<uint4*>(out + ((size_t)s * HW * HW + px) * 64 + cg * 8) = o;
    }
  }
}

// AdaptiveAvgPool3d((1,1,1)) / AdaptiveAvgPool2d((1,1)): bf16 [B][P][C] -> bf16 [B][C] (fp32 sum); one CTA per batch element, C = 256 threads
__global__ void __launch_bounds__(256) avgpool_kernel(const __nv_bfloat16* __restrict__ in, int B, int P, int C,
                                                      __nv_bfloat16* __restrict__ out) {
  for (int b = blockIdx.x; b < B; b += gridDim.x)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float a = 0.f;
      const __nv_bfloat16* src = in + (size_t)b * P * C + c;
      for (int i = 0; i < P; ++i) a += __bfloat162float(src[(size_t)i * C]);
      out[(size_t)b * C + c] = __float2bfloat16_rn(a / (float)P);
    }
}

struct Fold { std::vector<double> scale, shift; };

}  // namespace
}  // namespace cetpick

using namespace cetpick;

struct cetpick_simsiam {
  int layers[3];
  int two_d = 0;            // 1: TomoResClassifier2D (simsiam_model_2d.py:617-774)
  int out_dim = 256;        // width of fc and the heads (head_conv in the 2-D variant)
  bool has_proj, has_pred;
  std::map<std::string, std::vector<float>> params;
  bool finalized = false;
  std::vector<uint8_t> blob;
  void* d_blob = nullptr;
  struct Conv { size_t w_off = 0, b_off = 0; int Cin = 0, Cout = 0, ntaps = 0, stride = 1; bool has_bias = false; };
  struct Block { Conv c1, c2, ds; bool has_ds = false; };
  size_t stem_w = 0, stem_b = 0;
  std::vector<Block> blocks;
  Conv f3d, fc, proj0, proj3, proj6, pred0, pred3;

  const std::vector<float>* get(const std::string& k, size_t numel) const {
    auto it = params.find(k);
    if (it == params.end() || it->second.size() != numel) return nullptr;
    return &it->second;
  }
};

namespace {

size_t balloc(cetpick_simsiam* m, size_t bytes) {
  const size_t off = align_up(m->blob.size(), 256);
  m->blob.resize(off + bytes, 0);
  return off;
}

bool bn_fold(const cetpick_simsiam* m, const std::string& p, int C, bool affine, Fold& f) {
  auto mu = m->get(p + ".running_mean", C), var = m->get(p + ".running_var", C);
  auto w = affine ? m->get(p + ".weight", C) : nullptr, b = affine ? m->get(p + ".bias", C) : nullptr;
  if (!mu || !var || (affine && (!w || !b))) return false;
  f.scale.resize(C); f.shift.resize(C);
  for (int c = 0; c < C; ++c) {
    const double s = (affine ? (double)(*w)[c] : 1.0) / std::sqrt((double)(*var)[c] + BN_EPS);
    f.scale[c] = s;
    f.shift[c] = (affine ? (double)(*b)[c] : 0.0) - (double)(*mu)[c] * s;
  }
  return true;
}

// conv / Linear weight (Cout, Cin, taps) with an optional BatchNorm behind it and an optional bias of its own
// CoutP / CinP >= Cout / Cin: the layer is built CoutP x CinP wide with zero weights, zero bias in the padding (head widths
// that are not multiples of 64 in the 2-D variant: padded activations are exactly 0 through Linear, BatchNorm and ReLU)
bool pack(cetpick_simsiam* m, const std::string& wkey, int Cout, int Cin, int ntaps, int stride, const Fold* fold,
          const std::vector<float>* own_bias, cetpick_simsiam::Conv& c, int CoutP = 0, int CinP = 0) {
  auto w0 = m->get(wkey, (size_t)Cout * Cin * ntaps);
  if (CoutP < Cout) CoutP = Cout;
  if (CinP < Cin) CinP = Cin;
  if (!w0 || (CinP % 64)) return false;
  std::vector<float> wpad;
  std::vector<double> spad;
  std::vector<float> bpad;
  Fold fpad;
  const std::vector<float>* w = w0;
  if (CoutP != Cout || CinP != Cin) {
    wpad.assign((size_t)CoutP * CinP * ntaps, 0.f);
    for (int n = 0; n < Cout; ++n)
      for (int k = 0; k < Cin; ++k)
        for (int t = 0; t < ntaps; ++t) wpad[((size_t)n * CinP + k) * ntaps + t] = (*w0)[((size_t)n * Cin + k) * ntaps + t];
    w = &wpad;
    if (fold) {
      fpad.scale.assign(CoutP, 1.0); fpad.shift.assign(CoutP, 0.0);
      for (int n = 0; n < Cout; ++n) { fpad.scale[n] = fold->scale[n]; fpad.shift[n] = fold->shift[n]; }
      fold = &fpad;
    }
    if (own_bias) {
      bpad.assign(CoutP, 0.f);
      for (int n = 0; n < Cout; ++n) bpad[n] = (*own_bias)[n];
      own_bias = &bpad;
    }
    Cout = CoutP; Cin = CinP;
  }
  const std::vector<uint16_t> pk = small_pack_weights(w->data(), Cout, Cin, ntaps, fold ? fold->scale.data() : nullptr);
  c.w_off = balloc(m, pk.size() * 2);
  memcpy(m->blob.data() + c.w_off, pk.data(), pk.size() * 2);
  c.Cin = Cin; c.Cout = Cout; c.ntaps = ntaps; c.stride = stride;
  c.has_bias = fold || own_bias;
  if (c.has_bias) {
    c.b_off = balloc(m, (size_t)Cout * 4);
    float* b = reinterpret_cast<float*>(m->blob.data() + c.b_off);
    for (int n = 0; n < Cout; ++n) {
      double v = own_bias ? (double)(*own_bias)[n] : 0.0;
      if (fold) v = v * fold->scale[n] + fold->shift[n];
      b[n] = (float)v;
    }
  }
  return true;
}

struct Geo { int h0, h1, h2, h3; };   // conv map, after pool (layer1), layer2, layer3

bool geometry(int64_t H, int64_t W, Geo& g) {
  if (H != W || (H != 32 && H != 16)) return false;     // small maps whose pixel count divides the 128-row M-tile
  g.h0 = (int)H / 2; g.h1 = g.h0 / 2; g.h2 = g.h1 / 2; g.h3 = std::max(1, g.h2 / 2);
  return g.h2 >= 2;
}

// 2-D variant: no stride in conv1 and no max-pool, the stages see H, H/2, H/4 (simsiam_model_2d.py:755-762)
bool geometry_2d(int64_t D, int64_t H, int64_t W, Geo& g) {
  if (D != 1 || H != W || (H != 8 && H != 16 && H != 32 && H != 64)) return false;
  g.h0 = g.h1 = (int)H; g.h2 = g.h1 / 2; g.h3 = g.h2 / 2;
  return true;
}

bool geometry_of(const cetpick_simsiam* m, int64_t D, int64_t H, int64_t W, Geo& g) {
  return m->two_d ? geometry_2d(D, H, W, g) : geometry(H, W, g);
}

size_t ws_bytes_for(const cetpick_simsiam* m, int64_t B, int64_t D, const Geo& g) {
  (void)m;
  const size_t n = (size_t)B * D;
  const size_t a1 = align_up(n * g.h1 * g.h1 * 64 * 2, 1024);       // layer1 maps (three rotating buffers)
  return 3 * a1 + 1024 + align_up((size_t)B * 256 * 2, 1024) * 4 + align_up((size_t)B * 256 * 4, 1024) * 2;
}

int run(const cetpick_simsiam* m, const cetpick_simsiam::Conv& c, const void* src, int B, int Z, int Hin, int Win, int Ho,
        int Wo, bool is3d, const void* residual, int relu, int out_f32, void* out, cudaStream_t st) {
  SmallLaunch L;
  L.src = src; L.C = c.Cin; L.B = B; L.Z = Z; L.Hin = Hin; L.Win = Win; L.stride = c.stride; L.Ho = Ho; L.Wo = Wo;
  L.wpk = static_cast<const uint8_t*>(m->d_blob) + c.w_off; L.N = c.Cout; L.ntaps = c.ntaps;
  int t = 0;
  if (c.ntaps == 1) { L.tap[0][0] = L.tap[0][1] = L.tap[0][2] = 0; }
  else if (c.ntaps == 9) {
    for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) { L.tap[t][0] = 0; L.tap[t][1] = ky - 1; L.tap[t][2] = kx - 1; ++t; }
  } else {
    for (int kz = 0; kz < 3; ++kz) for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) {
      L.tap[t][0] = kz - 1; L.tap[t][1] = ky - 1; L.tap[t][2] = kx - 1; ++t;
    }
  }
  (void)is3d;
  L.bias = c.has_bias ? reinterpret_cast<const float*>(static_cast<const uint8_t*>(m->d_blob) + c.b_off) : nullptr;
  L.residual = residual; L.relu = relu; L.out_f32 = out_f32; L.out = out;
  return conv_small_launch(L, st);
}

}  // namespace

extern "C" int cetpick_simsiam_create(cetpick_simsiam** plan, int blocks1, int blocks2, int blocks3, int has_proj,
                                      int has_pred) {
  if (!plan || blocks1 < 1 || blocks2 < 1 || blocks3 < 1 || blocks1 > 8 || blocks2 > 8 || blocks3 > 8) return CETPICK_ERR_BAD_ARG;
  if (has_pred && !has_proj) return CETPICK_ERR_BAD_ARG;     // 'pred' is applied to the 'proj' output (:360-363)
  cetpick_simsiam* m = new cetpick_simsiam();
  m->layers[0] = blocks1; m->layers[1] = blocks2; m->layers[2] = blocks3;
  m->has_proj = has_proj != 0; m->has_pred = has_pred != 0;
  *plan = m;
  return CETPICK_OK;
}

extern "C" int cetpick_simsiam_create_2d(cetpick_simsiam** plan, int blocks1, int blocks2, int blocks3, int out_dim,
                                         int has_proj, int has_pred) {
  if (out_dim < 1 || out_dim > 256) return plan ? CETPICK_ERR_UNSUPPORTED : CETPICK_ERR_BAD_ARG;
  const int rc = cetpick_simsiam_create(plan, blocks1, blocks2, blocks3, has_proj, has_pred);
  if (rc != CETPICK_OK) return rc;
  (*plan)->two_d = 1;
  (*plan)->out_dim = out_dim;
  return CETPICK_OK;
}

extern "C" void cetpick_simsiam_destroy(cetpick_simsiam* m) {
  if (!m) return;
  if (m->d_blob) cudaFree(m->d_blob);
  delete m;
}

extern "C" int cetpick_simsiam_set_param(cetpick_simsiam* m, const char* key, const float* data, int64_t numel) {
  if (!m || !key || (!data && numel > 0) || numel < 0) return CETPICK_ERR_BAD_ARG;
  const std::string k(key);
  if (k.size() >= 19 && k.compare(k.size() - 19, 19, "num_batches_tracked") == 0) return CETPICK_OK;
  m->params[k].assign(data, data + numel);
  m->finalized = false;
  return CETPICK_OK;
}

extern "C" int cetpick_simsiam_finalize(cetpick_simsiam* m) {
  if (!m) return CETPICK_ERR_BAD_ARG;
  m->blob.clear();
  m->blocks.clear();
  Fold f;
  {
    const int kk = m->two_d ? 9 : 49;                       // conv1: 3x3 s1 (2-D variant) or 7x7 s2
    auto w = m->get("conv1.weight", (size_t)64 * kk);
    if (!w || !bn_fold(m, "bn1", 64, true, f)) return CETPICK_ERR_STATE;
    m->stem_w = balloc(m, (size_t)kk * 64 * 4);
    m->stem_b = balloc(m, 64 * 4);
    float* sw = reinterpret_cast<float*>(m->blob.data() + m->stem_w);
    float* sb = reinterpret_cast<float*>(m->blob.data() + m->stem_b);
    for (int c = 0; c < 64; ++c) {
      for (int t = 0; t < kk; ++t) sw[t * 64 + c] = (float)((double)(*w)[c * kk + t] * f.scale[c]);
      sb[c] = (float)f.shift[c];
    }
  }
  int inpl = 64;
  const int planes[3] = {64, 128, 256};
  for (int li = 0; li < 3; ++li)
    for (int b = 0; b < m->layers[li]; ++b) {
      const std::string p = "layer" + std::to_string(li + 1) + "." + std::to_string(b);
      const int pl = planes[li], cin = b == 0 ? inpl : pl, stride = (li > 0 && b == 0) ? 2 : 1;
      cetpick_simsiam::Block blk;
      if (!bn_fold(m, p + ".bn1", pl, true, f) || !pack(m, p + ".conv1.weight", pl, cin, 9, stride, &f, nullptr, blk.c1)) return CETPICK_ERR_STATE;
      if (!bn_fold(m, p + ".bn2", pl, true, f) || !pack(m, p + ".conv2.weight", pl, pl, 9, 1, &f, nullptr, blk.c2)) return CETPICK_ERR_STATE;
      blk.has_ds = m->params.count(p + ".downsample.0.weight") != 0;
      if (blk.has_ds && !pack(m, p + ".downsample.0.weight", pl, cin, 1, stride, nullptr, nullptr, blk.ds)) return CETPICK_ERR_STATE;
      if (!blk.has_ds && (stride != 1 || cin != pl)) return CETPICK_ERR_STATE;
      m->blocks.push_back(blk);
      inpl = pl;
    }
  if (!m->two_d &&
      (!bn_fold(m, "feature_3d.1", 256, true, f) || !pack(m, "feature_3d.0.weight", 256, 256, 27, 1, &f, nullptr, m->f3d)))
    return CETPICK_ERR_STATE;
  const int od = m->out_dim;                                // 256, or head_conv in the 2-D variant (simsiam_model_2d.py:639-640)
  const int op = (od + 63) / 64 * 64;                       // width the head layers are built with
  {
    auto b = m->get("fc.bias", od);
    if (!b || !pack(m, "fc.weight", od, 256, 1, 1, nullptr, b, m->fc, op, 256)) return CETPICK_ERR_STATE;
  }
  if (m->has_proj) {
    if (!bn_fold(m, "proj.1", od, true, f) || !pack(m, "proj.0.weight", od, od, 1, 1, &f, nullptr, m->proj0, op, op)) return CETPICK_ERR_STATE;
    if (!bn_fold(m, "proj.4", od, true, f) || !pack(m, "proj.3.weight", od, od, 1, 1, &f, nullptr, m->proj3, op, op)) return CETPICK_ERR_STATE;
    if (!bn_fold(m, "proj.7", od, false, f) || !pack(m, "proj.6.weight", od, od, 1, 1, &f, nullptr, m->proj6, op, op)) return CETPICK_ERR_STATE;
  }
  if (m->has_pred) {
    auto b = m->get("pred.3.bias", od);
    if (!bn_fold(m, "pred.1", od, true, f) || !pack(m, "pred.0.weight", od, od, 1, 1, &f, nullptr, m->pred0, op, op)) return CETPICK_ERR_STATE;
    if (!b || !pack(m, "pred.3.weight", od, od, 1, 1, nullptr, b, m->pred3, op, op)) return CETPICK_ERR_STATE;
  }
  if (m->d_blob) { cudaFree(m->d_blob); m->d_blob = nullptr; }
  CETPICK_CUDA(cudaMalloc(&m->d_blob, m->blob.size()));
  CETPICK_CUDA(cudaMemcpy(m->d_blob, m->blob.data(), m->blob.size(), cudaMemcpyHostToDevice));
  m->finalized = true;
  return CETPICK_OK;
}

extern "C" int cetpick_simsiam_workspace_bytes(const cetpick_simsiam* m, int64_t B, int64_t D, int64_t H, int64_t W,
                                               size_t* bytes) {
  Geo g;
  if (!m || !bytes || B <= 0 || D <= 0) return CETPICK_ERR_BAD_ARG;
  if (!geometry_of(m, D, H, W, g)) return CETPICK_ERR_UNSUPPORTED;
  *bytes = ws_bytes_for(m, B, D, g);
  return CETPICK_OK;
}

extern "C" int cetpick_simsiam_forward(cetpick_simsiam* m, const float* x, int64_t B64, int64_t D64, int64_t H, int64_t W,
                                       float* proj, float* pred, void* ws, size_t ws_bytes, void* stream) {
  g_launches = 0;
  if (!m || !x || B64 <= 0 || D64 <= 0 || B64 > (1 << 24) || D64 > 4096) return CETPICK_ERR_BAD_ARG;
  if (!m->finalized) return CETPICK_ERR_STATE;
  if ((proj && !m->has_proj) || (pred && !m->has_pred) || (!proj && !pred)) return CETPICK_ERR_BAD_ARG;
  Geo g;
  if (!geometry_of(m, D64, H, W, g)) return CETPICK_ERR_UNSUPPORTED;
  const int B = (int)B64, D = (int)D64;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~(uintptr_t)1023);
  if (!ws || ws_bytes < ws_bytes_for(m, B, D, g)) return CETPICK_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)B * D;
  const size_t a1 = align_up((size_t)n * g.h1 * g.h1 * 64 * 2, 1024);
  __nv_bfloat16* buf[3] = {reinterpret_cast<__nv_bfloat16*>(base), reinterpret_cast<__nv_bfloat16*>(base + a1),
                           reinterpret_cast<__nv_bfloat16*>(base + 2 * a1)};
  const size_t v1 = align_up((size_t)B * 256 * 2, 1024);
  __nv_bfloat16* vec[4];
  for (int i = 0; i < 4; ++i) vec[i] = reinterpret_cast<__nv_bfloat16*>(base + 3 * a1 + (size_t)i * v1);
  const uint8_t* blob = static_cast<const uint8_t*>(m->d_blob);
  int rc;

  if (m->two_d) {  // conv1 3x3 + bn1 + relu -> buf[0]: (n, H, H, 64)
    const int grid = (int)std::min<long long>(n, (long long)num_sms() * 4);
    const size_t smem = (size_t)(H + 2) * (H + 2) * 4;
    stem2d_kernel<<<grid, 256, smem, st>>>(x, n, (int)H, reinterpret_cast<const float*>(blob + m->stem_w),
                                           reinterpret_cast<const float*>(blob + m->stem_b), buf[0]);
    CETPICK_LAUNCH_CHECK();
  } else {  // conv1 + bn1 + relu + maxpool -> buf[0]: (n, h1, h1, 64)
    const int grid = (int)std::min<long long>(n, (long long)num_sms() * 8);
    const float* sw = reinterpret_cast<const float*>(blob + m->stem_w);
    const float* sb = reinterpret_cast<const float*>(blob + m->stem_b);
    auto smem_of = [](int hw) { const int ip = hw + 6, cw = hw / 2; return 49 * 64 * 4 + ((ip * (ip + 1) * 4 + 15) & ~15) + cw * cw * 66 * 2; };
    static DeviceOnce attr_once;
    if (attr_once.first()) {
      CETPICK_CUDA(cudaFuncSetAttribute(stem_pool_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(32)));
      CETPICK_CUDA(cudaFuncSetAttribute(stem_pool_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_of(16)));
    }
    static DeviceOnce mma_once;
    if (mma_once.first())
      CETPICK_CUDA(cudaFuncSetAttribute(stem_pool_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM));
    static const bool fp32_stem = getenv("CETPICK_SIMSIAM_FP32_STEM") != nullptr;     // A/B switch for the profiles
    if (H == 32 && !fp32_stem) {
      const int gm = (int)std::min<long long>(n, (long long)num_sms() * 3);
      stem_pool_mma_kernel<<<gm, 256, SM_SMEM, st>>>(x, n, sw, sb, buf[0]);
    } else if (H == 32) stem_pool_kernel<32><<<grid, 256, smem_of(32), st>>>(x, n, sw, sb, buf[0]);
    else stem_pool_kernel<16><<<grid, 256, smem_of(16), st>>>(x, n, sw, sb, buf[0]);
    CETPICK_LAUNCH_CHECK();
  }
  // BasicBlocks: x -> relu(bn1(conv1(x))) -> bn2(conv2(.)) + (x | downsample(x)) -> relu
  int cur = 0, hin = g.h1, bi = 0;
  for (int li = 0; li < 3; ++li)
    for (int b = 0; b < m->layers[li]; ++b, ++bi) {
      const cetpick_simsiam::Block& blk = m->blocks[bi];
      const int hout = blk.c1.stride == 2 ? hin / 2 : hin;
      if (hout < 1) return CETPICK_ERR_UNSUPPORTED;
      const int t1 = (cur + 1) % 3, t2 = (cur + 2) % 3;
      if ((rc = run(m, blk.c1, buf[cur], 1, (int)n, hin, hin, hout, hout, false, nullptr, 1, 0, buf[t1], st))) return rc;
      const void* res = buf[cur];
      if (blk.has_ds) {
        if ((rc = run(m, blk.ds, buf[cur], 1, (int)n, hin, hin, hout, hout, false, nullptr, 0, 0, buf[t2], st))) return rc;
        res = buf[t2];
      }
      // conv2 + bn2 + residual + relu; the output may overwrite the block input unless that IS the residual
      __nv_bfloat16* dst = blk.has_ds ? buf[cur] : buf[t2];
      if ((rc = run(m, blk.c2, buf[t1], 1, (int)n, hout, hout, hout, hout, false, res, 1, 0, dst, st))) return rc;
      cur = blk.has_ds ? cur : t2;
      hin = hout;
    }
  const int P = D * hin * hin;
  if (m->two_d) {
    // AdaptiveAvgPool2d((1,1)) over the last stage's map (simsiam_model_2d.py:764-765)
    avgpool_kernel<<<std::min(B, num_sms() * 8), 256, 0, st>>>(buf[cur], B, P, 256, vec[0]);
    CETPICK_LAUNCH_CHECK();
  } else {
    // (B, D, h, w, 256) -> Conv3d 3x3x3 pad 1 + BN3d + ReLU (:348-354); one sub-volume (D*h*w positions) per M-tile
    // an M-tile holds up to 128 / (h*w) slices of one sub-volume (a 32^3 sub-volume fills it exactly; shallower ones, like the
    // D = 1 slab sums of the reference's own exploration dataset, leave rows idle; deeper ones span several tiles)
    if (hin * hin > 128) return CETPICK_ERR_UNSUPPORTED;
    const int t1 = (cur + 1) % 3;
    if ((rc = run(m, m->f3d, buf[cur], B, D, hin, hin, hin, hin, true, nullptr, 1, 0, buf[t1], st))) return rc;
    avgpool_kernel<<<std::min(B, num_sms() * 8), 256, 0, st>>>(buf[t1], B, P, 256, vec[0]);
    CETPICK_LAUNCH_CHECK();
  }
  // fc, then the heads: Linear (+ folded BatchNorm1d) (+ ReLU) as 1x1 "convolutions" over the batch axis
  if ((rc = run(m, m->fc, vec[0], 1, B, 1, 1, 1, 1, false, nullptr, 0, 0, vec[1], st))) return rc;
  if (m->has_proj) {
    if ((rc = run(m, m->proj0, vec[1], 1, B, 1, 1, 1, 1, false, nullptr, 1, 0, vec[2], st))) return rc;
    if ((rc = run(m, m->proj3, vec[2], 1, B, 1, 1, 1, 1, false, nullptr, 1, 0, vec[3], st))) return rc;
    // proj.6 + BN(affine=False): fp32 out for the caller, bf16 copy as the input of 'pred'
    // head widths that are not multiples of 64 are computed 64-padded into a staging buffer and copied out row by row
    const int od = m->out_dim, op = m->proj6.Cout;
    float* stage_f32[2] = {reinterpret_cast<float*>(base + 3 * a1 + 4 * v1),
                           reinterpret_cast<float*>(base + 3 * a1 + 4 * v1 + align_up((size_t)B * 256 * 4, 1024))};
    auto emit = [&](const cetpick_simsiam::Conv& c, const void* src, float* dst, int which) -> int {
      float* o = op == od ? dst : stage_f32[which];
      const int r = run(m, c, src, 1, B, 1, 1, 1, 1, false, nullptr, 0, 1, o, st);
      if (r) return r;
      if (op != od)
        CETPICK_CUDA(cudaMemcpy2DAsync(dst, (size_t)od * 4, o, (size_t)op * 4, (size_t)od * 4, (size_t)B, cudaMemcpyDeviceToDevice, st));
      return CETPICK_OK;
    };
    if (proj && (rc = emit(m->proj6, vec[3], proj, 0))) return rc;
    if (pred) {
      if ((rc = run(m, m->proj6, vec[3], 1, B, 1, 1, 1, 1, false, nullptr, 0, 0, vec[2], st))) return rc;
      if ((rc = run(m, m->pred0, vec[2], 1, B, 1, 1, 1, 1, false, nullptr, 1, 0, vec[0], st))) return rc;
      if ((rc = emit(m->pred3, vec[0], pred, 1))) return rc;
    }
  }
  return CETPICK_OK;
}
