// Host-side description of one "marching" convolution launch (conv_march.cu): the kernel used for
// the narrow layers (Cout 32 / 64) of the detector, 2-D 3x3 and 3-D 3x3x3 dilated (1,d,d).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace cetpick {

enum MarchMode {
  MARCH_2D_ROWS = 0,    // Conv2d 3x3 pad 1 per image: march along y, M-tile = 128 pixels of one row
  MARCH_3D_PLANES = 1   // Conv3d 3x3x3 dilation (1,d,d) pad (1,d,d): march along z, M-tile = 8 x 16 (x,y)
};

struct MarchLaunch {
  int mode = MARCH_2D_ROWS;
  int dil = 1;              // in-plane dilation of the 3-D mode (4 for feature_head)
  int nsrc = 1;             // two sources of equal C = torch.cat((up, skip), 1) (unet.py:390)
  const void* src[2] = {nullptr, nullptr};   // bf16 [NIMG][H][W][C]
  int C = 0;                // channels per source: 16 / 32 / 64
  int NIMG = 0, H = 0, W = 0;
  int z_origin = 0;         // 3-D mode: absolute z of plane 0 (z-slab streaming / z-sharding keep results bit-identical)
  const void* wpk = nullptr;    // device, layout of march_pack_weights()
  int Cout = 0;             // 32 or 64
  const float* bias = nullptr;  // [Cout] fp32 or null (device)
  const float* bias_host = nullptr;      // the same values on the host: they travel in the kernel parameters
  int relu = 0;
  void* out = nullptr;      // bf16 [NIMG][H][W][Cout] (may be null when hm_out is set)
  // 2-D mode: also write MaxPool2d(2, ceil_mode=True) of the output (unet.py:225,237-238) as bf16
  // [NIMG][(H+1)/2][(W+1)/2][Cout] from the epilogue registers (needs relu: padding compares as 0)
  void* pool_out = nullptr;
  // 3-D mode extras -------------------------------------------------------------------------
  // bias_tab: [64][Cout] fp32 or null: a bias that depends on which taps fall inside the volume,
  // row = (cz*4 + cy)*4 + cx with c = (lower neighbour inside) + 2*(upper neighbour inside) per axis.
  // This is how a 1x1 conv with bias in FRONT of a zero-padded conv is folded into it exactly
  // (conv_final -> feature_head.0, unet.py:882 -> unet_small.py:85).
  const float* bias_tab = nullptr;
  const float* bias_tab_host = nullptr;
  // fused `hm` head (unet_small.py:53-61,89): Conv3d(Cout,1,(3,1,1),pad (1,0,0)) on the ReLU output,
  // kept in fp32 registers across the march; hm_w = [3][Cout] fp32 (kz major), hm_out = (D,H,W) fp32.
  const float* hm_w = nullptr;
  const float* hm_w_host = nullptr;
  float* hm_out = nullptr;
  int hm_sigmoid = 0;       // models/utils.py:167-169 _sigmoid fused
};

// True when conv_march_launch supports (mode, C per source, nsrc, Cout): weights must fit in shared
// memory next to the activation ring.
bool march_supported(int mode, int C, int nsrc, int Cout);

// Pack a PyTorch-layout weight into the kernel's shared-memory image.
//   2-D: w = (Cout, nsrc*C, 3, 3);  3-D: w = (Cout, nsrc*C, 3, 3, 3);  scale[Cout] (BN fold) or null.
// Layout: [block b = (source, channel chunk, in-step tap j)][row n = slot*Cout + co][KC] bf16 where
// slot 0/1/2 are the outputs one step behind / at / one step ahead of the input row (kernel index
// 2 - slot along the march axis) and j is kx (2-D) or ky*3+kx (3-D).
std::vector<uint16_t> march_pack_weights(int mode, const float* w, int Cout, int nsrc, int C,
                                         const double* scale);

int conv_march_launch(const MarchLaunch& L, cudaStream_t stream);

}  // namespace cetpick
