// Host-side description of one halo-tile 3x3 convolution launch (conv_halo.cu): the wide trunk levels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace cetpick {

struct HaloLaunch {
  int nsrc = 1;                 // two sources of equal C = torch.cat((up, skip), 1) (unet.py:390)
  const void* src[2] = {nullptr, nullptr};   // bf16 [NIMG][H][W][C]
  int C = 0;                    // channels per source, multiple of 64
  int NIMG = 0, H = 0, W = 0;
  const void* wpk = nullptr;    // device, layout of halo_pack_weights()
  const float* bias = nullptr;  // device [Cout] fp32 (BN folded)
  int Cout = 0;                 // multiple of 128
  int relu = 0;
  void* out = nullptr;          // bf16 [NIMG][H][W][Cout]
};

bool halo_supported(int C, int nsrc, int Cout);

// w = PyTorch Conv2d weight (Cout, nsrc*C, 3, 3); scale[Cout] (BN fold) or null.
// Layout: [(source, chunk of 64, tap ky*3+kx, block of 128 output channels)][128][64] bf16.
std::vector<uint16_t> halo_pack_weights(const float* w, int Cout, int nsrc, int C, const double* scale);

int conv_halo_launch(const HaloLaunch& L, cudaStream_t stream);

}  // namespace cetpick
