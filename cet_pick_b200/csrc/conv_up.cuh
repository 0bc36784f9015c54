// Host-side description of one ConvTranspose2d(k2,s2)+BN+ReLU launch (conv_up.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace cetpick {

struct UpLaunch {
  const void* src = nullptr;    // bf16 [NIMG][h][w][Cin]
  int Cin = 0, NIMG = 0, h = 0, w = 0;
  const void* wpk = nullptr;    // device, layout of upconv_pack_weights()
  const float* bias = nullptr;  // device [4*Cout] fp32, (dy,dx,co) order (BN folded)
  int Cout = 0;
  void* out = nullptr;          // bf16 [NIMG][Ho][Wo][Cout], Ho <= 2h, Wo <= 2w (autocrop)
  int Ho = 0, Wo = 0;
};

bool upconv_supported(int Cin, int Cout);

// w = PyTorch ConvTranspose2d weight (Cin, Cout, 2, 2); scale[Cout] (BN fold) or null.
// Layout: [column split][channel chunk of 64][NB columns (dy,dx,co)][64] bf16, NB = min(4*Cout, 256).
std::vector<uint16_t> upconv_pack_weights(const float* w, int Cin, int Cout, const double* scale);

int conv_up_launch(const UpLaunch& L, cudaStream_t stream);

}  // namespace cetpick
