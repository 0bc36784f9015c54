// Host-side description of one launch of the small-map implicit-GEMM convolution (conv_small.cu): the BasicBlock
// convolutions, the 3-D feature layer and the Linear layers of cet_pick/models/networks/simsiam_model.py:44-73,
// 181-215, 325-366.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace cetpick {

struct SmallLaunch {
  const void* src = nullptr;     // bf16 [B][Z][Hin][Win][C]
  int C = 0;                     // input channels, multiple of 64
  int B = 1, Z = 1;              // batch elements, maps per batch element (taps may move in z inside one element)
  int Hin = 1, Win = 1;
  int stride = 1;                // 1 or 2 (in x and y)
  int Ho = 1, Wo = 1;            // output map size: Wo*Ho divides 128 (whole maps per tile), or Wo divides 128 and
                                 // Ho is a multiple of 128 / Wo (row bands of one map per tile)
  const void* wpk = nullptr;     // device, layout of small_pack_weights()
  int N = 0;                     // output channels, multiple of 16, <= 256
  int ntaps = 1;
  int tap[27][3] = {};           // (dz, dy, dx): input offset of tap t relative to stride * output position
  const float* bias = nullptr;   // [N] fp32 device or null
  const void* residual = nullptr;   // bf16 [B][Z][Ho][Wo][N] added before the ReLU, or null
  int relu = 0;
  int out_f32 = 0;
  void* out = nullptr;           // bf16 (or fp32) [B][Z][Ho][Wo][N]
};

// (Cout, Cin, ntaps) PyTorch layout -> [tap][64-channel chunk][Cout][64] bf16, scale[Cout] folded in
std::vector<uint16_t> small_pack_weights(const float* w, int Cout, int Cin, int ntaps, const double* scale);

int conv_small_launch(const SmallLaunch& L, cudaStream_t stream);

}  // namespace cetpick
