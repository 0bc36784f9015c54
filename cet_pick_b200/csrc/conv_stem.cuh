// Host-side description of the tensor-core stem launch (conv_stem.cu):
// Conv2d(1,16,7,stride 2,pad 3,bias=False) + BN + ReLU of cet_pick/models/networks/unet_small.py:35-37,72-74.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

namespace cetpick {

struct StemLaunch {
  const float* in = nullptr;     // fp32 (D,H,W), rows 16-byte aligned (W % 4 == 0) -- or:
  const uint8_t* in_u8 = nullptr;   // uint8 levels (D,H,W), rows 16-byte aligned (W % 16 == 0), with
  uint16_t lut[256] = {};           // lut[k] = bf16 bits of level k's value (lut[0] must be 0)
  int D = 0, H = 0, W = 0;
  const void* wpk = nullptr;     // device, layout of stem_pack_weights()
  float bias[16] = {};           // BN shift (host values: they travel in the kernel parameters)
  void* out = nullptr;           // bf16 (D,h,w,16), h = (H-1)/2+1, w = (W-1)/2+1
};

// true when the TMA path can read `in` (16-byte aligned base and row pitch)
bool stem_tc_supported(const float* in, int W);
bool stem_tc_supported_u8(const uint8_t* in, int W);

// [4 row slots][16 co][16 k] bf16: slot d = output row (j-1)+d fed by input row pair j; k = e*8 + c is
// input row 2j+e, input column 2*ox-4+c; taps outside the 7x7 kernel are zero.  scale[16] = BN scale.
std::vector<uint16_t> stem_pack_weights(const float* w /*[16][49]*/, const double* scale);

int conv_stem_launch(const StemLaunch& L, cudaStream_t stream);

}  // namespace cetpick
