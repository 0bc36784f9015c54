// Host-side description of one fused block launch (conv_block.cu): conv3x3+BN+ReLU -> conv3x3+BN+ReLU
// (-> MaxPool2d(2, ceil)) of the 32-channel full-resolution blocks, cet_pick/models/networks/unet.py:198-249
// (DownConv level 0) and :375-399 (last UpConv after the concat).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace cetpick {

struct BlockLaunch {
  int nsrc = 1;                              // conv1 reads the concat of two sources (torch.cat((up, skip), 1))
  const void* src[2] = {nullptr, nullptr};   // bf16 [NIMG][H][W][C1]
  int C1 = 0;                                // channels per source of conv1: 16 (nsrc 1) or 32 (nsrc 1 / 2)
  int NIMG = 0, H = 0, W = 0;
  const void* w1pk = nullptr;                // device: march_pack_weights(MARCH_2D_ROWS, w1, 32, nsrc, C1, scale1)
  const void* w2pk = nullptr;                // device: march_pack_weights(MARCH_2D_ROWS, w2, 32, 1, 32, scale2)
  const float* bias1_host = nullptr;         // [32] folded BN shift of conv1 (HOST: travels in the kernel parameters)
  const float* bias2_host = nullptr;         // [32] of conv2
  void* out = nullptr;                       // bf16 [NIMG][H][W][32]: conv2 output (the skip / block output)
  void* pool_out = nullptr;                  // optional bf16 [NIMG][(H+1)/2][(W+1)/2][32]: fused MaxPool2d(2, ceil)
};

// conv1 channel configuration and a row width one thread-block cluster (<= 8 CTAs x 128 pixels) can span
bool block_supported(int C1, int nsrc, int W);

int conv_block_launch(const BlockLaunch& L, cudaStream_t stream);

}  // namespace cetpick
