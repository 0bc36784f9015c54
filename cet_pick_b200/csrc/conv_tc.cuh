// Host-side description of one implicit-GEMM convolution launch (conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace cetpick {

enum ConvEpilogue {
  EPI_BF16_NHWC = 0,     // out[img][y][x][col]            bf16, + bias, optional ReLU
  EPI_UPCONV_2X2 = 1,    // out[img][2y+dy][2x+dx][col%C]  bf16 (ConvTranspose2d k2 s2), + bias, ReLU
  EPI_F32_ROWMAJOR = 2,  // out[pixel][col]                fp32 (self-test)
  EPI_F32_L2NORM_NCDHW = 3  // out[col][img][y][x]         fp32, L2-normalised over the columns
};

struct ConvLaunch {
  // sources: bf16 activations [NIMG][H][W][C[i]] (NHWC; for 3-D convs NIMG is the z axis).
  // The K loop runs over source 0 then source 1 (the reference's torch.cat((up, skip), 1)).
  int nsrc = 1;
  const void* src[2] = {nullptr, nullptr};
  int C[2] = {0, 0};
  int NIMG = 0, H = 0, W = 0;
  // weights: bf16 [k-block][Ntot][KC], k-block = ((source, tap, channel chunk)) in loop order
  const void* wpk = nullptr;
  int KC = 64;              // channels per k-block: 16 / 32 / 64 (selects the 32/64/128-byte swizzle); tf32: 8 / 16 / 32
  int tf32 = 0;             // 1: sources, weights and NHWC outputs are fp32 holding TF32 values (kind::tf32 UMMA)
  int ntaps = 1;
  int tap[27][3] = {};      // (dz, dy, dx) input offsets per tap
  int Ntot = 0;             // GEMM N (output columns), multiple of 16
  const float* bias = nullptr;  // [Ntot] fp32 or null
  int relu = 0;
  int epi = EPI_BF16_NHWC;
  void* out = nullptr;
  int out_cstride = 0;      // channels per output pixel (EPI_BF16_NHWC)
  int Ho = 0, Wo = 0, Cout = 0;  // EPI_UPCONV_2X2: output size (after autocrop) and channels
};

// bf16 tiled tensor map (rank <= 5) whose innermost box dimension is KC channels = the swizzle span
// (KC 16 / 32 / 64 -> SWIZZLE_32B / 64B / 128B); out-of-bounds elements read as zero.
// (hits, misses) of the process-wide tensor-map cache behind the tmap_encode_* helpers
void tmap_cache_stats(int64_t* hits, int64_t* misses);
int tmap_encode_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int KC);

// fp32 tiled tensor map without swizzle whose out-of-bounds elements read as NaN.
int tmap_encode_f32_nanfill(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box);

// fp32 tiled tensor map without swizzle whose out-of-bounds elements read as zero.
int tmap_encode_u8_zerofill(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box);
int tmap_encode_f32_zerofill(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box);

// Enqueue the convolution.  Returns a CETPICK_* code.
int conv_tc_launch(const ConvLaunch& L, cudaStream_t stream);

}  // namespace cetpick
