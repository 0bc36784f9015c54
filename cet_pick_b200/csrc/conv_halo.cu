// 3x3 Conv2d (pad 1) + BN + ReLU for the wide trunk levels (Cout a multiple of 128, C per source a
// multiple of 64): cet_pick/models/networks/unet.py:127-145 as used by DownConv (:198-249) and UpConv
// (:319-399) at the 128 / 256 (/512) channel levels.
//
// Implicit GEMM with the activation operand loaded ONCE per (source, 64-channel chunk): the CTA owns a
// 16 x 16 pixel tile (two 8 x 16 UMMA M-tiles side by side) and 128 output channels.  One TMA box brings
// the (16+2) x (16+2) halo tile of a chunk into shared memory (out-of-bounds = the conv's zero padding);
// the nine taps are nine SHIFTED UMMA descriptors into that one tile (8-pixel row groups, stride = the
// halo row pitch), so L2->smem traffic per tile is 1.27x the tile instead of 9x, and only the weights
// (16 KB per tap and chunk, shared by both M-tiles) stream through their own ring.  Accumulators
// (2 M-tiles x 128 columns) are double-buffered in TMEM so the epilogue of tile i (bias, ReLU, bf16,
// NHWC store, 8 warps) overlaps the MMAs of tile i+1.  Two sources = torch.cat((up, skip), 1).
//   warp 0: TMA producer   warp 1: UMMA issuer (warp-uniform loop)   warp 2: TMEM allocator
//   warps 4-11: epilogue (warps 4-7: M-tile 0, warps 8-11: M-tile 1)
#include "conv_halo.cuh"
#include "common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <algorithm>

namespace cetpick {

namespace {

constexpr int HALO_THREADS = 384;
constexpr int KC = 64, PIX = KC * 2;            // channels per chunk, bytes per pixel chunk (128-byte swizzle)
constexpr int TW = 16, TH = 16;                 // output tile
constexpr int BX = TW + 2, BY = TH + 2;         // halo tile
constexpr int A_BYTES = BX * BY * PIX;          // 41 472
constexpr int A_STAGE = (A_BYTES + 1023) / 1024 * 1024;
constexpr int NB = 128;                         // output channels per CTA tile
constexpr int B_STAGE = NB * PIX;               // 16 384: one tap of one chunk
constexpr int A_STAGES = 2, B_STAGES = 6;
constexpr int SBO_A = BX * PIX, SBO_B = 8 * PIX;

struct alignas(64) HaloParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  int nsrc, chunks, n_nb;
  int NIMG, H, W, tiles_x, tiles_y, Cout;
  long long total_tiles;
  int relu;
  const float* bias;
  __nv_bfloat16* out;
};

__device__ __forceinline__ void decode_tile(const HaloParams& p, long long t, int& img, int& y0, int& x0, int& nb) {
  nb = (int)(t % p.n_nb);
  t /= p.n_nb;
  x0 = (int)(t % p.tiles_x) * TW;
  t /= p.tiles_x;
  y0 = (int)(t % p.tiles_y) * TH;
  img = (int)(t / p.tiles_y);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(HALO_THREADS, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull[A_STAGES], bar_aempty[A_STAGES], bar_bfull[B_STAGES], bar_bempty[B_STAGES];
  __shared__ __align__(8) uint64_t bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t s_tmem_base;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_STAGES * A_STAGE;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkc = p.nsrc * p.chunks;             // (source, chunk) steps per tile

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tmA[0]);
    if (p.nsrc > 1) ptx::prefetch_tensormap(&p.tmA[1]);
    ptx::prefetch_tensormap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < A_STAGES; ++s) { ptx::mbar_init(&bar_afull[s], 1); ptx::mbar_init(&bar_aempty[s], 1); }
    for (int s = 0; s < B_STAGES; ++s) { ptx::mbar_init(&bar_bfull[s], 1); ptx::mbar_init(&bar_bempty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&bar_tfull[a], 1); ptx::mbar_init(&bar_tempty[a], 8); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&s_tmem_base, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        int img, y0, x0, nb;
        decode_tile(p, t, img, y0, x0, nb);
        for (int kc = 0; kc < nkc; ++kc) {
          const int src = kc / p.chunks, c = kc - src * p.chunks;
          ptx::mbar_wait(&bar_aempty[as], aph ^ 1u);
          ptx::mbar_arrive_expect_tx(&bar_afull[as], (uint32_t)A_BYTES);
          ptx::tma_load_4d(sA + (size_t)as * A_STAGE, &p.tmA[src], &bar_afull[as], c * KC, x0 - 1, y0 - 1, img);
          if (++as == A_STAGES) { as = 0; aph ^= 1u; }
          for (int tap = 0; tap < 9; ++tap) {
            ptx::mbar_wait(&bar_bempty[bs], bph ^ 1u);
            ptx::mbar_arrive_expect_tx(&bar_bfull[bs], (uint32_t)B_STAGE);
            ptx::tma_load_2d(sB + (size_t)bs * B_STAGE, &p.tmB, &bar_bfull[bs], 0,
                             ((kc * 9 + tap) * p.n_nb + nb) * NB);
            if (++bs == B_STAGES) { bs = 0; bph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ UMMA issuer =================================
    constexpr uint32_t A_HI = ptx::smem_desc_hi(SBO_A, 2), B_HI = ptx::smem_desc_hi(SBO_B, 2);
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sA_lo = ptx::smem_desc_lo(ptx::smem_u32(sA)), sB_lo = ptx::smem_desc_lo(ptx::smem_u32(sB));
    int as = 0, bs = 0, acc = 0;
    uint32_t aph = 0, bph = 0, acc_phase = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      ptx::mbar_wait(&bar_tempty[acc], acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(acc * 2 * NB);
      for (int kc = 0; kc < nkc; ++kc) {
        ptx::mbar_wait(&bar_afull[as], aph);
        ptx::tc_fence_after();
        const uint32_t a_lo = sA_lo + (uint32_t)(as * (A_STAGE >> 4));
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          ptx::mbar_wait(&bar_bfull[bs], bph);
          ptx::tc_fence_after();
          const uint32_t b_lo = sB_lo + (uint32_t)(bs * (B_STAGE >> 4));
          const int ky = tap / 3, kx = tap - ky * 3;
          const uint32_t a_tap = a_lo + (uint32_t)(((ky * BX + kx) * PIX) >> 4);
          if (ptx::elect_one()) {
            if (kc == 0 && tap == 0) {
              // first MMA of each accumulator overwrites
              ptx::umma_bf16(d0, ((uint64_t)A_HI << 32) | a_tap, ((uint64_t)B_HI << 32) | b_lo, IDESC, 0u);
              ptx::umma_bf16(d0 + NB, ((uint64_t)A_HI << 32) | (a_tap + ((8 * PIX) >> 4)), ((uint64_t)B_HI << 32) | b_lo, IDESC, 0u);
#pragma unroll
              for (int kk = 1; kk < KC / 16; ++kk) {
                ptx::umma_bf16_lohi(d0, a_tap + kk * 2, A_HI, b_lo + kk * 2, B_HI, IDESC);
                ptx::umma_bf16_lohi(d0 + NB, a_tap + ((8 * PIX) >> 4) + kk * 2, A_HI, b_lo + kk * 2, B_HI, IDESC);
              }
            } else {
#pragma unroll
              for (int kk = 0; kk < KC / 16; ++kk) {
                ptx::umma_bf16_lohi(d0, a_tap + kk * 2, A_HI, b_lo + kk * 2, B_HI, IDESC);
                ptx::umma_bf16_lohi(d0 + NB, a_tap + ((8 * PIX) >> 4) + kk * 2, A_HI, b_lo + kk * 2, B_HI, IDESC);
              }
            }
            ptx::umma_commit(&bar_bempty[bs]);
            if (tap == 8) {
              ptx::umma_commit(&bar_aempty[as]);
              if (kc == nkc - 1) ptx::umma_commit(&bar_tfull[acc]);
            }
          }
          __syncwarp();
          if (++bs == B_STAGES) { bs = 0; bph ^= 1u; }
        }
        if (++as == A_STAGES) { as = 0; aph ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int quad = warp & 3, mt = (warp - 4) >> 2;     // TMEM lane quadrant, M-tile
    const int m = quad * 32 + lane;                      // row of the M-tile: pixel (m & 7, m >> 3)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int img, y0, x0, nb;
      decode_tile(p, t, img, y0, x0, nb);
      const int x = x0 + mt * 8 + (m & 7), y = y0 + (m >> 3);
      const bool valid = (x < p.W) && (y < p.H);
      ptx::mbar_wait(&bar_tfull[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 2 * NB + mt * NB);
      __nv_bfloat16* dst = p.out + (((size_t)img * p.H + y) * p.W + x) * p.Cout + nb * NB;
      const float* bias = p.bias + nb * NB;
#pragma unroll 1
      for (int c0 = 0; c0 < NB; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld16(t_row + c0, v);
        ptx::tmem_ld16(t_row + c0 + 16, v + 16);
        ptx::tmem_ld_wait();
        if (c0 + 32 == NB) {                 // last read of this accumulator: hand it back early
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bar_tempty[acc]);
        }
        if (valid) {
          uint4 wprev = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; i += 4) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + g * 8 + i));
              f[i] = __uint_as_float(v[g * 8 + i]) + b.x;
              f[i + 1] = __uint_as_float(v[g * 8 + i + 1]) + b.y;
              f[i + 2] = __uint_as_float(v[g * 8 + i + 2]) + b.z;
              f[i + 3] = __uint_as_float(v[g * 8 + i + 3]) + b.w;
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            uint4 w;
            w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]);
            w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
            if (g & 1) ptx::st_global_256(reinterpret_cast<uint4*>(dst + c0) + g - 1, wprev, w);   // one full sector per store
            else wprev = w;
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool halo_supported(int C, int nsrc, int Cout) {
  return C > 0 && (C % KC) == 0 && Cout > 0 && (Cout % NB) == 0 && nsrc >= 1 && nsrc <= 2;
}

std::vector<uint16_t> halo_pack_weights(const float* w, int Cout, int nsrc, int C, const double* scale) {
  const int chunks = C / KC, n_nb = Cout / NB, Cin = nsrc * C;
  std::vector<uint16_t> out((size_t)9 * Cin * Cout);
  for (int s = 0; s < nsrc; ++s)
    for (int c = 0; c < chunks; ++c)
      for (int tap = 0; tap < 9; ++tap)
        for (int nb = 0; nb < n_nb; ++nb)
          for (int j = 0; j < NB; ++j)
            for (int k = 0; k < KC; ++k) {
              const int co = nb * NB + j, ci = s * C + c * KC + k;
              const double v = (double)w[((size_t)co * Cin + ci) * 9 + tap] * (scale ? scale[co] : 1.0);
              const size_t blk = ((size_t)((s * chunks + c) * 9 + tap) * n_nb + nb);
              out[(blk * NB + j) * KC + k] = f2bf_host((float)v);
            }
  return out;
}

int conv_halo_launch(const HaloLaunch& L, cudaStream_t stream) {
  if (!halo_supported(L.C, L.nsrc, L.Cout)) return CETPICK_ERR_UNSUPPORTED;
  if (!L.src[0] || (L.nsrc > 1 && !L.src[1]) || !L.wpk || !L.bias || !L.out || L.NIMG <= 0 || L.H <= 0 || L.W <= 0)
    return CETPICK_ERR_BAD_ARG;
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.nsrc = L.nsrc; p.chunks = L.C / KC; p.n_nb = L.Cout / NB;
  p.NIMG = L.NIMG; p.H = L.H; p.W = L.W; p.Cout = L.Cout;
  p.tiles_x = ceil_div(L.W, TW); p.tiles_y = ceil_div(L.H, TH);
  p.total_tiles = (long long)p.tiles_x * p.tiles_y * L.NIMG * p.n_nb;
  p.relu = L.relu; p.bias = L.bias; p.out = static_cast<__nv_bfloat16*>(L.out);
  int rc;
  for (int s = 0; s < L.nsrc; ++s) {
    const uint64_t C = (uint64_t)L.C;
    const uint64_t dims[4] = {C, (uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.NIMG};
    const uint64_t strides[3] = {C * 2, C * 2 * L.W, C * 2 * (uint64_t)L.W * L.H};
    const uint32_t box[4] = {KC, BX, BY, 1};
    if ((rc = tmap_encode_bf16(&p.tmA[s], L.src[s], 4, dims, strides, box, KC))) return rc;
  }
  {
    const uint64_t rows = (uint64_t)L.nsrc * p.chunks * 9 * p.n_nb * NB;
    const uint64_t dims[2] = {KC, rows};
    const uint64_t strides[1] = {PIX};
    const uint32_t box[2] = {KC, NB};
    if ((rc = tmap_encode_bf16(&p.tmB, L.wpk, 2, dims, strides, box, KC))) return rc;
  }
  static DeviceOnce attr_once;
  static int static_smem = 0;
  if (attr_once.first()) {
    cudaFuncAttributes fa;
    CETPICK_CUDA(cudaFuncGetAttributes(&fa, conv_halo_kernel));
    CETPICK_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024 - (int)fa.sharedSizeBytes));
    static_smem = (int)fa.sharedSizeBytes;
  }
  const size_t smem = 1024 + (size_t)A_STAGES * A_STAGE + (size_t)B_STAGES * B_STAGE;
  const int grid = (int)std::min<long long>(p.total_tiles, num_sms());
  conv_halo_kernel<<<grid, HALO_THREADS, smem, stream>>>(p);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

}  // namespace cetpick

using namespace cetpick;

#ifdef CETPICK_TEST_HOOKS   // test / tuning hooks: built into libcetpick_test_sm100a.so only (include/cetpick_test.h)
// Test hook: one 3x3 conv through conv_halo.cu from a PyTorch-layout fp32 HOST weight (Cout, nsrc*C, 3, 3)
// and a HOST bias [Cout] (packs, uploads, launches, synchronises) -- tests/test_gpu_conv.py.
extern "C" int cetpick_conv_halo_bf16(int nsrc, const void* src0, const void* src1, int C, int NIMG, int H, int W,
                                      const float* w_host, const float* bias_host, int Cout, int relu, void* out,
                                      void* stream) {
  g_launches = 0;
  if (!w_host || !bias_host) return CETPICK_ERR_BAD_ARG;
  if (!halo_supported(C, nsrc, Cout)) return CETPICK_ERR_UNSUPPORTED;
  std::vector<uint16_t> pk = halo_pack_weights(w_host, Cout, nsrc, C, nullptr);
  void *dw = nullptr, *db = nullptr;
  CETPICK_CUDA(cudaMalloc(&dw, pk.size() * 2));
  if (cudaMalloc(&db, (size_t)Cout * 4) != cudaSuccess) { cudaFree(dw); return CETPICK_ERR_CUDA; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = CETPICK_OK;
  if (cudaMemcpyAsync(dw, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(db, bias_host, (size_t)Cout * 4, cudaMemcpyHostToDevice, st) != cudaSuccess)
    rc = CETPICK_ERR_CUDA;
  if (rc == CETPICK_OK) {
    HaloLaunch L;
    L.nsrc = nsrc; L.src[0] = src0; L.src[1] = src1; L.C = C; L.NIMG = NIMG; L.H = H; L.W = W;
    L.wpk = dw; L.bias = static_cast<const float*>(db); L.Cout = Cout; L.relu = relu; L.out = out;
    rc = conv_halo_launch(L, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dw);
  cudaFree(db);
  if (rc == CETPICK_OK && e != cudaSuccess) return cuda_fail(e, "conv_halo");
  return rc;
}

#endif  // CETPICK_TEST_HOOKS
