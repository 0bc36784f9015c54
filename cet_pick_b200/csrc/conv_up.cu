// ConvTranspose2d(k=2, s=2, bias) + BN + ReLU of UpConv (cet_pick/models/networks/unet.py:148-161,
// 375-388) as ONE tcgen05 GEMM per level: every input pixel produces a 2x2 block of output pixels,
//     D[pixel, (dy,dx,co)] = sum_ci A[pixel, ci] * W[ci, co, dy, dx]            (N = 4*Cout)
// The layer is output-bandwidth bound (it writes 4x the pixels it reads, K = Cin is only 64-256), so
// the kernel is built around the epilogue: the weights of the CTA's column block (<= 256 columns)
// are loaded once and stay in shared memory, the activations stream through a TMA ring as flat
// 128-pixel tiles (no halo: the op is pointwise), the fp32 accumulator is double-buffered in TMEM, and
// EIGHT epilogue warps (two per TMEM lane quadrant, half the columns each) do bias + ReLU + bf16 and
// the pixel-shuffle store (with autocrop, unet.py:285-292) while the next tile's MMAs run.
//   warp 0: TMA producer   warp 1: UMMA issuer (warp-uniform loop)   warp 2: TMEM allocator
//   warps 4-11: epilogue
#include "conv_up.cuh"
#include "common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <algorithm>

namespace cetpick {

namespace {

constexpr int UP_THREADS = 384;
constexpr int MAX_STAGES = 8;
constexpr int KC = 64;                      // channels per chunk = 128-byte swizzle span
constexpr int A_STAGE = 128 * KC * 2;       // one 128-pixel x 64-channel chunk

struct alignas(64) UpParams {
  CUtensorMap tmA, tmB;
  int chunks, nsplit;
  long long P, tiles;                        // input pixels, 128-pixel tiles
  int h, w, Ho, Wo, Cout;
  int stages;
  const float* bias;                         // [4*Cout] fp32, (dy,dx,co) order
  __nv_bfloat16* out;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int NB>   // columns per CTA (128 or 256)
__global__ void __launch_bounds__(UP_THREADS, 1) conv_up_kernel(const __grid_constant__ UpParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2], bar_w;
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_bias[NB];

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                        // [chunk][NB][64] bf16
  uint8_t* sA = smem + (size_t)p.chunks * NB * KC * 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x % p.nsplit;
  const long long tile0 = blockIdx.x / p.nsplit, tstep = gridDim.x / p.nsplit;

  if (warp == 0 && lane == 0) { ptx::prefetch_tensormap(&p.tmA); ptx::prefetch_tensormap(&p.tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&bar_full[s], 1); ptx::mbar_init(&bar_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&bar_tfull[a], 1); ptx::mbar_init(&bar_tempty[a], 8); }
    ptx::mbar_init(&bar_w, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&s_tmem_base, 2 * NB);
    ptx::tmem_relinquish();
  }
  if (warp == 3)
    for (int c = lane; c < NB; c += 32) s_bias[c] = p.bias[split * NB + c];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bar_w, (uint32_t)(p.chunks * NB * KC * 2));
      for (int c = 0; c < p.chunks; ++c)
        ptx::tma_load_2d(sW + (size_t)c * NB * KC * 2, &p.tmB, &bar_w, 0, (split * p.chunks + c) * NB);
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = tile0; t < p.tiles; t += tstep)
        for (int c = 0; c < p.chunks; ++c) {
          ptx::mbar_wait(&bar_empty[stage], phase ^ 1u);
          ptx::mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)A_STAGE);
          ptx::tma_load_2d(sA + (size_t)stage * A_STAGE, &p.tmA, &bar_full[stage], c * KC, (int)(t * 128));
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
    }
  } else if (warp == 1) {
    constexpr uint32_t HI = ptx::smem_desc_hi(8 * KC * 2, 2);     // 128-byte swizzle, dense 8-row groups
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sW_lo = ptx::smem_desc_lo(ptx::smem_u32(sW)), sA_lo = ptx::smem_desc_lo(ptx::smem_u32(sA));
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    ptx::mbar_wait(&bar_w, 0);
    for (long long t = tile0; t < p.tiles; t += tstep) {
      ptx::mbar_wait(&bar_tempty[acc], acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * NB);
      for (int c = 0; c < p.chunks; ++c) {
        ptx::mbar_wait(&bar_full[stage], phase);
        ptx::tc_fence_after();
        const uint32_t a_lo = sA_lo + (uint32_t)(stage * (A_STAGE >> 4));
        const uint32_t w_lo = sW_lo + (uint32_t)(c * ((NB * KC * 2) >> 4));
        if (ptx::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk) {
            if (c == 0 && kk == 0)
              ptx::umma_bf16(d, ((uint64_t)HI << 32) | a_lo, ((uint64_t)HI << 32) | w_lo, IDESC, 0u);
            else
              ptx::umma_bf16_lohi(d, a_lo + kk * 2, HI, w_lo + kk * 2, HI, IDESC);
          }
          ptx::umma_commit(&bar_empty[stage]);
          if (c == p.chunks - 1) ptx::umma_commit(&bar_tfull[acc]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3, half = (warp - 4) >> 2;
    const int m = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long t = tile0; t < p.tiles; t += tstep) {
      const long long pix = t * 128 + m;
      const bool valid = pix < p.P;
      const long long plane = (long long)p.h * p.w;
      const int img = (int)(pix / plane);
      const int rem = (int)(pix - (long long)img * plane);
      const int y = rem / p.w, x = rem - y * p.w;
      ptx::mbar_wait(&bar_tfull[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * NB + half * (NB / 2));
#pragma unroll 2
      for (int c0 = 0; c0 < NB / 2; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        ptx::tmem_ld16(t_row + c0, v);
        ptx::tmem_ld16(t_row + c0 + 16, v + 16);
        ptx::tmem_ld_wait();
        if (c0 + 32 >= NB / 2) {                 // last read of this accumulator: hand it back early
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bar_tempty[acc]);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int lcol = half * (NB / 2) + c0 + g * 16;         // column inside the CTA's block
          const int col = split * NB + lcol;                      // (dy,dx,co) column
          const int qd = col / p.Cout, ch = col - qd * p.Cout;
          const int oy = 2 * y + (qd >> 1), ox = 2 * x + (qd & 1);
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = fmaxf(__uint_as_float(v[g * 16 + i]) + s_bias[lcol + i], 0.f);
          if (valid && oy < p.Ho && ox < p.Wo) {                  // autocrop (unet.py:285-292)
            uint4* dst = reinterpret_cast<uint4*>(p.out + (((size_t)img * p.Ho + oy) * p.Wo + ox) * p.Cout + ch);
            uint4 w0, w1;
            w0.x = pack_bf16x2(f[0], f[1]);   w0.y = pack_bf16x2(f[2], f[3]);
            w0.z = pack_bf16x2(f[4], f[5]);   w0.w = pack_bf16x2(f[6], f[7]);
            w1.x = pack_bf16x2(f[8], f[9]);   w1.y = pack_bf16x2(f[10], f[11]);
            w1.z = pack_bf16x2(f[12], f[13]); w1.w = pack_bf16x2(f[14], f[15]);
            ptx::st_global_256(dst, w0, w1);
          }
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 2 * NB);
}

template <int NB>
int launch_up(UpParams& p, cudaStream_t stream) {
  auto kern = conv_up_kernel<NB>;
  static DeviceOnce attr_once;
  static int static_smem = 0;
  if (attr_once.first()) {
    cudaFuncAttributes fa;
    CETPICK_CUDA(cudaFuncGetAttributes(&fa, kern));
    CETPICK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes));
    static_smem = (int)fa.sharedSizeBytes;
  }
  const size_t wbytes = (size_t)p.chunks * NB * KC * 2;
  const size_t avail = (size_t)227 * 1024 - static_smem - 1024 - wbytes;
  p.stages = (int)std::min<size_t>(MAX_STAGES, avail / A_STAGE);
  if (p.stages < 2) return CETPICK_ERR_UNSUPPORTED;
  const size_t smem = 1024 + wbytes + (size_t)p.stages * A_STAGE;
  const int sms = num_sms();
  long long grid = std::min<long long>(p.tiles * p.nsplit, sms);
  grid = std::max<long long>(p.nsplit, grid / p.nsplit * p.nsplit);
  kern<<<(int)grid, UP_THREADS, smem, stream>>>(p);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

}  // namespace

bool upconv_supported(int Cin, int Cout) {
  if (Cin % KC || Cout % 16 || Cout < 32) return false;
  const int N = 4 * Cout, NB = std::min(N, 256);
  if (N % NB) return false;
  return (size_t)(Cin / KC) * NB * KC * 2 + 3 * A_STAGE + 4096 <= (size_t)227 * 1024;
}

std::vector<uint16_t> upconv_pack_weights(const float* w, int Cin, int Cout, const double* scale) {
  const int N = 4 * Cout, NB = std::min(N, 256), nsplit = N / NB, chunks = Cin / KC;
  std::vector<uint16_t> out((size_t)N * Cin);
  for (int sp = 0; sp < nsplit; ++sp)
    for (int c = 0; c < chunks; ++c)
      for (int j = 0; j < NB; ++j)
        for (int k = 0; k < KC; ++k) {
          const int col = sp * NB + j, qd = col / Cout, co = col % Cout, ci = c * KC + k;
          const double v = (double)w[((size_t)ci * Cout + co) * 4 + qd] * (scale ? scale[co] : 1.0);
          out[(((size_t)sp * chunks + c) * NB + j) * KC + k] = f2bf_host((float)v);
        }
  return out;
}

int conv_up_launch(const UpLaunch& L, cudaStream_t stream) {
  if (!upconv_supported(L.Cin, L.Cout)) return CETPICK_ERR_UNSUPPORTED;
  if (!L.src || !L.wpk || !L.bias || !L.out || L.NIMG <= 0 || L.h <= 0 || L.w <= 0 || L.Ho <= 0 || L.Wo <= 0)
    return CETPICK_ERR_BAD_ARG;
  UpParams p;
  memset(&p, 0, sizeof(p));
  const int N = 4 * L.Cout, NB = std::min(N, 256);
  p.chunks = L.Cin / KC;
  p.nsplit = N / NB;
  p.P = (long long)L.NIMG * L.h * L.w;
  p.tiles = ceil_div<long long>(p.P, 128);
  p.h = L.h; p.w = L.w; p.Ho = L.Ho; p.Wo = L.Wo; p.Cout = L.Cout;
  p.bias = L.bias; p.out = static_cast<__nv_bfloat16*>(L.out);
  int rc;
  {
    const uint64_t dims[2] = {(uint64_t)L.Cin, (uint64_t)p.P};
    const uint64_t strides[1] = {(uint64_t)L.Cin * 2};
    const uint32_t box[2] = {KC, 128};
    if ((rc = tmap_encode_bf16(&p.tmA, L.src, 2, dims, strides, box, KC))) return rc;
  }
  {
    const uint64_t dims[2] = {KC, (uint64_t)p.nsplit * p.chunks * NB};
    const uint64_t strides[1] = {KC * 2};
    const uint32_t box[2] = {KC, (uint32_t)NB};
    if ((rc = tmap_encode_bf16(&p.tmB, L.wpk, 2, dims, strides, box, KC))) return rc;
  }
  return NB == 128 ? launch_up<128>(p, stream) : launch_up<256>(p, stream);
}

}  // namespace cetpick

using namespace cetpick;

#ifdef CETPICK_TEST_HOOKS   // test / tuning hooks: built into libcetpick_test_sm100a.so only (include/cetpick_test.h)
// Test hook: ConvTranspose2d(k2,s2)+bias+ReLU from a PyTorch-layout fp32 HOST weight (Cin,Cout,2,2)
// (packs, uploads, launches, synchronises) -- tests/test_gpu_conv.py.
extern "C" int cetpick_upconv_bf16(const void* src, int Cin, int NIMG, int h, int w, const float* w_host,
                                   const float* bias_host, int Cout, void* out, int Ho, int Wo, void* stream) {
  g_launches = 0;
  if (!w_host || !bias_host) return CETPICK_ERR_BAD_ARG;
  if (!upconv_supported(Cin, Cout)) return CETPICK_ERR_UNSUPPORTED;
  std::vector<uint16_t> pk = upconv_pack_weights(w_host, Cin, Cout, nullptr);
  std::vector<float> b4((size_t)4 * Cout);
  for (int q = 0; q < 4; ++q)
    for (int c = 0; c < Cout; ++c) b4[(size_t)q * Cout + c] = bias_host[c];
  void *dw = nullptr, *db = nullptr;
  CETPICK_CUDA(cudaMalloc(&dw, pk.size() * 2));
  if (cudaMalloc(&db, b4.size() * 4) != cudaSuccess) { cudaFree(dw); return CETPICK_ERR_CUDA; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = CETPICK_OK;
  if (cudaMemcpyAsync(dw, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(db, b4.data(), b4.size() * 4, cudaMemcpyHostToDevice, st) != cudaSuccess)
    rc = CETPICK_ERR_CUDA;
  if (rc == CETPICK_OK) {
    UpLaunch L;
    L.src = src; L.Cin = Cin; L.NIMG = NIMG; L.h = h; L.w = w; L.wpk = dw; L.bias = static_cast<const float*>(db);
    L.Cout = Cout; L.out = out; L.Ho = Ho; L.Wo = Wo;
    rc = conv_up_launch(L, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dw);
  cudaFree(db);
  if (rc == CETPICK_OK && e != cudaSuccess) return cuda_fail(e, "conv_up");
  return rc;
}

#endif  // CETPICK_TEST_HOOKS
