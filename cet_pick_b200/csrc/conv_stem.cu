// Stem of the detector on tcgen05 tensor cores (sm_100a):
//   Conv2d(1, 16, 7, stride 2, pad 3, bias=False) + BatchNorm2d(eval) + ReLU
//   (cet_pick/models/networks/unet_small.py:35-37,72-74), fp32 (D,H,W) in -> bf16 NHWC16 out.
//
// As a CUDA-core kernel this layer is FFMA-issue bound (784 FMA per output pixel, 3.5 ms at
// 1024x1024x256); its HBM floor is 0.5 ms.  GEMM view used here (a "march" down the image like
// conv_march.cu): input rows are taken in PAIRS j = (2j, 2j+1).  For an output pixel (oy, ox) the 7x7
// window covers rows 2oy-3..2oy+3 = pairs oy-2..oy+1 and columns 2ox-3..2ox+3, so with
//     A_j[pixel ox, k = e*8 + c] = in[2j+e][2ox-4+c]            (K = 16 = one UMMA K step)
// pair j contributes to the FOUR output rows j-1..j+2, and ONE UMMA
//     D[128 px, 4 rows x 16 ch] += A_j[128 px, 16] * Wst[64, 16]^T
// feeds their accumulators, which sit in a ring of 16-column TMEM slots (one slot per output row).
// A_j is not a TMA-able view of the fp32 input (8-byte row stride between pixels, fp32 -> bf16), so
// converter warps build it: TMA stages the fp32 row pair in shared memory, each converter thread
// turns the 2 x 8 floats of its pixel into two 16-byte bf16 chunks and stores them in the
// SWIZZLE_32B K-major layout the UMMA descriptor expects (fence.proxy.async before the hand-off).
// Output row j-1 is complete when pair j has been consumed: the epilogue warps drain its slot
// (BN shift, ReLU, bf16, 32-byte NHWC store), zero it and hand it back.
//
// Roles (640 threads, 1 CTA/SM, persistent over strips of (z, 256-pixel x block, row range)):
//   warp 0 TMA producer | warp 1 UMMA issuer | warp 2 TMEM allocator | warps 4-11 converters |
//   warps 12-19 epilogue (warp % 4 = TMEM lane quadrant)
#include "conv_stem.cuh"
#include "common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <algorithm>

namespace cetpick {

namespace {

constexpr int STEM_THREADS = 640;
constexpr int MT = 2;                       // M-tiles (128 output pixels of one row) per step
constexpr int S = 16;                       // accumulator slots = output rows in flight per M-tile
constexpr int COUT = 16;
constexpr int TILE_COLS = S * COUT;         // TMEM columns of one M-tile's ring
constexpr int BOX_W = 136;                  // fp32 columns per TMA box: 64 output pixels read 2*63 + 8 = 134
constexpr int ROW_BYTES = BOX_W * 4;        // 544
constexpr int BOX_BYTES = 2 * ROW_BYTES;    // both rows of the pair
constexpr int BOX_STRIDE = 1152;            // 128-byte aligned slot of one box
constexpr int NBOX = MT * 2;
constexpr int RAW_STAGE = NBOX * BOX_STRIDE;
constexpr int RAW_STAGES = 6;
constexpr int A_TILE = 128 * 32;            // bf16 [128 px][16 k], 32-byte rows
constexpr int A_STAGE = MT * A_TILE;
constexpr int A_STAGES = 4;
constexpr int W_BYTES = 4 * COUT * 32;      // [4 slots][16 co][16 k] bf16
constexpr int SMEM_BYTES = 1024 + 1024 * ((W_BYTES + 1023) / 1024) + A_STAGES * A_STAGE + RAW_STAGES * RAW_STAGE;
// uint8 input (quantised tomogram, loader.py:16-25,117-120): a box row is 160 bytes starting 16 bytes left of the
// block's first pixel pair (a TMA box must start on a 16-byte boundary of the innermost dimension; 64 output pixels
// read bytes 12 .. 12 + 2*63 + 8); the converters map level -> bf16 through a per-lane copy of the 256-entry table
// (bank = lane: conflict-free), kept behind the raw stages.
constexpr int U8_ROW_BYTES = 160;
constexpr int U8_LEAD = 12;                 // bytes between the box start and column 2*ox - 4 of its first pixel
constexpr int U8_BOX_BYTES = 2 * U8_ROW_BYTES;
constexpr int U8_BOX_STRIDE = 384;
constexpr int LUT_BYTES = 256 * 32 * 4;
constexpr int SMEM_BYTES_U8 = SMEM_BYTES + LUT_BYTES;

struct alignas(64) StemParams {
  CUtensorMap tmIn;     // fp32 (W,H,D), box (BOX_W, 2, 1), zero fill  |  uint8 (W,H,D), box (144, 2, 1), zero fill
  CUtensorMap tmW;      // bf16 (16, 64), SWIZZLE_32B
  int D, H, W, h, w;
  int R, nchunk, nxb;   // rows per strip, strips along y, 256-pixel blocks along x
  long long total_strips;
  float bias_c[16];
  __nv_bfloat16* out;
  uint16_t lut[256];    // uint8 input: bf16 bits of level k (lut[0] == 0: the conv's zero padding is level 0)
};

struct Strip { int ma, mb, x0, z, j_lo, j_hi; };

__device__ __forceinline__ void decode_strip(const StemParams& p, long long k, Strip& s) {
  const int ch = (int)(k % p.nchunk);
  k /= p.nchunk;
  s.ma = ch * p.R;
  s.mb = min(s.ma + p.R, p.h);
  s.x0 = (int)(k % p.nxb) * (128 * MT);
  s.z = (int)(k / p.nxb);
  s.j_lo = max(s.ma - 2, 0);                 // pairs above the image are all padding
  s.j_hi = min(s.mb, (p.H - 1) >> 1);        // pairs below it too
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <bool U8>
__global__ void __launch_bounds__(STEM_THREADS, 1) stem_tc_kernel(const __grid_constant__ StemParams p) {
  constexpr int BSTRIDE = U8 ? U8_BOX_STRIDE : BOX_STRIDE;
  constexpr int BBYTES = U8 ? U8_BOX_BYTES : BOX_BYTES;
  constexpr int RSTAGE = NBOX * BSTRIDE;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t raw_full[RAW_STAGES], raw_empty[RAW_STAGES];
  __shared__ __align__(8) uint64_t a_full[A_STAGES], a_empty[A_STAGES];
  __shared__ __align__(8) uint64_t slot_full[S], slot_empty[S], bar_w;
  __shared__ uint32_t s_tmem_base;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sA = smem + 1024 * ((W_BYTES + 1023) / 1024);
  uint8_t* sRaw = sA + A_STAGES * A_STAGE;
  uint32_t* sLut = reinterpret_cast<uint32_t*>(sRaw + RAW_STAGES * RAW_STAGE);   // [256 levels][32 lanes]
  if (U8)
    for (int i = threadIdx.x; i < 256 * 32; i += STEM_THREADS) sLut[i] = p.lut[i >> 5];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tmIn);
    ptx::prefetch_tensormap(&p.tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < RAW_STAGES; ++s) { ptx::mbar_init(&raw_full[s], 1); ptx::mbar_init(&raw_empty[s], 8); }
    for (int s = 0; s < A_STAGES; ++s) { ptx::mbar_init(&a_full[s], 8); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < S; ++s) { ptx::mbar_init(&slot_full[s], 1); ptx::mbar_init(&slot_empty[s], 4 * MT); }
    ptx::mbar_init(&bar_w, 1);
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

  if (warp >= 12) {   // every accumulator slot starts at zero: all UMMAs accumulate
    const uint32_t row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp - 12) >> 2) * TILE_COLS);
    for (int c = 0; c < TILE_COLS; c += 16) ptx::tmem_st16_fill(row + c, 0u);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  constexpr uint32_t SMASK = S - 1, SSHIFT = 4;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bar_w, (uint32_t)W_BYTES);
      ptx::tma_load_2d(sW, &p.tmW, &bar_w, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (long long k = blockIdx.x; k < p.total_strips; k += gridDim.x) {
        Strip s;
        decode_strip(p, k, s);
        for (int j = s.j_lo; j <= s.j_hi; ++j) {
          ptx::mbar_wait(&raw_empty[stage], phase ^ 1u);
          ptx::mbar_arrive_expect_tx(&raw_full[stage], (uint32_t)(NBOX * BBYTES));
          uint8_t* dst = sRaw + (size_t)stage * RSTAGE;
#pragma unroll
          for (int b = 0; b < NBOX; ++b)
            ptx::tma_load_3d(dst + b * BSTRIDE, &p.tmIn, &raw_full[stage], 2 * (s.x0 + b * 64) - (U8 ? 16 : 4), 2 * j, s.z);
          if (++stage == RAW_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ UMMA issuer =================================
    constexpr uint32_t D_HI = ptx::smem_desc_hi(256, 6);     // SWIZZLE_32B: 8-row groups 256 bytes apart
    constexpr uint32_t IDESC0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint32_t sW_lo = ptx::smem_desc_lo(ptx::smem_u32(sW)), sA_lo = ptx::smem_desc_lo(ptx::smem_u32(sA));
    int astage = 0;
    uint32_t aphase = 0;
    uint32_t q0 = 0, q_touched = 0;
    ptx::mbar_wait(&bar_w, 0);
    for (long long k = blockIdx.x; k < p.total_strips; k += gridDim.x) {
      Strip s;
      decode_strip(p, k, s);
      for (int j = s.j_lo; j <= s.j_hi; ++j) {
        const int r_lo = max(s.ma, j - 1), r_hi = min(s.mb - 1, j + 2);
        const uint32_t q_lo = q0 + (uint32_t)(r_lo - s.ma);
        const int n = r_hi - r_lo + 1;
        while (q_touched < q_lo + (uint32_t)n) {    // rows touched for the first time: slot must be drained
          ptx::mbar_wait(&slot_empty[q_touched & SMASK], ((q_touched >> SSHIFT) & 1u) ^ 1u);
          ++q_touched;
        }
        const uint32_t s_lo = q_lo & SMASK;
        const int n1 = min(n, S - (int)s_lo), n2 = n - n1;       // ring wrap splits the column range
        const uint32_t id1 = IDESC0 | ((uint32_t)(n1 * COUT >> 3) << 17), id2 = IDESC0 | ((uint32_t)(n2 * COUT >> 3) << 17);
        const uint32_t boff1 = (uint32_t)(r_lo - (j - 1)) * ((COUT * 32) >> 4);
        const uint32_t boff2 = boff1 + (uint32_t)n1 * ((COUT * 32) >> 4);
        ptx::mbar_wait(&a_full[astage], aphase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
#pragma unroll
          for (int t = 0; t < MT; ++t) {
            const uint32_t a_lo = sA_lo + (uint32_t)((astage * A_STAGE + t * A_TILE) >> 4);
            ptx::umma_bf16_lohi(tmem_base + s_lo * COUT + t * TILE_COLS, a_lo, D_HI, sW_lo + boff1, D_HI, id1);
            if (n2) ptx::umma_bf16_lohi(tmem_base + t * TILE_COLS, a_lo, D_HI, sW_lo + boff2, D_HI, id2);
          }
          ptx::umma_commit(&a_empty[astage]);
          // row j-1 has now seen its four pairs; at the bottom of the image the last row never gets pair j+1
          if (j - 1 >= s.ma) ptx::umma_commit(&slot_full[(q0 + (uint32_t)(j - 1 - s.ma)) & SMASK]);
          if (j == s.j_hi)
            for (int r = max(s.ma, j); r < s.mb; ++r) ptx::umma_commit(&slot_full[(q0 + (uint32_t)(r - s.ma)) & SMASK]);
        }
        __syncwarp();
        if (++astage == A_STAGES) { astage = 0; aphase ^= 1u; }
      }
      q0 += (uint32_t)(s.mb - s.ma);
    }
  } else if (warp >= 4 && warp < 12) {
    // ================================ converters ==================================
    const int c = threadIdx.x - 128;
    const int t = c >> 7, px = c & 127, hb = px >> 6, pp = px & 63;
    const uint32_t raw_off = U8 ? (uint32_t)((t * 2 + hb) * BSTRIDE + ((pp * 2 + U8_LEAD) & ~3)) : (uint32_t)((t * 2 + hb) * BSTRIDE + pp * 8);
    const uint32_t u8_shift = (uint32_t)((pp * 2) & 3) * 8u;     // the pixel's 8 bytes start 0 or 2 bytes into an aligned word
    const uint32_t* lut = sLut + lane;
    const uint32_t a_row = (uint32_t)(px * 32);
    // SWIZZLE_32B: 16-byte chunk index (address bit 4) ^= address bit 7; tiles are 1024-byte aligned
    const uint32_t a_off0 = (uint32_t)(t * A_TILE) + (a_row ^ (((a_row >> 7) & 1u) << 4));
    const uint32_t a_off1 = (uint32_t)(t * A_TILE) + ((a_row + 16u) ^ (((a_row >> 7) & 1u) << 4));
    int stage = 0, astage = 0;
    uint32_t phase = 0, aphase = 0;
    for (long long k = blockIdx.x; k < p.total_strips; k += gridDim.x) {
      Strip s;
      decode_strip(p, k, s);
      for (int j = s.j_lo; j <= s.j_hi; ++j) {
        ptx::mbar_wait(&raw_full[stage], phase);
        const uint8_t* src = sRaw + (size_t)stage * RSTAGE + raw_off;
        uint4 c0, c1;
        if (U8) {
          uint32_t w[2][3];
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int i = 0; i < 3; ++i) w[e][i] = *reinterpret_cast<const uint32_t*>(src + e * U8_ROW_BYTES + i * 4);
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&raw_empty[stage]);
          uint32_t o[2][4];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint32_t lo = __funnelshift_r(w[e][0], w[e][1], u8_shift), hi = __funnelshift_r(w[e][1], w[e][2], u8_shift);
            o[e][0] = lut[(lo & 0xffu) << 5] | (lut[((lo >> 8) & 0xffu) << 5] << 16);
            o[e][1] = lut[((lo >> 16) & 0xffu) << 5] | (lut[(lo >> 24) << 5] << 16);
            o[e][2] = lut[(hi & 0xffu) << 5] | (lut[((hi >> 8) & 0xffu) << 5] << 16);
            o[e][3] = lut[((hi >> 16) & 0xffu) << 5] | (lut[(hi >> 24) << 5] << 16);
          }
          c0 = make_uint4(o[0][0], o[0][1], o[0][2], o[0][3]);
          c1 = make_uint4(o[1][0], o[1][1], o[1][2], o[1][3]);
        } else {
          float2 f[2][4];
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int i = 0; i < 4; ++i) f[e][i] = *reinterpret_cast<const float2*>(src + e * ROW_BYTES + i * 8);
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&raw_empty[stage]);
          c0.x = pack2(f[0][0].x, f[0][0].y); c0.y = pack2(f[0][1].x, f[0][1].y);
          c0.z = pack2(f[0][2].x, f[0][2].y); c0.w = pack2(f[0][3].x, f[0][3].y);
          c1.x = pack2(f[1][0].x, f[1][0].y); c1.y = pack2(f[1][1].x, f[1][1].y);
          c1.z = pack2(f[1][2].x, f[1][2].y); c1.w = pack2(f[1][3].x, f[1][3].y);
        }
        ptx::mbar_wait(&a_empty[astage], aphase ^ 1u);
        uint8_t* dst = sA + (size_t)astage * A_STAGE;
        *reinterpret_cast<uint4*>(dst + a_off0) = c0;
        *reinterpret_cast<uint4*>(dst + a_off1) = c1;
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_full[astage]);
        if (++stage == RAW_STAGES) { stage = 0; phase ^= 1u; }
        if (++astage == A_STAGES) { astage = 0; aphase ^= 1u; }
      }
    }
  } else if (warp >= 12) {
    // ================================ epilogue ====================================
    const int quad = warp & 3, t = (warp - 12) >> 2;
    const int m = quad * 32 + lane;
    uint32_t q0 = 0;
    for (long long k = blockIdx.x; k < p.total_strips; k += gridDim.x) {
      Strip s;
      decode_strip(p, k, s);
      const int x = s.x0 + t * 128 + m;
      const bool valid = x < p.w;
      for (int r = s.ma; r < s.mb; ++r) {
        const uint32_t q = q0 + (uint32_t)(r - s.ma);
        const uint32_t slot = q & SMASK, par = (q >> SSHIFT) & 1u;
        ptx::mbar_wait(&slot_full[slot], par);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t * TILE_COLS + (int)slot * COUT);
        uint32_t v[16];
        __syncwarp();
        ptx::tmem_ld16(taddr, v);
        ptx::tmem_ld_wait();
        ptx::tmem_st16_fill(taddr, 0u);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&slot_empty[slot]);
        if (valid) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            pk[i] = pack2(fmaxf(__uint_as_float(v[2 * i]) + p.bias_c[2 * i], 0.f),
                          fmaxf(__uint_as_float(v[2 * i + 1]) + p.bias_c[2 * i + 1], 0.f));
          uint4* dst = reinterpret_cast<uint4*>(p.out + (((size_t)s.z * p.h + r) * p.w + x) * COUT);
          ptx::st_global_256(dst, make_uint4(pk[0], pk[1], pk[2], pk[3]), make_uint4(pk[4], pk[5], pk[6], pk[7]));
        }
      }
      q0 += (uint32_t)(s.mb - s.ma);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool stem_tc_supported(const float* in, int W) {
  return (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
}

bool stem_tc_supported_u8(const uint8_t* in, int W) {
  return (W % 16 == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
}

std::vector<uint16_t> stem_pack_weights(const float* w, const double* scale) {
  std::vector<uint16_t> out((size_t)4 * COUT * 16, 0);
  for (int d = 0; d < 4; ++d)
    for (int co = 0; co < COUT; ++co)
      for (int e = 0; e < 2; ++e)
        for (int c = 0; c < 8; ++c) {
          const int ky = 2 * (1 - d) + e + 3, kx = c - 1;   // output row (j-1)+d sees input row 2j+e as tap ky
          float v = 0.f;
          if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = (float)((double)w[co * 49 + ky * 7 + kx] * (scale ? scale[co] : 1.0));
          out[((size_t)d * COUT + co) * 16 + e * 8 + c] = f2bf_host(v);
        }
  return out;
}

int conv_stem_launch(const StemLaunch& L, cudaStream_t stream) {
  const bool u8 = L.in_u8 != nullptr;
  if ((!L.in && !u8) || (L.in && u8) || !L.wpk || !L.out || L.D <= 0 || L.H <= 0 || L.W <= 0) return CETPICK_ERR_BAD_ARG;
  if (u8 ? !stem_tc_supported_u8(L.in_u8, L.W) : !stem_tc_supported(L.in, L.W)) return CETPICK_ERR_UNSUPPORTED;
  if (u8 && L.lut[0] != 0) return CETPICK_ERR_BAD_ARG;     // level 0 doubles as the zero padding
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    CETPICK_CUDA(cudaFuncSetAttribute(stem_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CETPICK_CUDA(cudaFuncSetAttribute(stem_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_U8));
  }
  StemParams p;
  memset(&p, 0, sizeof(p));
  if (u8) memcpy(p.lut, L.lut, sizeof(p.lut));
  p.D = L.D; p.H = L.H; p.W = L.W;
  p.h = (L.H - 1) / 2 + 1; p.w = (L.W - 1) / 2 + 1;
  p.out = static_cast<__nv_bfloat16*>(L.out);
  memcpy(p.bias_c, L.bias, sizeof(p.bias_c));
  p.nxb = ceil_div(p.w, 128 * MT);
  const long long base_strips = (long long)p.nxb * L.D;
  // strips along y: enough to balance the persistent grid, long enough that the 3 extra pairs stay cheap
  const int sms = num_sms();
  int best_n = 1;
  double best_eff = -1.0;
  for (int n = 1; n <= std::max(1, p.h / 8); ++n) {
    const int R = ceil_div(p.h, n);
    if (ceil_div(p.h, R) != n) continue;
    const long long strips = base_strips * n;
    const double waves = (double)ceil_div<long long>(strips, sms);
    const double eff = ((double)strips / (waves * sms)) * ((double)R / (R + 3));
    if (eff > best_eff + 1e-9) { best_eff = eff; best_n = n; }
  }
  p.nchunk = best_n;
  p.R = ceil_div(p.h, best_n);
  p.total_strips = base_strips * p.nchunk;
  int rc;
  if (u8) {
    const uint64_t dims[3] = {(uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.D};
    const uint64_t strides[2] = {(uint64_t)L.W, (uint64_t)L.W * L.H};
    const uint32_t box[3] = {(uint32_t)U8_ROW_BYTES, 2, 1};
    if ((rc = tmap_encode_u8_zerofill(&p.tmIn, L.in_u8, 3, dims, strides, box))) return rc;
  } else {
    const uint64_t dims[3] = {(uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.D};
    const uint64_t strides[2] = {(uint64_t)L.W * 4, (uint64_t)L.W * L.H * 4};
    const uint32_t box[3] = {(uint32_t)BOX_W, 2, 1};
    if ((rc = tmap_encode_f32_zerofill(&p.tmIn, L.in, 3, dims, strides, box))) return rc;
  }
  {
    const uint64_t dims[2] = {16, (uint64_t)4 * COUT};
    const uint64_t strides[1] = {32};
    const uint32_t box[2] = {16, (uint32_t)(4 * COUT)};
    if ((rc = tmap_encode_bf16(&p.tmW, L.wpk, 2, dims, strides, box, 16))) return rc;
  }
  const int grid = (int)std::min<long long>(p.total_strips, sms);
  if (u8) stem_tc_kernel<true><<<grid, STEM_THREADS, SMEM_BYTES_U8, stream>>>(p);
  else stem_tc_kernel<false><<<grid, STEM_THREADS, SMEM_BYTES, stream>>>(p);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

}  // namespace cetpick

using namespace cetpick;

#ifdef CETPICK_TEST_HOOKS   // test / tuning hooks: built into libcetpick_test_sm100a.so only (include/cetpick_test.h)
// Test hook: the stem from a PyTorch-layout fp32 host weight (16,1,7,7), BN scale/shift on the host
// (packs, uploads, launches, synchronises, frees) -- tests/test_gpu_conv.py.
extern "C" int cetpick_conv_stem_bf16(const float* in, int D, int H, int W, const float* w_host,
                                      const float* scale_host, const float* shift_host, void* out, void* stream) {
  g_launches = 0;
  if (!in || !w_host || !out) return CETPICK_ERR_BAD_ARG;
  if (!stem_tc_supported(in, W)) return CETPICK_ERR_UNSUPPORTED;
  double sc[16];
  for (int c = 0; c < 16; ++c) sc[c] = scale_host ? (double)scale_host[c] : 1.0;
  std::vector<uint16_t> pk = stem_pack_weights(w_host, sc);
  void* d = nullptr;
  CETPICK_CUDA(cudaMalloc(&d, pk.size() * 2));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = CETPICK_OK;
  if (cudaMemcpyAsync(d, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = CETPICK_ERR_CUDA;
  if (rc == CETPICK_OK) {
    StemLaunch L;
    L.in = in; L.D = D; L.H = H; L.W = W; L.wpk = d; L.out = out;
    for (int c = 0; c < 16; ++c) L.bias[c] = shift_host ? shift_host[c] : 0.f;
    rc = conv_stem_launch(L, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (rc == CETPICK_OK && e != cudaSuccess) return cuda_fail(e, "conv_stem");
  return rc;
}

#endif  // CETPICK_TEST_HOOKS
