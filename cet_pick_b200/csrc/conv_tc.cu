// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), used for every conv of the detector
// except the 1-channel stem and the 1-channel `hm` head.
//
// GEMM view: M = 128 output pixels (a 16 x 8 tile of one image / z-plane), N = output channels,
// K = taps x input channels.  Nothing is im2col'ed: for every (tap, channel chunk) the producer
// issues ONE TMA box load of the activation tensor shifted by the tap offset; TMA's out-of-bounds
// zero fill is the convolution's zero padding (in x, y and, for the 3-D head, z).  The weights of
// the same k-block arrive by a second TMA load.  Operands are bf16, K-major, hardware-swizzled;
// the fp32 accumulator lives in TMEM and is double-buffered so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Persistent grid, one CTA per SM, warp-specialised:
//   warp 0: TMA producer   warp 1: MMA issuer (one thread)   warp 2: TMEM allocator
//   warps 4-7: epilogue (TMEM -> registers -> bias/ReLU/bf16 -> global)
#include "conv_tc.cuh"
#include "common.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <algorithm>
#include <mutex>
#include <unordered_map>

namespace cetpick {

namespace {

constexpr int TILE_W = 16, TILE_H = 8, TILE_M = TILE_W * TILE_H;  // 128 pixels = UMMA M
constexpr int CONV_THREADS = 256;
constexpr int MAX_STAGES = 8;
constexpr int TMEM_COLS = 512;

struct alignas(64) ConvParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  int nsrc, chunks0, chunks1;
  int ntaps;
  int KC, NKB, nkb;            // channels per k-block, k-blocks per stage, k-blocks in total
  int Ntot, NB, n_nb;          // GEMM N, columns per tile, tiles along N
  int NIMG, H, W, tiles_x, tiles_y;
  long long total_tiles;
  int stages, a_sub, b_sub;    // pipeline depth, bytes of one A / B k-block
  int layout_type;             // UMMA swizzle code
  int epi, relu;
  int tf32;                    // 1: fp32 activations / weights (pre-rounded to TF32), kind::tf32 UMMA, fp32 NHWC out
  const float* bias;
  void* out;
  int out_cstride, Ho, Wo, Cout;
  signed char tdz[27], tdy[27], tdx[27];
};

// fp32 -> nearest TF32 (10-bit mantissa), ties away from zero: stored activations are exactly what the UMMA reads
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 x TF32 -> FP32 (K = 8 per instruction) or BF16 (K = 16):
// (lo, hi) descriptor words, accumulate flag as an operand; kind::tf32 or kind::f16 (warp-uniform choice)
__device__ __forceinline__ void umma_lohi(bool tf32, uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
  if (tf32) {
    asm volatile(
        "{\n\t"
        ".reg .b64 da, db;\n\t"
        ".reg .pred p;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    ptx::umma_bf16_lohi_acc(tmem_d, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
  }
}

__device__ __forceinline__ void decode_tile(const ConvParams& p, long long t, int& img, int& y0,
                                            int& x0, int& nb) {
  nb = (int)(t % p.n_nb);
  t /= p.n_nb;
  x0 = (int)(t % p.tiles_x) * TILE_W;
  t /= p.tiles_x;
  y0 = (int)(t % p.tiles_y) * TILE_H;
  img = (int)(t / p.tiles_y);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(CONV_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t s_tmem_base;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int a_stage = p.a_sub * p.NKB, b_stage = p.b_sub * p.NKB;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.stages * a_stage;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tmA[0]);
    if (p.nsrc > 1) ptx::prefetch_tensormap(&p.tmA[1]);
    ptx::prefetch_tensormap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&bar_full[s], 1); ptx::mbar_init(&bar_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { ptx::mbar_init(&bar_tfull[a], 1); ptx::mbar_init(&bar_tempty[a], 4); }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&s_tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // Producer and issuer run their loops on the WHOLE warp with warp-uniform values and elect one lane for the
    // asynchronous instructions: issued from inside an `if (lane == 0)` region, every UTMALDG / UTCHMMA is wrapped in a
    // lane-serialising loop (the operands must be in uniform registers) and the issuer's scalar chain bounds the kernel.
    int stage = 0;
    uint32_t phase = 0;
    const int kb_src0 = p.ntaps * p.chunks0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int img, y0, x0, nb;
      decode_tile(p, t, img, y0, x0, nb);
      for (int kb = 0; kb < p.nkb; kb += p.NKB) {
        const int nvalid = min(p.NKB, p.nkb - kb);
        ptx::mbar_wait(&bar_empty[stage], phase ^ 1u);
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)nvalid * (uint32_t)(p.a_sub + p.b_sub));
          for (int j = 0; j < nvalid; ++j) {
            const int k = kb + j;
            int src, tap, chunk;
            if (k < kb_src0) { src = 0; tap = k / p.chunks0; chunk = k - tap * p.chunks0; }
            else { const int k1 = k - kb_src0; src = 1; tap = k1 / p.chunks1; chunk = k1 - tap * p.chunks1; }
            ptx::tma_load_4d(sA + (size_t)stage * a_stage + (size_t)j * p.a_sub, &p.tmA[src],
                             &bar_full[stage], chunk * p.KC, x0 + p.tdx[tap], y0 + p.tdy[tap],
                             img + p.tdz[tap]);
            ptx::tma_load_2d(sB + (size_t)stage * b_stage + (size_t)j * p.b_sub, &p.tmB,
                             &bar_full[stage], 0, k * p.Ntot + nb * p.NB);
          }
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    // kind::tf32: a_format = b_format = 2 (TF32) in bits [7,10) / [10,13), c_format = 1 (F32)
    const bool tf32 = p.tf32 != 0;
    const uint32_t idesc = tf32 ? ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NB >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24))
                                : ptx::make_idesc_bf16(TILE_M, p.NB);
    const uint32_t esz = tf32 ? 4u : 2u;
    const uint32_t sbo = 8u * esz * (uint32_t)p.KC;     // 8 rows x (KC * element bytes)
    const int k16s = (int)(p.KC * esz / 32u);           // one UMMA consumes 32 bytes of K per row
    const uint32_t D_HI = ptx::smem_desc_hi(sbo, (uint32_t)p.layout_type);
    const uint32_t sA_lo = ptx::smem_desc_lo(ptx::smem_u32(sA)), sB_lo = ptx::smem_desc_lo(ptx::smem_u32(sB));
    const uint32_t a_stage16 = (uint32_t)a_stage >> 4, b_stage16 = (uint32_t)b_stage >> 4;
    const uint32_t a_sub16 = (uint32_t)p.a_sub >> 4, b_sub16 = (uint32_t)p.b_sub >> 4;
    const uint32_t a_full = ptx::smem_u32(&bar_full[0]), a_empty = ptx::smem_u32(&bar_empty[0]);
    const uint32_t a_tfull = ptx::smem_u32(&bar_tfull[0]), a_tempty = ptx::smem_u32(&bar_tempty[0]);
    uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      ptx::mbar_wait_a(a_tempty + 8u * acc, acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.NB;
      uint32_t accumulate = 0;
      for (int kb = 0; kb < p.nkb; kb += p.NKB) {
        const int nvalid = min(p.NKB, p.nkb - kb);
        ptx::mbar_wait_a(a_full + 8u * stage, phase);
        ptx::tc_fence_after();
        const uint32_t a_lo = sA_lo + stage * a_stage16, b_lo = sB_lo + stage * b_stage16;
        if (ptx::elect_one()) {
          for (int j = 0; j < nvalid; ++j)
            for (int k = 0; k < k16s; ++k) {
              umma_lohi(tf32, d_tmem, a_lo + (uint32_t)j * a_sub16 + 2u * k, D_HI, b_lo + (uint32_t)j * b_sub16 + 2u * k, D_HI,
                        idesc, accumulate);
              accumulate = 1;
            }
          ptx::umma_commit_a(a_empty + 8u * stage);   // frees the smem slot when these MMAs retire
        }
        __syncwarp();
        accumulate = 1;
        if (++stage == (uint32_t)p.stages) { stage = 0; phase ^= 1u; }
      }
      if (ptx::elect_one()) ptx::umma_commit_a(a_tfull + 8u * acc);       // accumulator complete -> epilogue
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int m = q * 32 + lane;            // row of the tile = pixel
    const int px = m % TILE_W, py = m / TILE_W;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long long t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int img, y0, x0, nb;
      decode_tile(p, t, img, y0, x0, nb);
      const int x = x0 + px, y = y0 + py;
      const bool valid = (x < p.W) && (y < p.H);
      ptx::mbar_wait(&bar_tfull[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.NB);
      float inv_norm = 1.0f;
      if (p.epi == EPI_F32_L2NORM_NCDHW) {
        float ss = 0.f;
        for (int c0 = 0; c0 < p.NB; c0 += 16) {
          uint32_t v[16];
          __syncwarp();
          ptx::tmem_ld16(t_row + c0, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float f = __uint_as_float(v[i]); ss = fmaf(f, f, ss); }
        }
        inv_norm = 1.0f / fmaxf(sqrtf(ss), 1e-12f);   // F.normalize(dim=1), eps = 1e-12
      }
      for (int c0 = 0; c0 < p.NB; c0 += 16) {
        uint32_t v[16];
        __syncwarp();                       // tcgen05.ld is .sync.aligned: reconverge first
        ptx::tmem_ld16(t_row + c0, v);
        ptx::tmem_ld_wait();
        const int col = nb * p.NB + c0;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col + i));
            f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
        }
        if (!valid) {
          // masked pixel of a ragged edge tile: nothing to store
        } else if (p.epi == EPI_BF16_NHWC || p.epi == EPI_UPCONV_2X2) {
          size_t off;
          bool store = true;
          if (p.epi == EPI_BF16_NHWC) {
            off = (((size_t)img * p.H + y) * p.W + x) * p.out_cstride + col;
          } else {
            const int qd = col / p.Cout, ch = col - qd * p.Cout;
            const int oy = 2 * y + (qd >> 1), ox = 2 * x + (qd & 1);
            store = (oy < p.Ho) && (ox < p.Wo);              // autocrop (unet.py:285-292)
            off = (((size_t)img * p.Ho + oy) * p.Wo + ox) * p.Cout + ch;
          }
          if (p.tf32) {                                      // fp32 NHWC activations, rounded to TF32 once here
            if (store) {
              float* d32 = reinterpret_cast<float*>(p.out) + off;
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(d32 + i) = make_float4(round_tf32(f[i]), round_tf32(f[i + 1]), round_tf32(f[i + 2]), round_tf32(f[i + 3]));
            }
            continue;
          }
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
          uint4 w0, w1;
          w0.x = pack_bf16x2(f[0], f[1]);   w0.y = pack_bf16x2(f[2], f[3]);
          w0.z = pack_bf16x2(f[4], f[5]);   w0.w = pack_bf16x2(f[6], f[7]);
          w1.x = pack_bf16x2(f[8], f[9]);   w1.y = pack_bf16x2(f[10], f[11]);
          w1.z = pack_bf16x2(f[12], f[13]); w1.w = pack_bf16x2(f[14], f[15]);
          if (store) {
            if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) ptx::st_global_256(dst, w0, w1);   // one full sector
            else { reinterpret_cast<uint4*>(dst)[0] = w0; reinterpret_cast<uint4*>(dst)[1] = w1; }
          }
        } else if (p.epi == EPI_F32_ROWMAJOR) {
          float* dst = reinterpret_cast<float*>(p.out) +
                       (((size_t)img * p.H + y) * p.W + x) * p.Ntot + col;
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(dst + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
        } else {  // EPI_F32_L2NORM_NCDHW
          const size_t plane = (size_t)p.H * p.W, vol = plane * p.NIMG;
          float* dst = reinterpret_cast<float*>(p.out) + (size_t)col * vol + (size_t)img * plane +
                       (size_t)y * p.W + x;
#pragma unroll
          for (int i = 0; i < 16; ++i) dst[(size_t)i * vol] = f[i] * inv_norm;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&bar_tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

CUtensorMapSwizzle swizzle_of(int KC) {
  return KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Tensor-map cache.  A CUtensorMap is a pure function of (base pointer, data type, rank, dims, strides, box, swizzle,
// L2 promotion, out-of-bounds fill); a forward of the same plan on the same buffers and shape asks for the same ~45
// descriptors every time (the detector loop: same workspace, same input staging buffer), so they are encoded once
// (cuTensorMapEncodeTiled is a 1-2 us driver call each) and copied from here afterwards.
// ---------------------------------------------------------------------------------------------
namespace {
struct TmapKey {
  const void* base;
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  int dtype, rank, swizzle, promo, oob, dev;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static_assert(sizeof(TmapKey) % 8 == 0, "hashed as 64-bit words");
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
int64_t g_tmap_hits = 0, g_tmap_misses = 0;
constexpr size_t TMAP_CACHE_MAX = 4096;

int tmap_encode_any(CUtensorMap* tm, CUtensorMapDataType dtype, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz, CUtensorMapL2promotion promo,
                    CUtensorMapFloatOOBfill oob, const char* what) {
  if (rank < 1 || rank > 5) return CETPICK_ERR_BAD_ARG;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.dtype = (int)dtype; key.rank = rank; key.swizzle = (int)swz; key.promo = (int)promo; key.oob = (int)oob;
  key.dev = current_device();
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> lock(g_tmap_mu);
    auto it = g_tmaps.find(key);
    if (it != g_tmaps.end()) { *tm = it->second; ++g_tmap_hits; return CETPICK_OK; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) { g_cuda_err = "cuTensorMapEncodeTiled not available"; return CETPICK_ERR_CUDA; }
  cuuint64_t d[5], st[4];
  cuuint32_t b[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  CUresult r = enc(tm, dtype, (cuuint32_t)rank, const_cast<void*>(base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, oob);
  if (r != CUDA_SUCCESS) {
    g_cuda_err = std::string("cuTensorMapEncodeTiled(") + what + ") failed: " + std::to_string((int)r);
    return CETPICK_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lock(g_tmap_mu);
  if (g_tmaps.size() >= TMAP_CACHE_MAX) g_tmaps.clear();
  g_tmaps.emplace(key, *tm);
  ++g_tmap_misses;
  return CETPICK_OK;
}
}  // namespace

void tmap_cache_stats(int64_t* hits, int64_t* misses) {
  std::lock_guard<std::mutex> lock(g_tmap_mu);
  if (hits) *hits = g_tmap_hits;
  if (misses) *misses = g_tmap_misses;
}

// bf16 tiled tensor map whose innermost box dimension is KC channels (= the swizzle span)
int tmap_encode_bf16(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int KC) {
  return tmap_encode_any(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle_of(KC),
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE, "bf16");
}

// fp32 tiled tensor map, no swizzle, out-of-bounds elements read as NaN (decode: fmaxf ignores them)
int tmap_encode_f32_nanfill(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box) {
  return tmap_encode_any(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA, "f32 nan-fill");
}

// fp32 tiled tensor map, no swizzle, out-of-bounds elements read as zero (stem: the conv's zero padding)
int tmap_encode_f32_zerofill(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box) {
  return tmap_encode_any(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE, "f32");
}

// uint8 tiled tensor map, no swizzle, out-of-bounds elements read as zero (quantised-input stem)
int tmap_encode_u8_zerofill(CUtensorMap* tm, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box) {
  return tmap_encode_any(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE, "u8");
}

int conv_tc_launch(const ConvLaunch& L, cudaStream_t stream) {
  if (L.KC != 16 && L.KC != 32 && L.KC != 64) return CETPICK_ERR_BAD_ARG;
  if (L.Ntot <= 0 || (L.Ntot % 16) || L.ntaps < 1 || L.ntaps > 27 || L.nsrc < 1 || L.nsrc > 2) return CETPICK_ERR_BAD_ARG;
  for (int s = 0; s < L.nsrc; ++s)
    if (!L.src[s] || L.C[s] <= 0 || (L.C[s] % L.KC)) return CETPICK_ERR_BAD_ARG;
  EncodeTiledFn enc = get_encode();
  if (!enc) { g_cuda_err = "cuTensorMapEncodeTiled not available"; return CETPICK_ERR_CUDA; }

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.nsrc = L.nsrc;
  p.chunks0 = L.C[0] / L.KC;
  p.chunks1 = L.nsrc > 1 ? L.C[1] / L.KC : 1;
  p.ntaps = L.ntaps;
  p.KC = L.KC;
  p.nkb = L.ntaps * (p.chunks0 + (L.nsrc > 1 ? p.chunks1 : 0));
  p.Ntot = L.Ntot;
  // columns per tile: the largest divisor of Ntot that is <= 256 and a multiple of 16
  int NB = std::min(L.Ntot, 256);
  while (L.Ntot % NB) NB -= 16;
  p.NB = NB;
  p.n_nb = L.Ntot / NB;
  const int esz = L.tf32 ? 4 : 2;
  if (L.tf32 && L.KC > 32) return CETPICK_ERR_BAD_ARG;      // the swizzle span is KC * 4 bytes <= 128
  p.tf32 = L.tf32;
  const int kstage_bytes = (NB <= 64) ? 256 : 128;
  p.NKB = std::max(1, kstage_bytes / (L.KC * esz));
  p.a_sub = TILE_M * L.KC * esz;
  p.b_sub = NB * L.KC * esz;
  const int stage_bytes = (p.a_sub + p.b_sub) * p.NKB;
  p.stages = std::max(2, std::min(MAX_STAGES, (int)((220 * 1024) / stage_bytes)));
  p.layout_type = L.KC * esz == 128 ? 2 : L.KC * esz == 64 ? 4 : 6;
  p.NIMG = L.NIMG; p.H = L.H; p.W = L.W;
  p.tiles_x = ceil_div(L.W, TILE_W);
  p.tiles_y = ceil_div(L.H, TILE_H);
  p.total_tiles = (long long)p.tiles_x * p.tiles_y * L.NIMG * p.n_nb;
  p.epi = L.epi; p.relu = L.relu; p.bias = L.bias; p.out = L.out;
  p.out_cstride = L.out_cstride; p.Ho = L.Ho; p.Wo = L.Wo; p.Cout = L.Cout;
  for (int t = 0; t < L.ntaps; ++t) {
    p.tdz[t] = (signed char)L.tap[t][0]; p.tdy[t] = (signed char)L.tap[t][1]; p.tdx[t] = (signed char)L.tap[t][2];
  }
  if (L.epi == EPI_UPCONV_2X2 && (L.Cout <= 0 || (L.Cout % 16) || L.Ntot != 4 * L.Cout)) return CETPICK_ERR_BAD_ARG;
  if (L.epi == EPI_F32_L2NORM_NCDHW && p.n_nb != 1) return CETPICK_ERR_UNSUPPORTED;

  const CUtensorMapSwizzle sw = swizzle_of(L.KC * esz / 2);
  const CUtensorMapDataType dt = L.tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  for (int s = 0; s < L.nsrc; ++s) {
    const cuuint64_t C = (cuuint64_t)L.C[s];
    cuuint64_t dims[4] = {C, (cuuint64_t)L.W, (cuuint64_t)L.H, (cuuint64_t)L.NIMG};
    cuuint64_t strides[3] = {C * esz, C * esz * L.W, C * esz * (cuuint64_t)L.W * L.H};
    cuuint32_t box[4] = {(cuuint32_t)L.KC, TILE_W, TILE_H, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&p.tmA[s], dt, 4, const_cast<void*>(L.src[s]), dims,
                     strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_cuda_err = "cuTensorMapEncodeTiled(A) failed: " + std::to_string((int)r); return CETPICK_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)L.KC, (cuuint64_t)p.nkb * L.Ntot};
    cuuint64_t strides[1] = {(cuuint64_t)L.KC * esz};
    cuuint32_t box[2] = {(cuuint32_t)L.KC, (cuuint32_t)NB};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&p.tmB, dt, 2, const_cast<void*>(L.wpk), dims, strides,
                     box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { g_cuda_err = "cuTensorMapEncodeTiled(B) failed: " + std::to_string((int)r); return CETPICK_ERR_CUDA; }
  }
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    cudaFuncAttributes fa;
    CETPICK_CUDA(cudaFuncGetAttributes(&fa, conv_tc_kernel));
    CETPICK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      227 * 1024 - (int)fa.sharedSizeBytes));
  }
  const int grid = (int)std::min<long long>(p.total_tiles, num_sms());
  conv_tc_kernel<<<grid, CONV_THREADS, smem, stream>>>(p);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

}  // namespace cetpick

using namespace cetpick;

#ifdef CETPICK_TEST_HOOKS   // test / tuning hooks: built into libcetpick_test_sm100a.so only (include/cetpick_test.h)
// C[M,N] = A[M,K] * B[N,K]^T as a 1x1 "convolution" over an (M/16) x 16 image -- validates the
// TMA / UMMA descriptor plumbing in isolation (tests/test_gpu_conv.py).
extern "C" int cetpick_selftest_gemm_bf16(const void* A, const void* B, float* C, int M, int N, int K,
                                          void* stream) {
  g_launches = 0;
  if (!A || !B || !C || M <= 0 || (M % 16) || N <= 0 || (N % 16) || K <= 0 || (K % 16)) return CETPICK_ERR_BAD_ARG;
  ConvLaunch L;
  L.nsrc = 1; L.src[0] = A; L.C[0] = K;
  L.NIMG = 1; L.H = M / 16; L.W = 16;
  L.wpk = B;
  L.KC = (K % 64 == 0) ? 64 : (K % 32 == 0) ? 32 : 16;
  // B is [N][K] row-major; the packed layout wants [k-block][N][KC]: only identical when K == KC,
  // so the self-test repacks on the caller's side (tests pass B already packed).
  L.ntaps = 1;
  L.Ntot = N;
  L.epi = EPI_F32_ROWMAJOR;
  L.out = C;
  return conv_tc_launch(L, static_cast<cudaStream_t>(stream));
}

#endif  // CETPICK_TEST_HOOKS
