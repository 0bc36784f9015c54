#ifdef CETPICK_TEST_HOOKS   // hardware probes: built into libcetpick_test_sm100a.so only (include/cetpick_test.h)
// Hardware probe (test hook, not on the product path): which shifted / strided shared-memory
// descriptors does tcgen05.mma accept for a TMA-written, hardware-swizzled K-major tile?
// One CTA: TMA-load A_big[R][KC] and B[32][KC], then D[128][32] = A_big[rows(m)] * B^T with the A
// descriptor starting `r0` rows into the tile, 8-row groups `sbo_bytes` apart and the given
// base_offset field.  The answers decide how conv_tc.cu may reuse one halo tile for several taps.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <mutex>

namespace cetpick {
namespace {

struct alignas(64) ProbeParams {
  CUtensorMap tmA, tmB;
  int R, KC, r0, sbo, base_offset, layout_type;
  float* out;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full, bar_done;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.R * p.KC * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar_full, 1); ptx::mbar_init(&bar_done, 1); ptx::fence_barrier_init(); }
  if (warp == 1) { ptx::tmem_alloc(&s_tmem, 32); ptx::tmem_relinquish(); }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar_full, (uint32_t)(p.R + 32) * p.KC * 2);
    for (int r = 0; r < p.R; r += 128)   // TMA boxes are limited to 256 rows: load in 128-row boxes
      ptx::tma_load_2d(sA + (size_t)r * p.KC * 2, &p.tmA, &bar_full, 0, r);
    ptx::tma_load_2d(sB, &p.tmB, &bar_full, 0, 0);
    ptx::mbar_wait(&bar_full, 0);
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::make_idesc_bf16(128, 32);
    for (int k = 0; k < p.KC / 16; ++k) {
      uint64_t da = ptx::make_smem_desc(ptx::smem_u32(sA) + p.r0 * p.KC * 2 + k * 32, p.sbo, p.layout_type);
      da |= (uint64_t)(p.base_offset & 7) << 49;
      const uint64_t db = ptx::make_smem_desc(ptx::smem_u32(sB) + k * 32, 16u * p.KC, p.layout_type);
      ptx::umma_bf16(tmem, da, db, idesc, k > 0);
    }
    ptx::umma_commit(&bar_done);
  }
  __syncwarp();
  ptx::mbar_wait(&bar_done, 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t v[16];
    __syncwarp();
    ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 16; ++i) p.out[(size_t)threadIdx.x * 32 + c0 + i] = __uint_as_float(v[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 32);
}

// MMA issue-rate probe: one thread issues `iters` tcgen05.mma (M = 128, N, K = 16) back to back on
// operands that stay in shared memory (contents irrelevant), cycling through `ntap` A start offsets
// (a_step bytes apart) and `ndst` accumulators, then commits and waits.  out[cta] = cycles per MMA.
struct RateParams {
  int N, KC, sbo_a, a_step, ntap, ndst, iters, layout_type;
  float* out;
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(const RateParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  // finite operand bytes (0x3c3c.. = small positive bf16)
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar_done, 1); ptx::fence_barrier_init(); }
  if (warp == 1) { ptx::tmem_alloc(&s_tmem, 512); ptx::tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 0) {
    const uint32_t a_hi = ptx::smem_desc_hi((uint32_t)p.sbo_a, (uint32_t)p.layout_type);
    const uint32_t b_hi = ptx::smem_desc_hi((uint32_t)(16 * p.KC), (uint32_t)p.layout_type);
    const uint32_t a_lo = ptx::smem_desc_lo(ptx::smem_u32(smem));
    const uint32_t b_lo = ptx::smem_desc_lo(ptx::smem_u32(smem + 64 * 1024));
    const uint32_t idesc = ptx::make_idesc_bf16(128, p.N);
    const int k16 = p.KC / 16;
    long long t0 = 0, t1 = 0;
    if (ptx::elect_one()) {
      t0 = clock64();
      int tap = 0, dst = 0, kk = 0;
      for (int i = 0; i < p.iters; ++i) {
        ptx::umma_bf16_lohi(tmem + (uint32_t)(dst * p.N), a_lo + (uint32_t)((tap * p.a_step + kk * 32) >> 4), a_hi,
                            b_lo + (uint32_t)((tap * p.N * p.KC * 2 / 1 % (24 * 1024) + kk * 32) >> 4), b_hi, idesc);
        if (++kk == k16) { kk = 0; if (++tap == p.ntap) { tap = 0; if (++dst == p.ndst) dst = 0; } }
      }
      ptx::umma_commit(&bar_done);
    }
    __syncwarp();
    ptx::mbar_wait(&bar_done, 0);
    t1 = clock64();
    const long long dt = __shfl_sync(0xffffffffu, t1, 0) - __shfl_sync(0xffffffffu, t0, 0);
    long long mx = t0;
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) p.out[blockIdx.x] = (float)((double)(t1 - mx) / p.iters);
    (void)dt;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, 512);
}

// Same probe for a CTA pair: tcgen05.mma.cta_group::2 (M = 256 over two SMs; each CTA supplies its own 128 rows
// of A and N/2 rows of B).  The leader CTA's elected thread issues; out[pair] = cycles per MMA.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma_rate2_kernel(const RateParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_done;
  __shared__ uint32_t s_tmem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  uint32_t cta_rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar_done, 1); ptx::fence_barrier_init(); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&s_tmem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  ptx::tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  ptx::tc_fence_after();
  const uint32_t tmem = s_tmem;
  if (warp == 0 && cta_rank == 0) {
    const uint32_t a_hi = ptx::smem_desc_hi((uint32_t)p.sbo_a, (uint32_t)p.layout_type);
    const uint32_t b_hi = ptx::smem_desc_hi((uint32_t)(16 * p.KC), (uint32_t)p.layout_type);
    const uint32_t a_lo = ptx::smem_desc_lo(ptx::smem_u32(smem));
    const uint32_t b_lo = ptx::smem_desc_lo(ptx::smem_u32(smem + 64 * 1024));
    const uint32_t idesc = ptx::make_idesc_bf16(256, p.N);
    const int k16 = p.KC / 16;
    long long t0 = 0, t1 = 0;
    if (ptx::elect_one()) {
      t0 = clock64();
      int tap = 0, dst = 0, kk = 0;
      for (int i = 0; i < p.iters; ++i) {
        const uint32_t al = a_lo + (uint32_t)((tap * p.a_step + kk * 32) >> 4);
        const uint32_t bl = b_lo + (uint32_t)(((tap * (p.N / 2) * p.KC * 2) % (24 * 1024) + kk * 32) >> 4);
        asm volatile(
            "{\n\t"
            ".reg .b64 da, db;\n\t"
            "mov.b64 da, {%1, %2};\n\t"
            "mov.b64 db, {%3, %4};\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, 1;\n\t"
            "}\n" ::"r"(tmem + (uint32_t)(dst * p.N)),
            "r"(al), "r"(a_hi), "r"(bl), "r"(b_hi), "r"(idesc)
            : "memory");
        if (++kk == k16) { kk = 0; if (++tap == p.ntap) { tap = 0; if (++dst == p.ndst) dst = 0; } }
      }
      asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                       ptx::smem_u32(&bar_done)), "h"((uint16_t)1)
                   : "memory");
    }
    __syncwarp();
    ptx::mbar_wait(&bar_done, 0);
    t1 = clock64();
    long long mx = t0;
    for (int o = 16; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (threadIdx.x == 0) p.out[blockIdx.x >> 1] = (float)((double)(t1 - mx) / p.iters);
  }
  ptx::tc_fence_before();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
}  // namespace
}  // namespace cetpick

using namespace cetpick;

extern "C" int cetpick_probe_umma(const void* A_big, int R, const void* B, int KC, int r0, int sbo_bytes,
                                  int base_offset, float* out, void* stream) {
  if (!A_big || !B || !out || R < 128 || (R % 128) || (KC != 16 && KC != 32 && KC != 64)) return CETPICK_ERR_BAD_ARG;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess || !f)
    return CETPICK_ERR_CUDA;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(f);
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  const CUtensorMapSwizzle sw = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : KC == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                                                  : CU_TENSOR_MAP_SWIZZLE_32B;
  cuuint32_t es[2] = {1, 1};
  {
    cuuint64_t dims[2] = {(cuuint64_t)KC, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)KC * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, 128};
    if (enc(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A_big), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return CETPICK_ERR_CUDA;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)KC, 32};
    cuuint64_t strides[1] = {(cuuint64_t)KC * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, 32};
    if (enc(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(B), dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return CETPICK_ERR_CUDA;
  }
  p.R = R; p.KC = KC; p.r0 = r0; p.sbo = sbo_bytes; p.base_offset = base_offset;
  p.layout_type = KC == 64 ? 2 : KC == 32 ? 4 : 6;
  p.out = out;
  const size_t smem = (size_t)(R + 32) * KC * 2 + 2048;
  CETPICK_CUDA(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

// Test / tuning hook: tcgen05.mma issue rate for (N, swizzle width KC, A group stride, tap pattern).
extern "C" int cetpick_probe_mma_rate(int N, int KC, int sbo_a, int a_step, int ntap, int ndst, int iters,
                                      float* out_cycles, int grid, void* stream) {
  if (!out_cycles || N < 16 || N > 256 || (N % 16) || (KC != 16 && KC != 32 && KC != 64) || ntap < 1 || ndst < 1 ||
      ndst * N > 512 || iters < 1 || grid < 1)
    return CETPICK_ERR_BAD_ARG;
  RateParams p;
  p.N = N; p.KC = KC; p.sbo_a = sbo_a; p.a_step = a_step; p.ntap = ntap; p.ndst = ndst; p.iters = iters;
  p.layout_type = KC == 64 ? 2 : KC == 32 ? 4 : 6;
  p.out = out_cycles;
  const size_t smem = 97 * 1024 + 1024;
  CETPICK_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_rate_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

// Tuning hook: the same measurement for a CTA pair (tcgen05.mma.cta_group::2, M = 256).  grid = number of pairs.
extern "C" int cetpick_probe_mma_rate2(int N, int KC, int sbo_a, int a_step, int ntap, int ndst, int iters,
                                       float* out_cycles, int pairs, void* stream) {
  if (!out_cycles || N < 32 || N > 256 || (N % 32) || (KC != 16 && KC != 32 && KC != 64) || ntap < 1 || ndst < 1 ||
      ndst * N > 512 || iters < 1 || pairs < 1)
    return CETPICK_ERR_BAD_ARG;
  RateParams p;
  p.N = N; p.KC = KC; p.sbo_a = sbo_a; p.a_step = a_step; p.ntap = ntap; p.ndst = ndst; p.iters = iters;
  p.layout_type = KC == 64 ? 2 : KC == 32 ? 4 : 6;
  p.out = out_cycles;
  const size_t smem = 97 * 1024 + 1024;
  CETPICK_CUDA(cudaFuncSetAttribute(mma_rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_rate2_kernel<<<2 * pairs, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

#endif  // CETPICK_TEST_HOOKS
