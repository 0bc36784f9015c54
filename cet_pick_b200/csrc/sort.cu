// Stable LSD radix sort (8 passes of 8 bits) and flag compaction, see sort.cuh.
// A pass = per-block digit histogram -> exclusive scan of the (digit-major) block counts -> stable scatter: every block
// owns a contiguous tile and walks it in sub-tiles of 256 keys; inside a sub-tile a key's rank among equal digits is
// (keys of earlier warps) + (earlier lanes of its own warp, __match_any_sync), so input order survives every pass.
#include "sort.cuh"

#include <algorithm>

namespace cetpick {
namespace {

constexpr int ST = 256;                 // threads per block = keys per sub-tile
constexpr int MAX_BLOCKS = 4096;        // tiles grow instead of the grid beyond this

struct Tiling { uint32_t tile, blocks; };
Tiling tiling(uint32_t n) {
  uint32_t tile = 2048;
  while ((uint64_t)tile * MAX_BLOCKS < n) tile *= 2;
  return {tile, std::max<uint32_t>(1, ceil_div<uint32_t>(n, tile))};
}

__device__ __forceinline__ uint32_t digit_of(unsigned long long k, int shift) {
  return 255u - (uint32_t)((k >> shift) & 0xFFull);          // descending order
}

__global__ void __launch_bounds__(ST) sort_hist_kernel(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t tile,
                                                       int shift, uint32_t* __restrict__ block_hist /*[256][blocks]*/) {
  __shared__ uint32_t s_cnt[256];
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t b0 = blockIdx.x * tile, b1 = min(n, b0 + tile);
  for (uint32_t i = b0 + threadIdx.x; i < b1; i += ST) atomicAdd(&s_cnt[digit_of(keys[i], shift)], 1u);
  __syncthreads();
  block_hist[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s_cnt[threadIdx.x];
}

// in-place exclusive scan of `len` counters by ONE block (len = 256 * blocks <= 1 M): chunked, 1024 threads
__global__ void __launch_bounds__(1024) scan_excl_kernel(uint32_t* __restrict__ a, uint32_t len, uint32_t* __restrict__ total) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t per = ceil_div<uint32_t>(len, 1024u);
  // each thread owns `per` consecutive entries: sum, block scan of the sums, then rewrite
  const uint32_t lo = min(len, threadIdx.x * per), hi = min(len, lo + per);
  uint32_t sum = 0;
  for (uint32_t i = lo; i < hi; ++i) sum += a[i];
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;
    if (lane == 31 && total) *total = wi;
  }
  __syncthreads();
  uint32_t run = s_warp[warp] + incl - sum;
  for (uint32_t i = lo; i < hi; ++i) { const uint32_t v = a[i]; a[i] = run; run += v; }
  (void)s_carry;
}

template <bool VALS>
__global__ void __launch_bounds__(ST) sort_scatter_kernel(const unsigned long long* __restrict__ keys_in,
                                                          unsigned long long* __restrict__ keys_out,
                                                          const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ vals_out,
                                                          uint32_t n, uint32_t tile, int shift,
                                                          const uint32_t* __restrict__ block_off /*[256][blocks], scanned*/) {
  __shared__ uint32_t s_base[256];             // next output slot of each digit for this block
  __shared__ uint32_t s_wcnt[ST / 32][256];    // per warp, per digit: keys of the current sub-tile
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  s_base[threadIdx.x] = block_off[(size_t)threadIdx.x * gridDim.x + blockIdx.x];
  const uint32_t b0 = blockIdx.x * tile, b1 = min(n, b0 + tile);
  for (uint32_t t0 = b0; t0 < b1; t0 += ST) {
#pragma unroll
    for (int w = 0; w < ST / 32; ++w) s_wcnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = t0 + threadIdx.x;
    const bool act = i < b1;
    const unsigned long long k = act ? keys_in[i] : 0ull;
    const uint32_t d = act ? digit_of(k, shift) : (0x1000u + (uint32_t)lane);
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (act && rank_in_warp == 0) s_wcnt[warp][d] = __popc(peers);
    __syncthreads();
    {   // thread d: exclusive prefix over the warps, advance the block's slot of digit d
      uint32_t run = s_base[threadIdx.x];
#pragma unroll
      for (int w = 0; w < ST / 32; ++w) { const uint32_t c = s_wcnt[w][threadIdx.x]; s_wcnt[w][threadIdx.x] = run; run += c; }
      s_base[threadIdx.x] = run;
    }
    __syncthreads();
    if (act) {
      const uint32_t pos = s_wcnt[warp][d] + rank_in_warp;
      keys_out[pos] = k;
      if (VALS) vals_out[pos] = vals_in[i];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(ST) flag_count_kernel(const uint8_t* __restrict__ flags, uint32_t n, uint32_t tile,
                                                        uint32_t* __restrict__ block_cnt) {
  __shared__ uint32_t s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const uint32_t b0 = blockIdx.x * tile, b1 = min(n, b0 + tile);
  uint32_t c = 0;
  for (uint32_t i = b0 + threadIdx.x; i < b1; i += ST) c += flags[i] ? 1u : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) block_cnt[blockIdx.x] = s_cnt;
}

__global__ void __launch_bounds__(ST) flag_scatter_kernel(const uint8_t* __restrict__ flags, uint32_t n, uint32_t tile,
                                                          const unsigned long long* __restrict__ src,
                                                          unsigned long long* __restrict__ out_u64, uint32_t* __restrict__ out_idx,
                                                          const uint32_t* __restrict__ block_off) {
  __shared__ uint32_t s_w[ST / 32];
  __shared__ uint32_t s_base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = block_off[blockIdx.x];
  const uint32_t b0 = blockIdx.x * tile, b1 = min(n, b0 + tile);
  for (uint32_t t0 = b0; t0 < b1; t0 += ST) {
    __syncthreads();
    const uint32_t i = t0 + threadIdx.x;
    const bool f = (i < b1) && flags[i];
    const unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) s_w[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = s_base;
    for (int w = 0; w < warp; ++w) before += s_w[w];
    if (f) {
      const uint32_t pos = before + __popc(bal & ((1u << lane) - 1u));
      if (out_u64) out_u64[pos] = src[i];
      else out_idx[pos] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t t = 0;
      for (int w = 0; w < ST / 32; ++w) t += s_w[w];
      s_base += t;
    }
  }
}

__global__ void store_count_kernel(const uint32_t* total, int* n_out) { *n_out = (int)*total; }

}  // namespace

size_t sort_tmp_bytes(size_t n, bool with_values) {
  const Tiling t = tiling((uint32_t)std::min<size_t>(n, 0xffffffffu));
  return align_up(n * 8, 256) + (with_values ? align_up(n * 4, 256) : 0) + align_up((size_t)256 * t.blocks * 4, 256) + 256;
}

size_t compact_tmp_bytes(size_t n) {
  const Tiling t = tiling((uint32_t)std::min<size_t>(n, 0xffffffffu));
  return align_up((size_t)t.blocks * 4, 256) + 256;
}

int radix_sort_desc_u64(const unsigned long long* keys_in, unsigned long long* keys_out, const uint32_t* vals_in,
                        uint32_t* vals_out, uint32_t n, void* tmp, size_t tmp_bytes, cudaStream_t s, int64_t* launches) {
  if (n == 0) return CETPICK_OK;
  const bool vals = vals_in != nullptr;
  if (!keys_in || !keys_out || (vals && !vals_out) || !tmp || tmp_bytes < sort_tmp_bytes(n, vals)) return CETPICK_ERR_WORKSPACE;
  const Tiling t = tiling(n);
  char* base = static_cast<char*>(tmp);
  unsigned long long* alt = reinterpret_cast<unsigned long long*>(base);
  size_t o = align_up((size_t)n * 8, 256);
  uint32_t* valt = nullptr;
  if (vals) { valt = reinterpret_cast<uint32_t*>(base + o); o += align_up((size_t)n * 4, 256); }
  uint32_t* hist = reinterpret_cast<uint32_t*>(base + o);
  const unsigned long long* kin = keys_in;
  const uint32_t* vin = vals_in;
  for (int pass = 0; pass < 8; ++pass) {
    unsigned long long* kout = (pass & 1) ? keys_out : alt;      // 8 passes: the last one lands in keys_out
    uint32_t* vout = (pass & 1) ? vals_out : valt;
    const int shift = 8 * pass;
    sort_hist_kernel<<<t.blocks, ST, 0, s>>>(kin, n, t.tile, shift, hist);
    CETPICK_LAUNCH_CHECK();
    scan_excl_kernel<<<1, 1024, 0, s>>>(hist, 256u * t.blocks, nullptr);
    CETPICK_LAUNCH_CHECK();
    if (vals) sort_scatter_kernel<true><<<t.blocks, ST, 0, s>>>(kin, kout, vin, vout, n, t.tile, shift, hist);
    else sort_scatter_kernel<false><<<t.blocks, ST, 0, s>>>(kin, kout, nullptr, nullptr, n, t.tile, shift, hist);
    CETPICK_LAUNCH_CHECK();
    kin = kout; vin = vout;
    if (launches) *launches += 3;
  }
  return CETPICK_OK;
}

int compact_flagged(const uint8_t* flags, uint32_t n, const unsigned long long* src, unsigned long long* out_u64,
                    uint32_t* out_idx, int* n_out, void* tmp, size_t tmp_bytes, cudaStream_t s, int64_t* launches) {
  if (!flags || !n_out || (!out_u64 && !out_idx) || (out_u64 && !src) || !tmp || tmp_bytes < compact_tmp_bytes(n))
    return CETPICK_ERR_WORKSPACE;
  if (n == 0) { CETPICK_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int), s)); return CETPICK_OK; }
  const Tiling t = tiling(n);
  char* base = static_cast<char*>(tmp);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(base);
  uint32_t* total = reinterpret_cast<uint32_t*>(base + align_up((size_t)t.blocks * 4, 256));
  flag_count_kernel<<<t.blocks, ST, 0, s>>>(flags, n, t.tile, cnt);
  CETPICK_LAUNCH_CHECK();
  scan_excl_kernel<<<1, 1024, 0, s>>>(cnt, t.blocks, total);
  CETPICK_LAUNCH_CHECK();
  flag_scatter_kernel<<<t.blocks, ST, 0, s>>>(flags, n, t.tile, src, out_u64, out_idx, cnt);
  CETPICK_LAUNCH_CHECK();
  store_count_kernel<<<1, 1, 0, s>>>(total, n_out);
  CETPICK_LAUNCH_CHECK();
  if (launches) *launches += 4;
  return CETPICK_OK;
}

}  // namespace cetpick
