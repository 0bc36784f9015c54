// Exploration-step candidate generator, device side (first half):
//   cet_pick/utils/image.py:138-183 `get_potential_coords_pyramid` = difference of 3-D Gaussians (float64,
//   scipy) -> `_nms_xy` (:81-87) -> max over the pyramid -> `non_maximum_suppression_3d` (:42-79) with d = 14.
// Everything there is float64 (numpy arrays from `load_rec`/`gaussian_filter`, torch double tensors), so the
// kernels here are the float64 twins of `cetpick_nms_f32` (decode.cu) and `cetpick_greedy_nms_f32`
// (greedy_nms.cu): an NMS map with a free (kz,ky,kx) window, and the greedy distance suppression with 64-bit keys.
// The Gaussians are `cetpick_pre_gauss1d_f64` (preproc.cu).
//
// Greedy suppression with double scores: the 64-bit (key,index) composite of the float32 version does not fit, so
// candidates are compacted IN INDEX ORDER (compact_flagged, sort.cu), keyed with the monotone 64-bit
// image of the double, and sorted with a STABLE descending radix sort: equal scores keep ascending index order
// (the canonical tie order of this repo; the reference's is numpy's unstable argsort).  The dependency rounds are
// those of greedy_nms.cu.  Scores are returned as float32 like the reference's `scores` array (:61,68).
#include "common.cuh"

#include "sort.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace cetpick {

namespace {

constexpr int EX_THREADS = 256;
constexpr int EX_MAX_DELTAS = 8192;
enum : uint8_t { EX_UNDECIDED = 0, EX_PICK = 1, EX_SUPPRESSED = 2 };

template <typename T>
__global__ void __launch_bounds__(EX_THREADS) nms_window_kernel(const T* __restrict__ heat, T* __restrict__ out, int D,
                                                                 int H, int W, int pz, int py, int px, size_t total) {
  const size_t n_vox = (size_t)D * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / n_vox, r = i - b * n_vox;
    const int x = (int)(r % W), y = (int)((r / W) % H), z = (int)(r / ((size_t)W * H));
    const T* hb = heat + b * n_vox;
    const T v = hb[r];
    T m = -INFINITY;
    bool nan = false;
    for (int dz = -pz; dz <= pz; ++dz) {
      const int zz = z + dz;
      if (zz < 0 || zz >= D) continue;
      for (int dy = -py; dy <= py; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -px; dx <= px; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= W) continue;
          const T u = hb[((size_t)zz * H + yy) * W + xx];
          nan |= (u != u);
          m = u > m ? u : m;
        }
      }
    }
    out[i] = (!nan && m == v) ? v : v * (T)0;      // heat * (hmax == heat).float(): suppressed -> 0 with heat's sign
  }
}

__device__ __forceinline__ unsigned long long key64(double v) {           // monotone: larger double -> larger key
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double unkey64(unsigned long long k) {
  return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k));
}

__global__ void __launch_bounds__(EX_THREADS) ex_flag_kernel(const double* __restrict__ x, size_t n, double threshold,
                                                              uint8_t* __restrict__ flag, int32_t* __restrict__ rank) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    flag[i] = x[i] > threshold;                    // NaN never compares greater: not a candidate
    rank[i] = 0x7fffffff;
  }
}

__global__ void ex_key_kernel(const double* __restrict__ x, const uint32_t* __restrict__ idx, uint32_t n,
                              unsigned long long* __restrict__ keys) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = key64(x[idx[i]]);
}

__global__ void ex_rank_kernel(const uint32_t* __restrict__ idx_sorted, uint32_t n, int32_t* __restrict__ rank,
                               uint8_t* __restrict__ state) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rank[idx_sorted[i]] = (int32_t)i;
  state[i] = EX_UNDECIDED;
}

__global__ void __launch_bounds__(EX_THREADS) ex_resolve_kernel(const uint32_t* __restrict__ idx_sorted, uint32_t n,
                                                                 const int32_t* __restrict__ rank, uint8_t* state,
                                                                 const long long* __restrict__ deltas, int n_deltas,
                                                                 long long n_vox, uint32_t* undecided) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (state[i] != EX_UNDECIDED) return;
  const long long idx = (long long)idx_sorted[i];
  bool pending = false, hit = false;
  for (int k = 0; k < n_deltas; ++k) {
    const long long nb = idx + deltas[k];
    if (nb < 0 || nb >= n_vox || nb == idx) continue;
    const int32_t r = rank[nb];
    if (r >= (int32_t)i) continue;                 // later in the visit order (or not a candidate)
    const uint8_t s = reinterpret_cast<volatile uint8_t*>(state)[r];
    if (s == EX_PICK) { hit = true; break; }
    if (s == EX_UNDECIDED) pending = true;
  }
  if (hit) state[i] = EX_SUPPRESSED;
  else if (!pending) state[i] = EX_PICK;
  else atomicAdd(undecided, 1u);
}

__global__ void ex_pickflag_kernel(const uint8_t* __restrict__ state, uint32_t n, uint8_t* __restrict__ flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = state[i] == EX_PICK;
}

__global__ void ex_write_kernel(const uint32_t* __restrict__ pos, const int* __restrict__ n_picks,
                                const unsigned long long* __restrict__ keys_sorted,
                                const uint32_t* __restrict__ idx_sorted, int H, int W, long long max_out,
                                float* __restrict__ scores, int32_t* __restrict__ coords) {
  const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= *n_picks || j >= max_out) return;
  const uint32_t i = pos[j];
  const uint32_t idx = idx_sorted[i];
  const int hw = H * W;
  const int z = (int)(idx / (uint32_t)hw), r = (int)(idx - (uint32_t)z * (uint32_t)hw);
  scores[j] = (float)unkey64(keys_sorted[i]);      // np.float32 slot of a float64 score (image.py:61,68)
  coords[3 * j + 0] = r % W;
  coords[3 * j + 1] = r / W;
  coords[3 * j + 2] = z;
}

// Candidate patches (datasets/tomo_pre_proj_angle_select_new3d_vol.py:117-128 `extract_subvols`): for candidate
// (x, y, z) sum the z-slab [z - sz/2, z + sz/2] of the (sy x sx) window, min-max normalise in float64, cast to
// float32.  One CTA per candidate; the slab sum is recomputed in the second pass instead of staged.
__global__ void __launch_bounds__(EX_THREADS) ex_patch_kernel(const double* __restrict__ vol, int D, int H, int W,
                                                               const int32_t* __restrict__ coords, int hz, int hy, int hx,
                                                               float* __restrict__ out) {
  __shared__ double s_mn[EX_THREADS], s_mx[EX_THREADS];
  const int n = blockIdx.x;
  const int x = coords[3 * n], y = coords[3 * n + 1], z = coords[3 * n + 2];
  const int z0 = z - hz, z1 = min(z + hz + 1, D);          // numpy clips the end of the slice
  const int py = 2 * hy, px = 2 * hx, np_ = py * px;
  auto slab = [&](int e) -> double {
    const int yy = y - hy + e / px, xx = x - hx + e % px;
    double a = 0.0;
    for (int zz = z0; zz < z1; ++zz) a += vol[((size_t)zz * H + yy) * W + xx];     // np.sum(axis=0): slice after slice
    return a;
  };
  double mn = INFINITY, mx = -INFINITY;
  for (int e = threadIdx.x; e < np_; e += EX_THREADS) {
    const double a = slab(e);
    mn = fmin(mn, a); mx = fmax(mx, a);
  }
  s_mn[threadIdx.x] = mn; s_mx[threadIdx.x] = mx;
  __syncthreads();
  for (int o = EX_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_mn[threadIdx.x] = fmin(s_mn[threadIdx.x], s_mn[threadIdx.x + o]);
      s_mx[threadIdx.x] = fmax(s_mx[threadIdx.x], s_mx[threadIdx.x + o]);
    }
    __syncthreads();
  }
  mn = s_mn[0]; mx = s_mx[0];
  const double range = mx - mn;
  for (int e = threadIdx.x; e < np_; e += EX_THREADS) out[(size_t)n * np_ + e] = (float)((slab(e) - mn) / range);
}

struct ExLayout {
  size_t off_ctr, off_deltas, off_flag, off_rank, off_idx, off_idx2, off_keys, off_keys2, off_state, off_pflag, off_pos,
      off_tmp, total;
  size_t tmp_bytes;
};

ExLayout ex_layout(size_t n_vox, size_t cap) {
  ExLayout L;
  const size_t t1 = compact_tmp_bytes(n_vox), t2 = sort_tmp_bytes(cap, true), t3 = compact_tmp_bytes(cap);
  L.tmp_bytes = std::max(t1, std::max(t2, t3));
  size_t o = 0;
  L.off_ctr = o;    o = align_up(o + 64, 256);
  L.off_deltas = o; o = align_up(o + (size_t)EX_MAX_DELTAS * 8, 256);
  L.off_flag = o;   o = align_up(o + n_vox, 256);
  L.off_rank = o;   o = align_up(o + n_vox * 4, 256);
  L.off_idx = o;    o = align_up(o + n_vox * 4, 256);      // sized for every voxel: the count is only known on the device
  L.off_idx2 = o;   o = align_up(o + cap * 4, 256);
  L.off_keys = o;   o = align_up(o + cap * 8, 256);
  L.off_keys2 = o;  o = align_up(o + cap * 8, 256);
  L.off_state = o;  o = align_up(o + cap, 256);
  L.off_pflag = o;  o = align_up(o + cap, 256);
  L.off_pos = o;    o = align_up(o + cap * 4, 256);
  L.off_tmp = o;    o = align_up(o + L.tmp_bytes, 256);
  L.total = o;
  return L;
}

std::vector<long long> ex_deltas(double d, double scale, int H, int W) {
  const double r = scale * d / 2.0;
  const int width = (int)std::ceil(r);
  std::vector<long long> out;
  for (int i = -width; i <= width; ++i)
    for (int j = -width; j <= width; ++j)
      for (int k = -width; k <= width; ++k)
        if ((double)(i * i + j * j + k * k) <= r * r) out.push_back((long long)i * H * W + (long long)j * W + k);
  return out;
}

template <typename T>
int nms_window_launch(const T* heat, T* out, int64_t B, int64_t D, int64_t H, int64_t W, int kz, int ky, int kx,
                      cudaStream_t s) {
  const size_t total = (size_t)B * D * H * W;
  const int grid = (int)std::min<size_t>(ceil_div<size_t>(total, EX_THREADS), (size_t)num_sms() * 16);
  nms_window_kernel<T><<<grid, EX_THREADS, 0, s>>>(heat, out, (int)D, (int)H, (int)W, (kz - 1) / 2, (ky - 1) / 2,
                                                    (kx - 1) / 2, total);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

}  // namespace
}  // namespace cetpick

using namespace cetpick;

// out = heat * (max_pool3d(heat, (kz,ky,kx), stride 1, pad (k-1)/2) == heat) for float32 (dtype 0) or float64 (1):
// image.py:81-105 `_nms_xy` (1,k,k), `_nms_z` (k,1,1), `_nms` (k,k,k); also decode.py's (3,k,k).
extern "C" int cetpick_nms_window(const void* heat, void* out, int dtype, int64_t B, int64_t D, int64_t H, int64_t W,
                                  int kz, int ky, int kx, void* stream) {
  g_launches = 0;
  if (!heat || !out || B <= 0 || D <= 0 || H <= 0 || W <= 0) return CETPICK_ERR_BAD_ARG;
  if (kz < 1 || ky < 1 || kx < 1 || !(kz & 1) || !(ky & 1) || !(kx & 1)) return CETPICK_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == 0) return nms_window_launch(static_cast<const float*>(heat), static_cast<float*>(out), B, D, H, W, kz, ky, kx, s);
  if (dtype == 1) return nms_window_launch(static_cast<const double*>(heat), static_cast<double*>(out), B, D, H, W, kz, ky, kx, s);
  return CETPICK_ERR_BAD_ARG;
}

extern "C" int cetpick_greedy_nms_f64_workspace_bytes(int64_t D, int64_t H, int64_t W, int64_t max_candidates,
                                                      size_t* bytes) {
  if (!bytes || D <= 0 || H <= 0 || W <= 0 || max_candidates <= 0) return CETPICK_ERR_BAD_ARG;
  const uint64_t n = (uint64_t)D * H * W;
  if (n > 0x7fffffffull) return CETPICK_ERR_BAD_ARG;
  *bytes = ex_layout((size_t)n, (size_t)std::min<uint64_t>((uint64_t)max_candidates, n)).total;
  return CETPICK_OK;
}

// float64 twin of cetpick_greedy_nms_f32 (same arguments and conventions; synchronous).
extern "C" int cetpick_greedy_nms_f64(const double* vol, int64_t D, int64_t H, int64_t W, double d, double scale,
                                      double threshold, int64_t max_candidates, float* scores, int32_t* coords,
                                      int64_t max_out, int64_t* n_out, int* rounds_out, void* ws, size_t ws_bytes,
                                      void* stream) {
  g_launches = 0;
  if (!vol || !scores || !coords || !n_out || D <= 0 || H <= 0 || W <= 0 || max_candidates <= 0 || max_out < 0)
    return CETPICK_ERR_BAD_ARG;
  const uint64_t n64 = (uint64_t)D * H * W;
  if (n64 > 0x7fffffffull || !(d >= 0.0) || !(scale >= 0.0)) return CETPICK_ERR_BAD_ARG;
  const size_t n = (size_t)n64, cap = (size_t)std::min<uint64_t>((uint64_t)max_candidates, n64);
  const ExLayout L = ex_layout(n, cap);
  if (!ws || ws_bytes < L.total || (reinterpret_cast<uintptr_t>(ws) & 255)) return CETPICK_ERR_WORKSPACE;
  const std::vector<long long> deltas = ex_deltas(d, scale, (int)H, (int)W);
  if (deltas.size() > (size_t)EX_MAX_DELTAS) return CETPICK_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  int* n_sel = reinterpret_cast<int*>(base + L.off_ctr);                 // [0] candidates, [1] picks
  uint32_t* undecided = reinterpret_cast<uint32_t*>(base + L.off_ctr + 16);
  long long* d_deltas = reinterpret_cast<long long*>(base + L.off_deltas);
  uint8_t* flag = reinterpret_cast<uint8_t*>(base + L.off_flag);
  int32_t* rank = reinterpret_cast<int32_t*>(base + L.off_rank);
  uint32_t* idx = reinterpret_cast<uint32_t*>(base + L.off_idx);
  uint32_t* idx2 = reinterpret_cast<uint32_t*>(base + L.off_idx2);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base + L.off_keys);
  unsigned long long* keys2 = reinterpret_cast<unsigned long long*>(base + L.off_keys2);
  uint8_t* state = reinterpret_cast<uint8_t*>(base + L.off_state);
  uint8_t* pflag = reinterpret_cast<uint8_t*>(base + L.off_pflag);
  uint32_t* pos = reinterpret_cast<uint32_t*>(base + L.off_pos);
  void* tmp = base + L.off_tmp;

  const int sms = num_sms();
  CETPICK_CUDA(cudaMemcpyAsync(d_deltas, deltas.data(), deltas.size() * 8, cudaMemcpyHostToDevice, s));
  ex_flag_kernel<<<sms * 8, EX_THREADS, 0, s>>>(vol, n, threshold, flag, rank);
  CETPICK_LAUNCH_CHECK();
  size_t tb = L.tmp_bytes;
  if (int rc = compact_flagged(flag, (uint32_t)n, nullptr, nullptr, idx, n_sel, tmp, tb, s, nullptr)) return rc;   // candidates in index order
  int nc_i = 0;
  CETPICK_CUDA(cudaMemcpyAsync(&nc_i, n_sel, sizeof(int), cudaMemcpyDeviceToHost, s));
  CETPICK_CUDA(cudaStreamSynchronize(s));
  *n_out = 0;
  if (rounds_out) *rounds_out = 0;
  if ((size_t)nc_i > cap) return CETPICK_ERR_WORKSPACE;     // more voxels above threshold than max_candidates
  if (nc_i == 0) return CETPICK_OK;
  const uint32_t nc = (uint32_t)nc_i;
  const int gb = (int)ceil_div<uint32_t>(nc, EX_THREADS);
  ex_key_kernel<<<gb, EX_THREADS, 0, s>>>(vol, idx, nc, keys);
  CETPICK_LAUNCH_CHECK();
  tb = L.tmp_bytes;
  if (int rc = radix_sort_desc_u64(keys, keys2, idx, idx2, nc, tmp, tb, s, nullptr)) return rc;   // stable
  ex_rank_kernel<<<gb, EX_THREADS, 0, s>>>(idx2, nc, rank, state);
  CETPICK_LAUNCH_CHECK();
  int rounds = 0;
  for (;;) {
    uint32_t h_und = 0;
    for (int k = 0; k < 4; ++k) {
      if (k == 3) CETPICK_CUDA(cudaMemsetAsync(undecided, 0, 4, s));
      ex_resolve_kernel<<<gb, EX_THREADS, 0, s>>>(idx2, nc, rank, state, d_deltas, (int)deltas.size(), (long long)n, undecided);
      CETPICK_LAUNCH_CHECK();
      ++rounds;
    }
    CETPICK_CUDA(cudaMemcpyAsync(&h_und, undecided, 4, cudaMemcpyDeviceToHost, s));
    CETPICK_CUDA(cudaStreamSynchronize(s));
    if (h_und == 0) break;
    if (rounds > 4 * (int)nc + 8) return CETPICK_ERR_STATE;
  }
  if (rounds_out) *rounds_out = rounds;
  ex_pickflag_kernel<<<gb, EX_THREADS, 0, s>>>(state, nc, pflag);
  CETPICK_LAUNCH_CHECK();
  tb = L.tmp_bytes;
  if (int rc = compact_flagged(pflag, nc, nullptr, nullptr, pos, n_sel + 1, tmp, tb, s, nullptr)) return rc;
  ex_write_kernel<<<gb, EX_THREADS, 0, s>>>(pos, n_sel + 1, keys2, idx2, (int)H, (int)W, (long long)max_out, scores, coords);
  CETPICK_LAUNCH_CHECK();
  int np = 0;
  CETPICK_CUDA(cudaMemcpyAsync(&np, n_sel + 1, sizeof(int), cudaMemcpyDeviceToHost, s));
  CETPICK_CUDA(cudaStreamSynchronize(s));
  *n_out = np;
  return CETPICK_OK;
}

// out[n] = min-max normalised z-slab sum around candidate n (float32, (2*(sub_y/2)) x (2*(sub_x/2)) per patch).
// coords: int32 device [n][3] = (x, y, z); windows must lie inside the volume in x, y and at their low z end
// (CETPICK_ERR_BAD_ARG otherwise is the caller's check: the kernel does not bounds-check).
extern "C" int cetpick_extract_subvols_f64(const double* vol, int64_t D, int64_t H, int64_t W, const int32_t* coords,
                                           int64_t n, int sub_z, int sub_y, int sub_x, float* out, void* stream) {
  g_launches = 0;
  if (!vol || !coords || !out || D <= 0 || H <= 0 || W <= 0 || n < 0 || sub_z < 1 || sub_y < 2 || sub_x < 2)
    return CETPICK_ERR_BAD_ARG;
  if (n == 0) return CETPICK_OK;
  ex_patch_kernel<<<(unsigned)n, EX_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(vol, (int)D, (int)H, (int)W, coords,
                                                                                    sub_z / 2, sub_y / 2, sub_x / 2, out);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}
