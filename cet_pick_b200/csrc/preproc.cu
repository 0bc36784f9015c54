// GPU pre-processing = the step in front of the localisation path:
//   cet_pick/utils/loader.py:27-88  load_rec   (axis order, 2-slice z max "compress", global z-score)
//   cet_pick/utils/loader.py:90-121 preprocess (3-D Gaussian, z-score, quantise to 256 levels, min-max)
//   cet_pick/utils/loader.py:16-25  quantize
// The reference does all of it on the host in float64 (numpy / scipy.ndimage.gaussian_filter), ~10-30 s for a
// 1024x1024x256 tomogram; here every step is a bandwidth-bound kernel in float64 (same precision, so the 256-level
// output agrees except where a value sits within rounding noise of a level boundary).  Reductions are two-stage
// with a fixed grid: bitwise reproducible run to run.
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace cetpick {

namespace {

constexpr int PRE_THREADS = 256;
constexpr int PRE_MAX_RADIUS = 64;
constexpr int PRE_RED_BLOCKS = 1184;      // 8 x 148: partial sums per reduction

__device__ __forceinline__ double load_as_f64(const void* src, int dtype, long long i) {
  switch (dtype) {
    case 0: return (double)static_cast<const float*>(src)[i];
    case 1: return (double)static_cast<const short*>(src)[i];
    case 2: return (double)static_cast<const unsigned short*>(src)[i];
    case 3: return (double)static_cast<const signed char*>(src)[i];
    default: return static_cast<const double*>(src)[i];
  }
}

// out[j][a][b] = src[a*sa + b*sb + z*sz], z = j, or max over z in {2j, 2j+1} (those < Z) when pair_max
__global__ void __launch_bounds__(PRE_THREADS) pre_gather_kernel(const void* __restrict__ src, int dtype, long long A,
                                                                  long long B, long long J, long long sa, long long sb,
                                                                  long long sz, long long Z, int pair_max,
                                                                  double* __restrict__ out) {
  const long long n = A * B * J;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i % B, a = (i / B) % A, j = i / (A * B);
    const long long base = a * sa + b * sb;
    double v;
    if (pair_max) {
      v = load_as_f64(src, dtype, base + 2 * j * sz);
      if (2 * j + 1 < Z) {
        const double u = load_as_f64(src, dtype, base + (2 * j + 1) * sz);
        v = (u > v || u != u) ? u : v;              // np.max propagates NaN
      }
    } else {
      v = load_as_f64(src, dtype, base + j * sz);
    }
    out[i] = v;
  }
}

// block partial of sum(f(x)): f = x (mode 0) or (x - mean)^2 (mode 1)
__global__ void __launch_bounds__(PRE_THREADS) pre_sum_kernel(const double* __restrict__ x, long long n, int mode,
                                                               const double* __restrict__ mean_ptr,
                                                               double* __restrict__ partial) {
  __shared__ double s[PRE_THREADS];
  const double m = mode ? mean_ptr[0] : 0.0;
  // contiguous chunk per block, strided inside the block: deterministic for a fixed grid
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long lo = blockIdx.x * per, hi = min(n, lo + per);
  double acc = 0.0;
  for (long long i = lo + threadIdx.x; i < hi; i += PRE_THREADS) {
    const double v = x[i];
    acc += mode ? (v - m) * (v - m) : v;
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = PRE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}

// stats[0] = mean (mode 0) / stats[1] = sqrt(sum / n) (mode 1, population std like np.std)
__global__ void __launch_bounds__(PRE_THREADS) pre_sum_final_kernel(const double* __restrict__ partial, int nparts,
                                                                     long long n, int mode, double* __restrict__ stats) {
  __shared__ double s[PRE_THREADS];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += PRE_THREADS) acc += partial[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = PRE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (mode == 0) stats[0] = s[0] / (double)n;
    else stats[1] = sqrt(s[0] / (double)n);
  }
}

__global__ void __launch_bounds__(PRE_THREADS) pre_zscore_kernel(double* __restrict__ x, long long n,
                                                                  const double* __restrict__ stats) {
  const double m = stats[0], sd = stats[1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    x[i] = (x[i] - m) / sd;
}

struct GaussW { double w[2 * PRE_MAX_RADIUS + 1]; };

// scipy.ndimage.correlate1d with a symmetric kernel, mode='reflect' (d c b a | a b c d | d c b a), along one axis
// of a C-contiguous (n0, n1, n2) volume; same operation order as NI_Correlate1D's symmetric branch
// (centre tap, then pairs from the outside in), without FMA contraction.
__global__ void __launch_bounds__(PRE_THREADS) pre_gauss1d_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                                   long long n0, long long n1, long long n2, int axis,
                                                                   int radius, const __grid_constant__ GaussW gw) {
  const long long n = n0 * n1 * n2;
  const long long len = axis == 0 ? n0 : axis == 1 ? n1 : n2;
  const long long stride = axis == 0 ? n1 * n2 : axis == 1 ? n2 : 1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long c = axis == 0 ? i / (n1 * n2) : axis == 1 ? (i / n2) % n1 : i % n2;
    const double* line = in + (i - c * stride);
    auto at = [&](long long k) -> double {
      if (len == 1) return line[0];
      while (k < 0 || k >= len) k = k < 0 ? -k - 1 : 2 * len - 1 - k;   // reflect about the edge, repeatedly
      return line[k * stride];
    };
    double tmp = __dmul_rn(at(c), gw.w[radius]);
    for (int jj = -radius; jj < 0; ++jj)
      tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(at(c + jj), at(c - jj)), gw.w[radius + jj]));
    out[i] = tmp;
  }
}

// loader.py:16-25 quantize: x = 255*(x - mi)/r; clip(0,255); np.round (half to even) -> uint8
__global__ void __launch_bounds__(PRE_THREADS) pre_quantize_kernel(const double* __restrict__ x, long long n, double mi,
                                                                    double r, unsigned char* __restrict__ q) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double v = __ddiv_rn(__dmul_rn(255.0, __dadd_rn(x[i], -mi)), r);
    v = fmin(fmax(v, 0.0), 255.0);
    q[i] = (unsigned char)rint(v);                  // rint = round half to even
  }
}

__global__ void __launch_bounds__(PRE_THREADS) pre_minmax_kernel(const unsigned char* __restrict__ q, long long n,
                                                                  int* __restrict__ mm /*[2]: min, max (pre-set 255, 0)*/) {
  int lo = 255, hi = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = q[i];
    lo = min(lo, v); hi = max(hi, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&mm[0], lo); atomicMax(&mm[1], hi); }
}

__global__ void pre_minmax_init_kernel(int* mm) { mm[0] = 255; mm[1] = 0; }

// loader.py:105,120: (im - min) / (max - min) on the uint8 levels, true division in float64
template <typename T>
__global__ void __launch_bounds__(PRE_THREADS) pre_normalize_kernel(const unsigned char* __restrict__ q, long long n,
                                                                     const int* __restrict__ mm, T* __restrict__ out) {
  const int lo = mm[0];
  const double range = (double)(mm[1] - mm[0]);       // 0 -> division by zero -> nan/inf like numpy
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (T)((double)(q[i] - lo) / range);
}

int grid_for(long long n) {
  return (int)std::min<long long>(std::max<long long>(1, (n + PRE_THREADS - 1) / PRE_THREADS), (long long)num_sms() * 16);
}

}  // namespace
}  // namespace cetpick

using namespace cetpick;

extern "C" int cetpick_pre_gather_f64(const void* src, int src_dtype, int64_t A, int64_t B, int64_t J, int64_t sa,
                                      int64_t sb, int64_t sz, int64_t Z, int pair_max, double* out, void* stream) {
  g_launches = 0;
  if (!src || !out || A <= 0 || B <= 0 || J <= 0 || Z <= 0 || src_dtype < 0 || src_dtype > 4) return CETPICK_ERR_BAD_ARG;
  pre_gather_kernel<<<grid_for(A * B * J), PRE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      src, src_dtype, A, B, J, sa, sb, sz, Z, pair_max, out);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

// stats (device, 2 doubles) = [mean, population std] of x; ws: >= cetpick_pre_stats_workspace_bytes() bytes
extern "C" int cetpick_pre_stats_workspace_bytes(size_t* bytes) {
  if (!bytes) return CETPICK_ERR_BAD_ARG;
  *bytes = (size_t)PRE_RED_BLOCKS * sizeof(double);
  return CETPICK_OK;
}

extern "C" int cetpick_pre_mean_std_f64(const double* x, int64_t n, double* stats, void* ws, size_t ws_bytes, void* stream) {
  g_launches = 0;
  if (!x || !stats || n <= 0) return CETPICK_ERR_BAD_ARG;
  if (!ws || ws_bytes < (size_t)PRE_RED_BLOCKS * sizeof(double)) return CETPICK_ERR_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(ws);
  const int nb = (int)std::min<long long>(PRE_RED_BLOCKS, (n + PRE_THREADS - 1) / PRE_THREADS);
  for (int mode = 0; mode < 2; ++mode) {
    pre_sum_kernel<<<nb, PRE_THREADS, 0, s>>>(x, n, mode, stats, partial);
    CETPICK_LAUNCH_CHECK();
    pre_sum_final_kernel<<<1, PRE_THREADS, 0, s>>>(partial, nb, n, mode, stats);
    CETPICK_LAUNCH_CHECK();
  }
  return CETPICK_OK;
}

extern "C" int cetpick_pre_zscore_f64(double* x, int64_t n, const double* stats, void* stream) {
  g_launches = 0;
  if (!x || !stats || n <= 0) return CETPICK_ERR_BAD_ARG;
  pre_zscore_kernel<<<grid_for(n), PRE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, n, stats);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

// one axis of scipy.ndimage.gaussian_filter (mode='reflect'); weights_host = the 2*radius+1 kernel taps
extern "C" int cetpick_pre_gauss1d_f64(const double* in, double* out, int64_t n0, int64_t n1, int64_t n2, int axis,
                                       const double* weights_host, int radius, void* stream) {
  g_launches = 0;
  if (!in || !out || in == out || !weights_host || n0 <= 0 || n1 <= 0 || n2 <= 0 || axis < 0 || axis > 2 || radius < 0)
    return CETPICK_ERR_BAD_ARG;
  if (radius > PRE_MAX_RADIUS) return CETPICK_ERR_UNSUPPORTED;
  GaussW gw;
  memset(&gw, 0, sizeof(gw));
  memcpy(gw.w, weights_host, (size_t)(2 * radius + 1) * sizeof(double));
  pre_gauss1d_kernel<<<grid_for(n0 * n1 * n2), PRE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n0, n1, n2,
                                                                                                   axis, radius, gw);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

extern "C" int cetpick_pre_quantize_u8(const double* x, int64_t n, double mi, double ma, unsigned char* q, void* stream) {
  g_launches = 0;
  if (!x || !q || n <= 0) return CETPICK_ERR_BAD_ARG;
  pre_quantize_kernel<<<grid_for(n), PRE_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, n, mi, ma - mi, q);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

// out = (q - min(q)) / (max(q) - min(q)); minmax: device int[2] scratch (also the result); out_f64: 1 = double, 0 = float
extern "C" int cetpick_pre_minmax_normalize(const unsigned char* q, int64_t n, int* minmax, void* out, int out_f64,
                                            void* stream) {
  g_launches = 0;
  if (!q || !minmax || !out || n <= 0) return CETPICK_ERR_BAD_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pre_minmax_init_kernel<<<1, 1, 0, s>>>(minmax);
  CETPICK_LAUNCH_CHECK();
  pre_minmax_kernel<<<grid_for(n), PRE_THREADS, 0, s>>>(q, n, minmax);
  CETPICK_LAUNCH_CHECK();
  if (out_f64) pre_normalize_kernel<double><<<grid_for(n), PRE_THREADS, 0, s>>>(q, n, minmax, static_cast<double*>(out));
  else pre_normalize_kernel<float><<<grid_for(n), PRE_THREADS, 0, s>>>(q, n, minmax, static_cast<float*>(out));
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}
