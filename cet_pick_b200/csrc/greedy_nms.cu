// Greedy distance-threshold suppression on the device: cet_pick/models/decode.py:42-79
// `non_maximum_suppression_3d` (called by `tomo_decode_classify`, decode.py:108-120, on the semiclass
// path detectors/tomo_det_classify.py:112,146).
//
// Reference semantics: visit voxels in descending score order, stop at the first
// score <= threshold; a visited voxel that is not in the suppressed set S becomes a pick and adds
// {i + delta} to S for every FLAT-index delta i*(H*W) + j*W + k with i^2+j^2+k^2 <= (scale*d/2)^2
// (deltas wrap across row / plane ends exactly as in the reference; out-of-range targets are inert).
// The reference's visit order among equal scores is numpy's unstable argsort; here ties are visited
// in ascending index order (same canonical order as the decode path and the oracle).
//
// The sequential loop is a dependency graph: candidate c (rank r_c in the sorted order) is a pick
// iff no PICK of smaller rank lies at c - delta.  It is resolved data-parallel: every round each
// undecided candidate looks at its earlier-ranked neighbours; it becomes SUPPRESSED if one of them is
// a pick, a PICK if all of them are decided and none is a pick, and stays undecided otherwise.  The
// smallest undecided rank is always decided, so the rounds terminate; on real heat-maps the
// dependency chains are a few tens of voxels long.
//
// Pipeline: compact (score > threshold) voxels into 64-bit composites (monotone key << 32 | ~index)
// -> stable radix sort descending (sort.cu) -> scatter ranks into a dense int32 map -> resolve rounds -> ordered
// compaction of the picks -> (score, x, y, z) writer.
#include "common.cuh"

#include "sort.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace cetpick {

namespace {

constexpr int GN_THREADS = 256;
constexpr int GN_MAX_DELTAS = 8192;
enum : uint8_t { ST_UNDECIDED = 0, ST_PICK = 1, ST_SUPPRESSED = 2 };

__device__ __forceinline__ uint32_t gn_key(float v) {        // monotone: larger float -> larger key
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float gn_unkey(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

struct GnCounters { unsigned long long n_cand; uint32_t undecided; uint32_t pad; };

__global__ void gn_fill_kernel(int32_t* __restrict__ rank, size_t n, GnCounters* ctr) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    rank[i] = 0x7fffffff;       // "not a candidate": never earlier than anyone
  if (blockIdx.x == 0 && threadIdx.x == 0) { ctr->n_cand = 0ull; ctr->undecided = 0u; }
}

// candidates: score > threshold (NaN never compares greater: not a candidate)
__global__ void __launch_bounds__(GN_THREADS) gn_compact_kernel(const float* __restrict__ heat, size_t n, double threshold,
                                                                 unsigned long long* __restrict__ cand,
                                                                 unsigned long long cap, GnCounters* ctr) {
  const int lane = threadIdx.x & 31;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n_round = (n + stride - 1) / stride * stride;      // warp-uniform trip count
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const float v = (i < n) ? heat[i] : 0.f;
    const bool take = (i < n) && ((double)v > threshold);   // numpy compares the float32 score with a Python float
    const unsigned b = __ballot_sync(0xffffffffu, take);
    if (b == 0) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&ctr->n_cand, (unsigned long long)__popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (take) {
      const unsigned long long o = base + __popc(b & ((1u << lane) - 1u));
      if (o < cap) cand[o] = ((unsigned long long)gn_key(v) << 32) | (unsigned long long)(~(uint32_t)i);
    }
  }
}

__global__ void gn_rank_kernel(const unsigned long long* __restrict__ sorted, uint32_t n, int32_t* __restrict__ rank,
                               uint8_t* __restrict__ state) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  rank[~(uint32_t)sorted[i]] = (int32_t)i;
  state[i] = ST_UNDECIDED;
}

__global__ void __launch_bounds__(GN_THREADS) gn_resolve_kernel(const unsigned long long* __restrict__ sorted, uint32_t n,
                                                                 const int32_t* __restrict__ rank,
                                                                 uint8_t* state, const long long* __restrict__ deltas,
                                                                 int n_deltas, long long n_vox, GnCounters* ctr) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (state[i] != ST_UNDECIDED) return;
  const long long idx = (long long)(~(uint32_t)sorted[i]);
  bool pending = false, hit = false;
  for (int k = 0; k < n_deltas; ++k) {
    const long long nb = idx + deltas[k];
    if (nb < 0 || nb >= n_vox || nb == idx) continue;
    const int32_t r = rank[nb];
    if (r >= (int32_t)i) continue;                         // later in the visit order (or not a candidate)
    const uint8_t s = reinterpret_cast<volatile uint8_t*>(state)[r];
    if (s == ST_PICK) { hit = true; break; }
    if (s == ST_UNDECIDED) pending = true;
  }
  // an earlier pick in range suppresses regardless of what the still-undecided neighbours become
  if (hit) state[i] = ST_SUPPRESSED;
  else if (!pending) state[i] = ST_PICK;
  else atomicAdd(&ctr->undecided, 1u);
}

__global__ void gn_flag_kernel(const uint8_t* __restrict__ state, uint32_t n, uint8_t* __restrict__ flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = state[i] == ST_PICK;
}

__global__ void gn_write_kernel(const unsigned long long* __restrict__ picks, const int* __restrict__ n_picks,
                                int H, int W, long long max_out, float* __restrict__ scores,
                                int32_t* __restrict__ coords) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= *n_picks || i >= max_out) return;
  const unsigned long long c = picks[i];
  const uint32_t idx = ~(uint32_t)c;
  const int hw = H * W;
  const int z = (int)(idx / (uint32_t)hw), r = (int)(idx - (uint32_t)z * (uint32_t)hw);
  scores[i] = gn_unkey((uint32_t)(c >> 32));
  coords[3 * i + 0] = r % W;      // decode.py:66-69: np.unravel_index -> (xx, yy, zz)
  coords[3 * i + 1] = r / W;
  coords[3 * i + 2] = z;
}

struct GnLayout {
  size_t off_ctr, off_npick, off_deltas, off_rank, off_cand, off_sorted, off_state, off_flags, off_picks, off_tmp, total;
  size_t tmp_bytes;
};

GnLayout gn_layout(size_t n_vox, size_t cap) {
  GnLayout L;
  const size_t sort_tmp = sort_tmp_bytes(cap, false), sel_tmp = compact_tmp_bytes(cap);
  L.tmp_bytes = std::max(sort_tmp, sel_tmp);
  size_t o = 0;
  L.off_ctr = o;    o = align_up(o + sizeof(GnCounters), 256);
  L.off_npick = o;  o = align_up(o + 16, 256);
  L.off_deltas = o; o = align_up(o + (size_t)GN_MAX_DELTAS * 8, 256);
  L.off_rank = o;   o = align_up(o + n_vox * 4, 256);
  L.off_cand = o;   o = align_up(o + cap * 8, 256);
  L.off_sorted = o; o = align_up(o + cap * 8, 256);
  L.off_state = o;  o = align_up(o + cap, 256);
  L.off_flags = o;  o = align_up(o + cap, 256);
  L.off_picks = o;  o = align_up(o + cap * 8, 256);
  L.off_tmp = o;    o = align_up(o + L.tmp_bytes, 256);
  L.total = o;
  return L;
}

// decode.py:44-56: flat-index deltas of the ball of radius r = scale*d/2 (double arithmetic like numpy)
std::vector<long long> gn_deltas(double d, double scale, int H, int W) {
  const double r = scale * d / 2.0;
  const int width = (int)std::ceil(r);
  std::vector<long long> out;
  for (int i = -width; i <= width; ++i)
    for (int j = -width; j <= width; ++j)
      for (int k = -width; k <= width; ++k)
        if ((double)(i * i + j * j + k * k) <= r * r)
          out.push_back((long long)i * H * W + (long long)j * W + k);
  return out;
}

}  // namespace
}  // namespace cetpick

using namespace cetpick;

extern "C" int cetpick_greedy_nms_workspace_bytes(int64_t D, int64_t H, int64_t W, int64_t max_candidates,
                                                  size_t* bytes) {
  if (!bytes || D <= 0 || H <= 0 || W <= 0 || max_candidates <= 0) return CETPICK_ERR_BAD_ARG;
  const uint64_t n = (uint64_t)D * H * W;
  if (n > 0x7fffffffull) return CETPICK_ERR_BAD_ARG;
  *bytes = gn_layout((size_t)n, (size_t)std::min<uint64_t>((uint64_t)max_candidates, n)).total;
  return CETPICK_OK;
}

// Synchronous by design (the number of picks is the result's length, as in the reference): the stream is
// synchronised once per block of resolve rounds and at the end.
extern "C" int cetpick_greedy_nms_f32(const float* heat, int64_t D, int64_t H, int64_t W, double d, double scale,
                                      double threshold, int64_t max_candidates, float* scores, int32_t* coords,
                                      int64_t max_out, int64_t* n_out, int* rounds_out, void* ws, size_t ws_bytes,
                                      void* stream) {
  g_launches = 0;
  if (!heat || !scores || !coords || !n_out || D <= 0 || H <= 0 || W <= 0 || max_candidates <= 0 || max_out < 0)
    return CETPICK_ERR_BAD_ARG;
  const uint64_t n64 = (uint64_t)D * H * W;
  if (n64 > 0x7fffffffull || !(d >= 0.0) || !(scale >= 0.0)) return CETPICK_ERR_BAD_ARG;
  const size_t n = (size_t)n64, cap = (size_t)std::min<uint64_t>((uint64_t)max_candidates, n64);
  const GnLayout L = gn_layout(n, cap);
  if (!ws || ws_bytes < L.total || (reinterpret_cast<uintptr_t>(ws) & 255)) return CETPICK_ERR_WORKSPACE;
  const std::vector<long long> deltas = gn_deltas(d, scale, (int)H, (int)W);
  if (deltas.size() > (size_t)GN_MAX_DELTAS) return CETPICK_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  GnCounters* ctr = reinterpret_cast<GnCounters*>(base + L.off_ctr);
  int* n_pick = reinterpret_cast<int*>(base + L.off_npick);
  long long* d_deltas = reinterpret_cast<long long*>(base + L.off_deltas);
  int32_t* rank = reinterpret_cast<int32_t*>(base + L.off_rank);
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(base + L.off_cand);
  unsigned long long* sorted = reinterpret_cast<unsigned long long*>(base + L.off_sorted);
  uint8_t* state = reinterpret_cast<uint8_t*>(base + L.off_state);
  uint8_t* flags = reinterpret_cast<uint8_t*>(base + L.off_flags);
  unsigned long long* picks = reinterpret_cast<unsigned long long*>(base + L.off_picks);
  void* tmp = base + L.off_tmp;

  const int sms = num_sms();
  CETPICK_CUDA(cudaMemcpyAsync(d_deltas, deltas.data(), deltas.size() * 8, cudaMemcpyHostToDevice, s));
  gn_fill_kernel<<<sms * 8, GN_THREADS, 0, s>>>(rank, n, ctr);
  CETPICK_LAUNCH_CHECK();
  gn_compact_kernel<<<sms * 8, GN_THREADS, 0, s>>>(heat, n, threshold, cand, (unsigned long long)cap, ctr);
  CETPICK_LAUNCH_CHECK();
  GnCounters h;
  CETPICK_CUDA(cudaMemcpyAsync(&h, ctr, sizeof(h), cudaMemcpyDeviceToHost, s));
  CETPICK_CUDA(cudaStreamSynchronize(s));
  if (h.n_cand > (unsigned long long)cap) return CETPICK_ERR_WORKSPACE;    // more voxels above threshold than max_candidates
  const uint32_t nc = (uint32_t)h.n_cand;
  *n_out = 0;
  if (rounds_out) *rounds_out = 0;
  if (nc == 0) return CETPICK_OK;

  size_t tb = L.tmp_bytes;
  if (int rc = radix_sort_desc_u64(cand, sorted, nullptr, nullptr, nc, tmp, tb, s, nullptr)) return rc;
  const int gb = (int)ceil_div<uint32_t>(nc, GN_THREADS);
  gn_rank_kernel<<<gb, GN_THREADS, 0, s>>>(sorted, nc, rank, state);
  CETPICK_LAUNCH_CHECK();
  int rounds = 0;
  for (;;) {
    // a block of rounds per host round trip; the counter is reset before the last round of the block
    for (int k = 0; k < 4; ++k) {
      if (k == 3) CETPICK_CUDA(cudaMemsetAsync(&ctr->undecided, 0, 4, s));
      gn_resolve_kernel<<<gb, GN_THREADS, 0, s>>>(sorted, nc, rank, state, d_deltas, (int)deltas.size(), (long long)n, ctr);
      CETPICK_LAUNCH_CHECK();
      ++rounds;
    }
    CETPICK_CUDA(cudaMemcpyAsync(&h, ctr, sizeof(h), cudaMemcpyDeviceToHost, s));
    CETPICK_CUDA(cudaStreamSynchronize(s));
    if (h.undecided == 0) break;
    if (rounds > 4 * (int)nc + 8) return CETPICK_ERR_STATE;   // cannot happen: one decision per round at least
  }
  if (rounds_out) *rounds_out = rounds;
  gn_flag_kernel<<<gb, GN_THREADS, 0, s>>>(state, nc, flags);
  CETPICK_LAUNCH_CHECK();
  tb = L.tmp_bytes;
  if (int rc = compact_flagged(flags, nc, sorted, picks, nullptr, n_pick, tmp, tb, s, nullptr)) return rc;
  gn_write_kernel<<<gb, GN_THREADS, 0, s>>>(picks, n_pick, (int)H, (int)W, (long long)max_out, scores, coords);
  CETPICK_LAUNCH_CHECK();
  int np = 0;
  CETPICK_CUDA(cudaMemcpyAsync(&np, n_pick, sizeof(int), cudaMemcpyDeviceToHost, s));
  CETPICK_CUDA(cudaStreamSynchronize(s));
  *n_out = np;               // may exceed max_out: the caller then sees how much room a full result needs
  return CETPICK_OK;
}
