// Fused full-resolution block of the detector trunk on tcgen05 tensor cores (sm_100a):
//     [Conv2d 3x3 pad 1 + BN + ReLU] -> [Conv2d 3x3 pad 1 + BN + ReLU] (-> MaxPool2d(2, ceil_mode=True))
// = conv1/conv2 of DownConv level 0 and of the last UpConv (cet_pick/models/networks/unet.py:198-249 and
// :375-399, after the concat), the 32-channel layers whose intermediate map costs a 4.3 GB HBM round trip
// per 1024x1024x256 tomogram when the two convolutions run as separate kernels.
//
// Both convolutions are "marches" down the image (conv_march.cu): the three ky taps are stacked on GEMM-N, an
// input row feeds the accumulators of three output rows that live in a ring of TMEM slots.  conv1's epilogue
// warps do bias/ReLU/bf16 and store the finished row -- in the SWIZZLE_64B K-major image a TMA load would have
// produced -- into a ring of rows in SHARED memory, which is conv2's A operand (fence.proxy.async hand-off,
// like the stem's converter warps).  The intermediate map never leaves the SM.
//
// x tiling: one CTA = one 128-pixel M-tile (TMEM holds 8 + 8 slots of 32 columns); the CTAs of a CLUSTER cover
// a full image row (cluster size = ceil(W / 128) <= 8), so conv2's one-pixel x halo is always a pixel that a
// neighbour CTA of the same cluster has just computed: the owner's epilogue thread stores it into the
// neighbour's ring through distributed shared memory (st.shared::cluster) and arrives on the neighbour's
// mbarrier.  No recompute in x, and the results are bit-identical to the two separate kernels.
// y strips overlap by one conv1 row (two input rows) per side.
//
// Roles (640 threads, 1 CTA/SM, persistent over (image, row strip)):
//   warp 0 TMA producer | warp 1 UMMA issuer of conv1 | warp 3 UMMA issuer of conv2 (tcgen05.commit tracks the issuing
//   thread's own UMMAs, so the two convolutions pipeline independently) | warp 2 TMEM allocator, then relay (tells the
//   neighbours when a ring row of this CTA is free again) | warps 4-11 conv1 epilogue, two groups alternating rows
//   (TMEM -> ring) | warps 12-19 conv2 epilogue, two groups (TMEM -> global, fused pool)
#include "conv_block.cuh"
#include "common.cuh"
#include "conv_march.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <algorithm>

namespace cetpick {

namespace {

constexpr int BLK_THREADS = 640;
constexpr int COUT = 32;
constexpr int SP = 8, SLOTS = 6;            // physical / logical TMEM slots per convolution (phantom ring, conv_march.cu)
constexpr int TMEM_C1 = 0, TMEM_C2 = SP * COUT;
constexpr int MAX_LAG = 12;                 // conv2 consumes ring row j while conv1 is fed input row j + lag
constexpr int MAX_STAGES = 10, MAX_RS = 16;
constexpr int RING_PIX = 130;               // x0-1 .. x0+128
constexpr int RING_ROW = 9216;              // 130 * 64 bytes, 1024-aligned
constexpr int W2BLK = 3 * COUT * 64;        // one dx block of conv2's weights (KC = 32)

struct alignas(64) BlockParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB1, tmB2;
  int nsrc, NIMG, H, W;
  int R, nchunk;                // rows per strip, strips per image
  long long total_strips;       // per cluster
  int stages, RS, NC;
  int lag, flags;               // tuning: bit0 = acquire.cluster polls on ring_full (A/B only)
  float bias1[COUT], bias2[COUT];
  __nv_bfloat16* out;
  __nv_bfloat16* pool_out;
};

template <int KC1>
struct G1 {
  static constexpr int PIX = KC1 * 2;
  static constexpr int K16 = KC1 / 16;
  static constexpr int BOX_BYTES = 130 * PIX;
  static constexpr int STAGE_BYTES = (BOX_BYTES + 1023) / 1024 * 1024;
  static constexpr int WBLK = 3 * COUT * PIX;
  static constexpr uint32_t LAYOUT = KC1 == 64 ? 2 : KC1 == 32 ? 4 : 6;
};

// Bring-up aid: when a host-mapped buffer is registered (cetpick_block_debug_buffer), a wait that times out records
// (cta, warp, wait-site tag, two values) there before trapping, so a protocol deadlock can be read post mortem.
__device__ uint32_t* g_blk_dbg = nullptr;

__device__ __noinline__ void dbg_timeout(uint32_t tag, uint32_t a, uint32_t b) {
  uint32_t* d = g_blk_dbg;
  if (d && (threadIdx.x & 31) == 0) {
    const uint32_t k = atomicAdd(d, 1u);
    if (k < 60) {
      d[4 + k * 4 + 0] = (blockIdx.x << 8) | (threadIdx.x >> 5);
      d[4 + k * 4 + 1] = tag;
      d[4 + k * 4 + 2] = a;
      d[4 + k * 4 + 3] = b;
    }
    __threadfence_system();
  }
}

template <bool CLUSTER>
__device__ __forceinline__ void wait_t(uint64_t* bar, uint32_t parity, uint32_t tag, uint32_t a = 0, uint32_t b = 0) {
  if (!CLUSTER && ptx::mbar_try_wait(bar, parity)) return;      // the common case: no clock read, no loop
  const long long t0 = clock64();
  uint32_t ok, spins = 0;
  bool reported = false;
  do {
    if (CLUSTER)
      asm volatile(
          "{\n\t"
          ".reg .pred P1;\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
          "selp.b32 %0, 1, 0, P1;\n\t"
          "}\n"
          : "=r"(ok)
          : "r"(ptx::smem_u32(bar)), "r"(parity)
          : "memory");
    else
      ok = ptx::mbar_try_wait(bar, parity) ? 1u : 0u;
    if (!ok && (++spins & 0xFF) == 0) {
      const long long dt = clock64() - t0;
      if (dt > 3000000000LL && !reported) { dbg_timeout(tag, a, b); reported = true; }
      if (dt > 6000000000LL) __trap();
    }
  } while (!ok);
}

struct Strip { int ma, mb, c1a, c1b, i_lo, i_hi, img; };

__device__ __forceinline__ void decode_strip(const BlockParams& p, long long k, Strip& s) {
  const int ch = (int)(k % p.nchunk);
  s.img = (int)(k / p.nchunk);
  s.ma = ch * p.R;
  s.mb = min(s.ma + p.R, p.H);
  s.c1a = max(s.ma - 1, 0);
  s.c1b = min(s.mb + 1, p.H);
  s.i_lo = max(s.c1a - 1, 0);
  s.i_hi = min(s.c1b, p.H - 1);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same variable in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t caddr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// asynchronous 16-byte store into another CTA's shared memory that reports its bytes to an mbarrier of THAT CTA
// (tx-count): no fence and no release on the sending side, the receiving barrier completes when the bytes have landed
__device__ __forceinline__ void st_async_v4(uint32_t caddr, const uint4& v, uint32_t cbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(caddr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cbar)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t caddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cta.b64 _, [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;\n" ::: "memory");
}

template <int KC1>
__global__ void __launch_bounds__(BLK_THREADS, 1) conv_block_kernel(const __grid_constant__ BlockParams p) {
  using G = G1<KC1>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t in_full[MAX_STAGES], in_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t c1_full[SLOTS], c1_empty[SLOTS], c2_full[SLOTS], c2_empty[SLOTS];
  __shared__ __align__(8) uint64_t ring_full[MAX_RS], ring_empty[MAX_RS], nb_empty[2][MAX_RS], bar_w;
  __shared__ uint32_t s_tmem_base;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int nblk1 = p.nsrc * 3;
  uint8_t* sW1 = smem;
  uint8_t* sW2 = sW1 + (((size_t)nblk1 * G::WBLK + 1023) & ~(size_t)1023);
  uint8_t* sRing = sW2 + (((size_t)3 * W2BLK + 1023) & ~(size_t)1023);
  uint8_t* sA = sRing + (size_t)p.RS * RING_ROW;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int x0 = (int)rank * 128;
  const bool has_left = rank > 0, has_right = (int)rank + 1 < p.NC;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tmA[0]);
    if (p.nsrc > 1) ptx::prefetch_tensormap(&p.tmA[1]);
    ptx::prefetch_tensormap(&p.tmB1);
    ptx::prefetch_tensormap(&p.tmB2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&in_full[s], 1); ptx::mbar_init(&in_empty[s], 1); }
    for (int s = 0; s < SLOTS; ++s) {
      // *_empty: the draining group AND the observing group arrive (8 warps): the issuer must not start a slot's next
      // phase before both have seen the current one (a parity wait cannot tell phase n from phase n + 2)
      ptx::mbar_init(&c1_full[s], 1); ptx::mbar_init(&c1_empty[s], 8);
      ptx::mbar_init(&c2_full[s], 1); ptx::mbar_init(&c2_empty[s], 8);
    }
    for (int s = 0; s < p.RS; ++s) {
      ptx::mbar_init(&ring_full[s], 4u);       // the four warps of the owning group; the neighbours' edge pixels arrive as tx bytes
      ptx::mbar_init(&ring_empty[s], 1);
      ptx::mbar_init(&nb_empty[0][s], 1);
      ptx::mbar_init(&nb_empty[1][s], 1);
    }
    ptx::mbar_init(&bar_w, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&s_tmem_base, 512);
    ptx::tmem_relinquish();
  }
  // the ring starts as zeros: halo pixels outside the image are never written and stay zero (conv2's padding)
  for (int i = threadIdx.x; i < p.RS * RING_ROW / 16; i += BLK_THREADS) reinterpret_cast<uint4*>(sRing)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_all();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if ((warp >= 4 && warp < 8) || (warp >= 12 && warp < 16)) {   // every accumulator slot starts at zero: all UMMAs accumulate
    const uint32_t row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp < 8 ? TMEM_C1 : TMEM_C2);
    for (int c = 0; c < SP * COUT; c += 16) ptx::tmem_st16_fill(row + c, 0u);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  cluster_sync_all();             // barriers and rings of every CTA of the cluster are ready before any remote access
  ptx::tc_fence_after();

  const long long kfirst = cluster_id_x(), kstep = num_clusters_x();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bar_w, (uint32_t)(nblk1 * G::WBLK + 3 * W2BLK));
      for (int b = 0; b < nblk1; ++b) ptx::tma_load_2d(sW1 + (size_t)b * G::WBLK, &p.tmB1, &bar_w, 0, b * 3 * COUT);
      for (int b = 0; b < 3; ++b) ptx::tma_load_2d(sW2 + (size_t)b * W2BLK, &p.tmB2, &bar_w, 0, b * 3 * COUT);
      int stage = 0;
      uint32_t phase = 0;
      for (long long k = kfirst; k < p.total_strips; k += kstep) {
        Strip s;
        decode_strip(p, k, s);
        for (int i = s.i_lo; i <= s.i_hi; ++i)
          for (int src = 0; src < p.nsrc; ++src) {
            wait_t<false>(&in_empty[stage], phase ^ 1u, 1, (uint32_t)i, (uint32_t)stage);
            ptx::mbar_arrive_expect_tx(&in_full[stage], (uint32_t)G::BOX_BYTES);
            ptx::tma_load_4d(sA + (size_t)stage * G::STAGE_BYTES, &p.tmA[src], &in_full[stage], 0, x0 - 1, i, s.img);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    // ================================ UMMA issuer, conv1 ===========================
    constexpr uint32_t A1_HI = ptx::smem_desc_hi(8 * G::PIX, G::LAYOUT), B1_HI = ptx::smem_desc_hi(8 * G::PIX, G::LAYOUT);
    constexpr uint32_t IDESC0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint32_t sW1_lo = ptx::smem_desc_lo(ptx::smem_u32(sW1)), sA_lo = ptx::smem_desc_lo(ptx::smem_u32(sA));
    int stage = 0;
    uint32_t phase = 0;
    uint32_t umask1 = 0;                // per logical TMEM slot: parity of the number of rows that have used it
    wait_t<false>(&bar_w, 0, 2);
    for (long long k = kfirst; k < p.total_strips; k += kstep) {
      Strip s;
      decode_strip(p, k, s);
      int r1_touched = s.c1a;
      for (int i = s.i_lo; i <= s.i_hi; ++i) {
        // input row i feeds conv1 rows i-1, i, i+1 (inside [c1a, c1b))
        const int r_lo = max(s.c1a, i - 1), r_hi = min(s.c1b - 1, i + 1);
        const int n = r_hi - r_lo + 1;
        for (; r1_touched <= r_hi; ++r1_touched) {          // rows touched for the first time: slot must be drained
          const uint32_t sl = (uint32_t)r1_touched % SLOTS;
          wait_t<false>(&c1_empty[sl], ((umask1 >> sl) & 1u) ^ 1u, 3, (uint32_t)r1_touched, (uint32_t)i);
          umask1 ^= 1u << sl;
        }
        const uint32_t idesc = IDESC0 | ((uint32_t)(n * COUT >> 3) << 17);
        const uint32_t boff = (uint32_t)(r_lo - (i - 1)) * ((COUT * G::PIX) >> 4);
        const uint32_t d = tmem_base + TMEM_C1 + ((uint32_t)(i + SLOTS - 1) % SLOTS + (uint32_t)(r_lo - (i - 1))) * COUT;
        for (int src = 0; src < p.nsrc; ++src) {
          wait_t<false>(&in_full[stage], phase, 4, (uint32_t)i, (uint32_t)stage);
          ptx::tc_fence_after();
          const uint32_t a_lo = sA_lo + (uint32_t)(stage * (G::STAGE_BYTES >> 4));
          const uint32_t w_lo = sW1_lo + (uint32_t)(src * 3 * (G::WBLK >> 4));
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
              for (int kk = 0; kk < G::K16; ++kk)
                ptx::umma_bf16_lohi(d, a_lo + ((j * G::PIX + kk * 32) >> 4), A1_HI,
                                    w_lo + boff + ((j * G::WBLK + kk * 32) >> 4), B1_HI, idesc);
            ptx::umma_commit(&in_empty[stage]);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (ptx::elect_one()) {
          if (i - 1 >= s.c1a) ptx::umma_commit(&c1_full[(uint32_t)(i - 1) % SLOTS]);
          if (i == p.H - 1 && i < s.c1b) ptx::umma_commit(&c1_full[(uint32_t)i % SLOTS]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ================================ UMMA issuer, conv2 ===========================
    constexpr uint32_t A2_HI = ptx::smem_desc_hi(8 * 64, 4), B2_HI = ptx::smem_desc_hi(8 * 64, 4);
    constexpr uint32_t IDESC0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint32_t sW2_lo = ptx::smem_desc_lo(ptx::smem_u32(sW2)), sR_lo = ptx::smem_desc_lo(ptx::smem_u32(sRing));
    uint32_t umask2 = 0;
    uint32_t rg_slot = 0, rg_par = 0;   // next ring row conv2 consumes
    wait_t<false>(&bar_w, 0, 2);
    for (long long k = kfirst; k < p.total_strips; k += kstep) {
      Strip s;
      decode_strip(p, k, s);
      int r2_touched = s.ma;
      for (int j = s.c1a; j < s.c1b; ++j) {
        // ring row j (= conv1 output row j) feeds conv2 rows j-1, j, j+1 (inside [ma, mb))
        const int r_lo = max(s.ma, j - 1), r_hi = min(s.mb - 1, j + 1);
        const int n = r_hi - r_lo + 1;
        for (; r2_touched <= r_hi; ++r2_touched) {
          const uint32_t sl = (uint32_t)r2_touched % SLOTS;
          wait_t<false>(&c2_empty[sl], ((umask2 >> sl) & 1u) ^ 1u, 5, (uint32_t)r2_touched, (uint32_t)j);
          umask2 ^= 1u << sl;
        }
        // The ring row lives in THIS SM's shared memory (a neighbour's edge pixel arrives through DSMEM before its
        // release.cluster arrive is counted), so a CTA-scope wait plus a consumer-side proxy fence orders it before
        // the UMMA reads; an acquire.cluster poll would invalidate L1 (CCTL.IVALL, ~500 cycles) on every try.
        if (p.flags & 1) wait_t<true>(&ring_full[rg_slot], rg_par, 6, (uint32_t)j, rg_slot);
        else wait_t<false>(&ring_full[rg_slot], rg_par, 6, (uint32_t)j, rg_slot);
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        ptx::tc_fence_after();
        const uint32_t idesc = IDESC0 | ((uint32_t)(n * COUT >> 3) << 17);
        const uint32_t boff = (uint32_t)(r_lo - (j - 1)) * ((COUT * 64) >> 4);
        const uint32_t d = tmem_base + TMEM_C2 + ((uint32_t)(j + SLOTS - 1) % SLOTS + (uint32_t)(r_lo - (j - 1))) * COUT;
        const uint32_t a_lo = sR_lo + (uint32_t)(rg_slot * (RING_ROW >> 4));
        if (ptx::elect_one()) {
#pragma unroll
          for (int t = 0; t < 3; ++t)
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
              ptx::umma_bf16_lohi(d, a_lo + ((t * 64 + kk * 32) >> 4), A2_HI,
                                  sW2_lo + boff + ((t * W2BLK + kk * 32) >> 4), B2_HI, idesc);
          ptx::umma_commit(&ring_empty[rg_slot]);
          if (j - 1 >= s.ma) ptx::umma_commit(&c2_full[(uint32_t)(j - 1) % SLOTS]);
          if (j == p.H - 1 && j < s.mb) ptx::umma_commit(&c2_full[(uint32_t)j % SLOTS]);
        }
        __syncwarp();
        if (++rg_slot == (uint32_t)p.RS) { rg_slot = 0; rg_par ^= 1u; }
      }
    }
  } else if (warp == 2) {
    // ================================ relay =======================================
    // a ring row of THIS CTA is free again once conv2 has consumed it: tell the neighbours, whose epilogue
    // threads store their edge pixel of a later row into it
    if (has_left || has_right) {
      const uint32_t left_bar = has_left ? mapa(ptx::smem_u32(&nb_empty[1][0]), rank - 1) : 0u;   // I am its right neighbour
      const uint32_t right_bar = has_right ? mapa(ptx::smem_u32(&nb_empty[0][0]), rank + 1) : 0u;
      uint32_t slot = 0, par = 0;
      for (long long k = kfirst; k < p.total_strips; k += kstep) {
        Strip s;
        decode_strip(p, k, s);
        for (int r = s.c1a; r < s.c1b; ++r) {
          wait_t<false>(&ring_empty[slot], par, 7, (uint32_t)r, slot);
          if (lane == 0) {
            if (has_left) mbar_arrive_remote(left_bar + slot * 8u);
            if (has_right) mbar_arrive_remote(right_bar + slot * 8u);
          }
          __syncwarp();
          if (++slot == (uint32_t)p.RS) { slot = 0; par ^= 1u; }
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ================================ conv1 epilogue: TMEM -> ring ==================
    const int quad = warp & 3, g1 = (warp - 4) >> 2;       // two groups alternate rows: a row's drain -> convert -> ring
                                                            // store -> arrive chain is longer than one row time
    const int m = quad * 32 + lane;
    const int x = x0 + m;
    const bool valid = x < p.W;
    const bool edge_l = (m == 0) && has_left, edge_r = (m == 127) && has_right;
    // swizzled byte offsets of this pixel's four 16-byte chunks inside a ring row (SWIZZLE_64B: chunk ^= (addr >> 7) & 3)
    auto sw = [](uint32_t off) { return off ^ (((off >> 7) & 3u) << 4); };
    const uint32_t own_off = (uint32_t)(m + 1) * 64u;
    const uint32_t ring_s = ptx::smem_u32(sRing);
    const uint32_t rem_base = edge_l ? mapa(ring_s, rank - 1) + 129u * 64u : edge_r ? mapa(ring_s, rank + 1) : 0u;
    const uint32_t rem_full = edge_l ? mapa(ptx::smem_u32(&ring_full[0]), rank - 1)
                                     : edge_r ? mapa(ptx::smem_u32(&ring_full[0]), rank + 1) : 0u;
    uint32_t emask = 0, rslot = 0, rpar = 0;
    for (long long k = kfirst; k < p.total_strips; k += kstep) {
      Strip s;
      decode_strip(p, k, s);
      for (int r = s.c1a; r < s.c1b; ++r) {
        const uint32_t slot = (uint32_t)r % SLOTS, par = (emask >> slot) & 1u;
        emask ^= 1u << slot;
        const uint32_t my_rslot = rslot, my_rpar = rpar;     // ring slot of this row; both groups count every row
        if (++rslot == (uint32_t)p.RS) { rslot = 0; rpar ^= 1u; }
        wait_t<false>(&c1_full[slot], par, 8, (uint32_t)r, slot);      // every group observes every phase
        if ((r & 1) != g1) { __syncwarp(); if (lane == 0) ptx::mbar_arrive(&c1_empty[slot]); continue; }
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(TMEM_C1 + (int)slot * COUT);
        uint32_t v[COUT];
        __syncwarp();
        ptx::tmem_ld16(taddr, v);
        ptx::tmem_ld16(taddr + 16, v + 16);
        ptx::tmem_ld_wait();
        ptx::tmem_st16_fill(taddr, 0u);
        ptx::tmem_st16_fill(taddr + 16, 0u);
        if (slot < 2u) {
          const uint32_t paddr = taddr + SLOTS * COUT;
#pragma unroll
          for (int c0 = 0; c0 < COUT; c0 += 16) {
            uint32_t u[16];
            ptx::tmem_ld16(paddr + c0, u);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[c0 + i] = __float_as_uint(__uint_as_float(v[c0 + i]) + __uint_as_float(u[i]));
            ptx::tmem_st16_fill(paddr + c0, 0u);
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&c1_empty[slot]);

        uint4 w4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = valid ? fmaxf(__uint_as_float(v[c * 8 + i]) + p.bias1[c * 8 + i], 0.f) : 0.f;
          w4[c] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
        // the ring row must have been consumed by conv2 (here, and in the neighbour the edge pixel goes to)
        wait_t<false>(&ring_empty[my_rslot], my_rpar ^ 1u, 9, (uint32_t)r, my_rslot);
        if (quad == 0 && has_left) wait_t<false>(&nb_empty[0][my_rslot], my_rpar ^ 1u, 10, (uint32_t)r, my_rslot);
        if (quad == 3 && has_right) wait_t<false>(&nb_empty[1][my_rslot], my_rpar ^ 1u, 11, (uint32_t)r, my_rslot);
        uint8_t* row = sRing + (size_t)my_rslot * RING_ROW;
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(row + sw(own_off + c * 16u)) = w4[c];
        if (edge_l || edge_r) {
          const uint32_t rb = rem_base + my_rslot * (uint32_t)RING_ROW;
          const uint32_t poff = edge_l ? 129u * 64u : 0u;    // swizzle is a function of the offset inside the (1024-aligned) row
#pragma unroll
          for (int c = 0; c < 4; ++c) st_async_v4(rb - poff + sw(poff + c * 16u), w4[c], rem_full + my_rslot * 8u);
        }
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          // quad 0's arrive also announces the 64 bytes each neighbour's edge thread sends into this ring row
          const uint32_t nb_bytes = (quad == 0) ? 64u * ((has_left ? 1u : 0u) + (has_right ? 1u : 0u)) : 0u;
          if (nb_bytes) ptx::mbar_arrive_expect_tx(&ring_full[my_rslot], nb_bytes);
          else ptx::mbar_arrive(&ring_full[my_rslot]);
        }
      }
    }
  } else if (warp >= 12) {
    // ================================ conv2 epilogue: TMEM -> global (+ pool) =======
    const int quad = warp & 3, eg = (warp - 12) >> 2;
    const int m = quad * 32 + lane;
    const int x = x0 + m;
    const bool valid = x < p.W;
    const bool pool = p.pool_out != nullptr;
    uint32_t q = 0, emask = 0;
    for (long long k = kfirst; k < p.total_strips; k += kstep) {
      Strip s;
      decode_strip(p, k, s);
      uint32_t prow[COUT / 2];
      for (int r = s.ma; r < s.mb; ++r, ++q) {
        const uint32_t slot = (uint32_t)r % SLOTS, par = (emask >> slot) & 1u;
        emask ^= 1u << slot;
        // with the fused pool a group keeps both rows of a pair (strips start on even rows)
        // Every group observes EVERY phase of a slot's barrier, also for the rows the other group drains: a parity
        // wait only tells the current phase from the one before it, and a group that skipped a phase could run two
        // phases ahead across a strip boundary and take the stale parity for "done".
        wait_t<false>(&c2_full[slot], par, 12, (uint32_t)r, slot);
        if (pool ? (((r >> 1) & 1) != eg) : ((int)(q & 1u) != eg)) { __syncwarp(); if (lane == 0) ptx::mbar_arrive(&c2_empty[slot]); continue; }
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(TMEM_C2 + (int)slot * COUT);
        uint32_t v[COUT];
        __syncwarp();
        ptx::tmem_ld16(taddr, v);
        ptx::tmem_ld16(taddr + 16, v + 16);
        ptx::tmem_ld_wait();
        ptx::tmem_st16_fill(taddr, 0u);
        ptx::tmem_st16_fill(taddr + 16, 0u);
        if (slot < 2u) {
          const uint32_t paddr = taddr + SLOTS * COUT;
#pragma unroll
          for (int c0 = 0; c0 < COUT; c0 += 16) {
            uint32_t u[16];
            ptx::tmem_ld16(paddr + c0, u);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[c0 + i] = __float_as_uint(__uint_as_float(v[c0 + i]) + __uint_as_float(u[i]));
            ptx::tmem_st16_fill(paddr + c0, 0u);
          }
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&c2_empty[slot]);

        auto mx = [](uint32_t a, uint32_t b) {
          __nv_bfloat162 r2 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
          return *reinterpret_cast<uint32_t*>(&r2);
        };
        uint4* dst = reinterpret_cast<uint4*>(p.out + (((size_t)s.img * p.H + r) * p.W + x) * COUT);
        uint4 wprev = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int c0 = 0; c0 < COUT; c0 += 8) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaxf(__uint_as_float(v[c0 + i]) + p.bias2[c0 + i], 0.f);
          uint4 w;
          w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]);
          w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
          if (c0 & 8) { if (valid) ptx::st_global_256(dst + c0 / 8 - 1, wprev, w); }
          else wprev = w;
          if (pool) {   // max with the even row of the pair (post-ReLU values: a missing pixel counts as 0)
            if (!valid) w = make_uint4(0u, 0u, 0u, 0u);
            if (r & 1) { w.x = mx(w.x, prow[c0 / 2]); w.y = mx(w.y, prow[c0 / 2 + 1]); w.z = mx(w.z, prow[c0 / 2 + 2]); w.w = mx(w.w, prow[c0 / 2 + 3]); }
            prow[c0 / 2] = w.x; prow[c0 / 2 + 1] = w.y; prow[c0 / 2 + 2] = w.z; prow[c0 / 2 + 3] = w.w;
          }
        }
        if (pool && ((r & 1) || r == p.H - 1)) {
          const int Hp = (p.H + 1) >> 1, Wp = (p.W + 1) >> 1;
          uint4* pd = reinterpret_cast<uint4*>(p.pool_out + (((size_t)s.img * Hp + (r >> 1)) * Wp + (x >> 1)) * COUT);
#pragma unroll
          for (int c4 = 0; c4 < COUT / 8; ++c4) {
            uint4 o;
            o.x = mx(prow[c4 * 4], __shfl_xor_sync(0xffffffffu, prow[c4 * 4], 1));
            o.y = mx(prow[c4 * 4 + 1], __shfl_xor_sync(0xffffffffu, prow[c4 * 4 + 1], 1));
            o.z = mx(prow[c4 * 4 + 2], __shfl_xor_sync(0xffffffffu, prow[c4 * 4 + 2], 1));
            o.w = mx(prow[c4 * 4 + 3], __shfl_xor_sync(0xffffffffu, prow[c4 * 4 + 3], 1));
            if (valid && !(m & 1)) pd[c4] = o;
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();           // no CTA leaves while a neighbour may still store into its ring or arrive on its barriers
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

size_t block_smem(int KC1, int nsrc, int stages, int RS) {
  const size_t pix = (size_t)KC1 * 2, wblk = 3 * COUT * pix;
  const size_t stage = align_up(130 * pix, 1024);
  return 1024 + align_up((size_t)nsrc * 3 * wblk, 1024) + align_up((size_t)3 * W2BLK, 1024) + (size_t)RS * RING_ROW + (size_t)stages * stage;
}

template <int KC1>
int launch_block(const BlockLaunch& L, cudaStream_t stream) {
  using G = G1<KC1>;
  auto kern = conv_block_kernel<KC1>;
  BlockParams p;
  memset(&p, 0, sizeof(p));
  p.nsrc = L.nsrc; p.NIMG = L.NIMG; p.H = L.H; p.W = L.W;
  p.NC = ceil_div(L.W, 128);
  auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e && *e ? atoi(e) : dflt; };
  p.lag = 0;
  p.RS = std::min(MAX_RS, std::max(4, env_int("CETPICK_BLOCK_RS", 8)));        // ring rows between the two convolutions
  p.flags = env_int("CETPICK_BLOCK_FLAGS", 0);
  const int max_stages = std::min(MAX_STAGES, std::max(3, env_int("CETPICK_BLOCK_STAGES", MAX_STAGES)));
  static DeviceOnce attr_once;
  static int static_smem = 0;
  if (attr_once.first()) {
    cudaFuncAttributes fa;
    CETPICK_CUDA(cudaFuncGetAttributes(&fa, kern));
    CETPICK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes));
    CETPICK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));
    static_smem = (int)fa.sharedSizeBytes;
  }
  const size_t avail = (size_t)227 * 1024 - static_smem;
  p.stages = max_stages;
  while (p.stages > 3 && block_smem(KC1, L.nsrc, p.stages, p.RS) > avail) --p.stages;
  if (block_smem(KC1, L.nsrc, p.stages, p.RS) > avail) return CETPICK_ERR_UNSUPPORTED;
  const size_t smem = block_smem(KC1, L.nsrc, p.stages, p.RS);
  for (int c = 0; c < COUT; ++c) { p.bias1[c] = L.bias1_host[c]; p.bias2[c] = L.bias2_host[c]; }
  p.out = static_cast<__nv_bfloat16*>(L.out);
  p.pool_out = static_cast<__nv_bfloat16*>(L.pool_out);

  // clusters that can be resident at once (a cluster = the CTAs of one image row; all on one GPC)
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)p.NC, 1, 1); cfg.blockDim = dim3(BLK_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem; cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = 1;
  int max_clusters = 0;
  CETPICK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
  if (max_clusters < 1) return CETPICK_ERR_UNSUPPORTED;

  // strips along y: enough strips to balance the resident clusters, long enough that the recomputed rows stay cheap
  int best_n = 1, best_R = L.H + (L.pool_out ? (L.H & 1) : 0);
  double best_eff = -1.0;
  for (int n = 1; n <= std::max(1, L.H / 16); ++n) {
    int R = ceil_div(L.H, n);
    if (L.pool_out) R += R & 1;               // row pairs of the fused pool stay inside one strip
    const int nn = ceil_div(L.H, R);
    const long long strips = (long long)L.NIMG * nn;
    const double waves = (double)ceil_div<long long>(strips, max_clusters);
    const double eff = ((double)strips / (waves * max_clusters)) * ((double)R / (R + 6));
    if (eff > best_eff + 1e-9) { best_eff = eff; best_n = nn; best_R = R; }
  }
  p.nchunk = best_n; p.R = best_R;
  p.total_strips = (long long)L.NIMG * p.nchunk;

  int rc;
  for (int s = 0; s < L.nsrc; ++s) {
    const uint64_t C = (uint64_t)KC1;
    const uint64_t dims[4] = {C, (uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.NIMG};
    const uint64_t strides[3] = {C * 2, C * 2 * L.W, C * 2 * (uint64_t)L.W * L.H};
    const uint32_t box[4] = {(uint32_t)KC1, 130, 1, 1};
    if ((rc = tmap_encode_bf16(&p.tmA[s], L.src[s], 4, dims, strides, box, KC1))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)KC1, (uint64_t)L.nsrc * 3 * 3 * COUT};
    const uint64_t strides[1] = {(uint64_t)G::PIX};
    const uint32_t box[2] = {(uint32_t)KC1, (uint32_t)(3 * COUT)};
    if ((rc = tmap_encode_bf16(&p.tmB1, L.w1pk, 2, dims, strides, box, KC1))) return rc;
  }
  {
    const uint64_t dims[2] = {32, (uint64_t)3 * 3 * COUT};
    const uint64_t strides[1] = {64};
    const uint32_t box[2] = {32, (uint32_t)(3 * COUT)};
    if ((rc = tmap_encode_bf16(&p.tmB2, L.w2pk, 2, dims, strides, box, 32))) return rc;
  }
  const int nclusters = (int)std::min<long long>(p.total_strips, max_clusters);
  cfg.gridDim = dim3((unsigned)(nclusters * p.NC), 1, 1);
  CETPICK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

}  // namespace

bool block_supported(int C1, int nsrc, int W) {
  if (!((C1 == 16 && nsrc == 1) || (C1 == 32 && (nsrc == 1 || nsrc == 2)))) return false;
  return W >= 1 && W <= 128 * 8;        // one cluster (<= 8 CTAs of 128 pixels) spans a row
}

int conv_block_launch(const BlockLaunch& L, cudaStream_t stream) {
  if (!block_supported(L.C1, L.nsrc, L.W)) return CETPICK_ERR_UNSUPPORTED;
  if (!L.src[0] || (L.nsrc > 1 && !L.src[1]) || !L.w1pk || !L.w2pk || !L.bias1_host || !L.bias2_host || !L.out ||
      L.NIMG <= 0 || L.H <= 0)
    return CETPICK_ERR_BAD_ARG;
  return L.C1 == 16 ? launch_block<16>(L, stream) : launch_block<32>(L, stream);
}

}  // namespace cetpick

using namespace cetpick;

#ifdef CETPICK_TEST_HOOKS   // test / tuning hooks: built into libcetpick_test_sm100a.so only (include/cetpick_test.h)
// Test hook: one fused block from PyTorch-layout fp32 host weights (packs, uploads, launches, synchronises, frees).
extern "C" int cetpick_conv_block_bf16(int nsrc, const void* src0, const void* src1, int C1, int NIMG, int H, int W,
                                       const float* w1_host, const float* bias1_host, const float* w2_host,
                                       const float* bias2_host, void* out, void* pool_out, void* stream) {
  g_launches = 0;
  if (!w1_host || !w2_host || !bias1_host || !bias2_host) return CETPICK_ERR_BAD_ARG;
  if (!block_supported(C1, nsrc, W)) return CETPICK_ERR_UNSUPPORTED;
  std::vector<uint16_t> pk1 = march_pack_weights(MARCH_2D_ROWS, w1_host, COUT, nsrc, C1, nullptr);
  std::vector<uint16_t> pk2 = march_pack_weights(MARCH_2D_ROWS, w2_host, COUT, 1, 32, nullptr);
  void *d1 = nullptr, *d2 = nullptr;
  CETPICK_CUDA(cudaMalloc(&d1, pk1.size() * 2));
  if (cudaMalloc(&d2, pk2.size() * 2) != cudaSuccess) { cudaFree(d1); return CETPICK_ERR_CUDA; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = CETPICK_OK;
  if (cudaMemcpyAsync(d1, pk1.data(), pk1.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemcpyAsync(d2, pk2.data(), pk2.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess)
    rc = CETPICK_ERR_CUDA;
  if (rc == CETPICK_OK) {
    BlockLaunch L;
    L.nsrc = nsrc; L.src[0] = src0; L.src[1] = src1; L.C1 = C1; L.NIMG = NIMG; L.H = H; L.W = W;
    L.w1pk = d1; L.w2pk = d2; L.bias1_host = bias1_host; L.bias2_host = bias2_host; L.out = out; L.pool_out = pool_out;
    rc = conv_block_launch(L, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(d1);
  cudaFree(d2);
  if (rc == CETPICK_OK && e != cudaSuccess) return cuda_fail(e, "conv_block");
  return rc;
}

// Bring-up aid: registers (first call) and returns a 256-word host-mapped buffer; word 0 counts the waits that timed
// out in conv_block_kernel, entries of 4 words follow from word 4: (cta << 8 | warp, wait-site tag, value, value).
extern "C" int cetpick_block_debug_buffer(uint32_t** host_buf) {
  static uint32_t* h = nullptr;
  if (!h) {
    CETPICK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h), 1024, cudaHostAllocMapped));
    memset(h, 0, 1024);
    uint32_t* d = nullptr;
    CETPICK_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0));
    CETPICK_CUDA(cudaMemcpyToSymbol(g_blk_dbg, &d, sizeof(d)));
  }
  if (host_buf) *host_buf = h;
  return CETPICK_OK;
}

#endif  // CETPICK_TEST_HOOKS
