// "Marching" convolution on tcgen05 tensor cores (sm_100a) for the narrow layers of the detector
// (Cout = 32 / 64; 2/3 of the network's FLOPs): cet_pick/models/networks/unet.py:127-145 (3x3
// Conv2d of DownConv / UpConv) and unet_small.py:39-46 (feature_head Conv3d, dilation (1,4,4)).
//
// Why not a plain implicit GEMM: with N = Cout = 32 every UMMA re-reads its 4 KB A operand from
// shared memory for 16 cycles of tensor work (the 128 B/clk shared-memory port allows ~40 % of
// peak), and every tap re-fetches the activation tile from L2.  Here the kernel marches along one
// axis (y for the 2-D convs, z for the 3-D ones).  The three taps along that axis are stacked on
// the GEMM N dimension:
//     input row i contributes to output rows i-1, i, i+1 with weights W[+1], W[0], W[-1]
//  => ONE UMMA  D[128 px, 3*Cout] += A[128 px, 16 ch] * [W[+1]; W[0]; W[-1]]^T        (N = 96 / 192)
// whose three column blocks are the accumulators of three different output rows, kept in a ring
// of TMEM slots.  Each input row (2-D: 1 x (128+2) pixels; 3-D: a (16+2d) x (8*MT+2d) plane tile)
// is loaded by TMA exactly once per strip, the in-step taps (dx, or (dy,dx)) are shifted UMMA
// descriptors into that one tile, and all weights stay resident in shared memory.  An accumulator
// slot is complete when the row after it has been consumed; the epilogue warps drain it
// (bias, ReLU, bf16, NHWC store), write zeros back and hand the slot to the MMA thread again, so
// every UMMA accumulates and no instruction needs a per-column-block "first touch" flag.
// The ring has SL logical slots and two more physical ones behind them: the window of a step starts at
// logical slot w = q mod SL and always spans the physical columns [w, w+3), so it never wraps and a step is
// ONE UMMA per (tap, k) -- an N <= 96 UMMA costs the same ~64 cycles whatever N is (operand-fetch bound), and
// splitting the window at the ring end would add 25 % to the issue count.  Rows whose home slot is 0 or 1
// collect part of their sum in the phantom slots SL / SL+1; the epilogue adds the two.  The home slot of a
// row is (ABSOLUTE row index) mod SL, so the order in which a row's partial sums are added depends on nothing
// but the row's position in the image / volume: strips, slabs and z-shards all give bit-identical results.
//
// Roles (384 threads, 1 CTA/SM, persistent over strips):
//   warp 0: TMA producer   warp 1: UMMA issuer   warp 2: TMEM allocator   warps 4-11: epilogue
#include "conv_march.cuh"
#include "common.cuh"
#include "conv_tc.cuh"
#include "ptx.cuh"

#include <cuda_bf16.h>
#include <algorithm>

namespace cetpick {

namespace {

constexpr int MARCH_THREADS = 384;
constexpr int MAX_STAGES = 8;
constexpr int MAX_SLOTS = 16;
constexpr int TMEM_COLS = 512;
constexpr int DIL3D = 4;         // feature_head dilation (unet_small.py:39-46)

struct alignas(64) MarchParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB;
  int nsrc, chunks, S, SL;   // S physical TMEM slots per M-tile, SL = S - 2 logical ring slots
  int row_origin;            // 3-D: absolute z of plane 0 of this call (z-slab streaming), taken mod SL
  int NIMG, H, W;
  int L, R, nchunk;        // march extent, rows per strip, strips along the march axis
  int nxb;                 // tile columns (2-D: x blocks per row; 3-D: x tiles per plane)
  long long total_strips;
  int stages;
  int nblk;                // weight blocks (source, chunk, tap)
  int relu;
  const float* bias;
  __nv_bfloat16* out;
  __nv_bfloat16* pool_out; // 2-D: fused MaxPool2d(2, ceil) output or null
  // epilogue constants live in the kernel-parameter constant bank: FADD/FFMA read them as c[0][imm]
  // operands, so the epilogue puts no load on the shared-memory pipe the UMMA operands stream through
  float bias_c[64];        // 2-D bias / 3-D bias of the interior tap-validity class (row 63 of bias_tab)
  float hmw_c[96];         // fused hm head weights [3][32]
  const float* bias_tab;   // 3-D: [64][COUT] tap-validity dependent bias (or null)
  const float* hm_w;       // 3-D: fused hm head weights [3][COUT] (or null)
  float* hm_out;
  int hm_sigmoid;
};

// Compile-time geometry of one instantiation.
template <int COUT, int KC, int MODE, int MT>
struct Geo {
  static constexpr int T = MODE == MARCH_2D_ROWS ? 3 : 9;            // in-step taps
  static constexpr int K16 = KC / 16;
  static constexpr int PIX = KC * 2;                                 // bytes of one pixel's chunk = swizzle span
  // TMA boxes of one stage: 2-D = one (128+2)-pixel row piece per M-tile (box extents are capped at
  // 256); 3-D = one (16+2d) x (8*MT+2d) plane tile shared by the M-tiles.
  static constexpr int NBOX = MODE == MARCH_2D_ROWS ? MT : 1;
  static constexpr int BX = MODE == MARCH_2D_ROWS ? 130 : 8 * MT + 2 * DIL3D;
  static constexpr int BY = MODE == MARCH_2D_ROWS ? 1 : 16 + 2 * DIL3D;
  static constexpr int BOX_BYTES = BX * BY * PIX;
  static constexpr int BOX_STRIDE = (BOX_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = NBOX * BOX_STRIDE;
  static constexpr int WBLK = 3 * COUT * PIX;                        // bytes of one weight block
  static constexpr int SLOT_BYTES = COUT * PIX;                      // B rows of one output row
  static constexpr int SBO_A = MODE == MARCH_2D_ROWS ? 8 * PIX : BX * PIX;
  static constexpr int SBO_B = 8 * PIX;
  static constexpr uint32_t LAYOUT = KC == 64 ? 2 : KC == 32 ? 4 : 6;
  // byte offset of the A operand of (M-tile t, in-step tap j, k16 step kk) inside a stage
  static __host__ __device__ constexpr int aoff(int t, int j, int kk) {
    return (MODE == MARCH_2D_ROWS ? t * BOX_STRIDE + j * PIX
                                  : (((j / 3) * DIL3D) * BX + (j % 3) * DIL3D + t * 8) * PIX) + kk * 32;
  }
};

struct Strip { int ma, mb, ha, hb, x0, y0, img; };   // rows computed [ma,mb); hm rows written [ha,hb)

template <int MODE, int MT>
__device__ __forceinline__ void decode_strip(const MarchParams& p, long long k, Strip& s) {
  const int ch = (int)(k % p.nchunk);
  k /= p.nchunk;
  s.ma = ch * p.R;
  s.mb = min(s.ma + p.R, p.L);
  s.ha = s.ma; s.hb = s.mb;
  if (MODE == MARCH_3D_PLANES && p.hm_out) {   // fused hm head needs the feature rows around its own
    s.ma = max(s.ma - 1, 0);
    s.mb = min(s.mb + 1, p.L);
  }
  const int bx = (int)(k % p.nxb);
  k /= p.nxb;
  if (MODE == MARCH_2D_ROWS) { s.x0 = bx * 128 * MT; s.y0 = 0; s.img = (int)k; }
  else { s.x0 = bx * 8 * MT; s.y0 = (int)k * 16; s.img = 0; }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int COUT, int KC, int MODE, int MT>
__global__ void __launch_bounds__(MARCH_THREADS, 1) conv_march_kernel(const __grid_constant__ MarchParams p) {
  using G = Geo<COUT, KC, MODE, MT>;
  // COUT = 32: N = 96 UMMAs are operand-fetch bound, a window split at the ring end would double the issue count of
  // that step -> phantom slots (SL = S - 2).  COUT = 64: N = 192 UMMAs run at the tensor rate, a split costs nothing
  // in time (two UMMAs of N = 64 + 128) -> plain ring (SL = S) with the window split where it wraps.
  constexpr bool PHANTOM = COUT == 32;
  // accumulator slots per M-tile and logical ring length: compile-time, so the slot arithmetic below is add / compare /
  // multiply-shift instead of a runtime division (the issuer warp's scalar chain is what bounds the 32-channel layers)
  constexpr int S_CT = (TMEM_COLS / (MT * COUT)) < MAX_SLOTS ? (TMEM_COLS / (MT * COUT)) : MAX_SLOTS;
  constexpr uint32_t SL = PHANTOM ? S_CT - 2 : S_CT;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[MAX_STAGES], bar_empty[MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_afull[MAX_SLOTS], bar_aempty[MAX_SLOTS], bar_w;
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_btab[MODE == MARCH_3D_PLANES ? 64 * COUT : 4];   // border classes only

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;
  uint8_t* sA = smem + (((size_t)p.nblk * G::WBLK + 1023) & ~(size_t)1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tmA[0]);
    if (p.nsrc > 1) ptx::prefetch_tensormap(&p.tmA[1]);
    ptx::prefetch_tensormap(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(&bar_full[s], 1); ptx::mbar_init(&bar_empty[s], 1); }
    // a slot is free again when BOTH epilogue groups are through with its row: MT = 2: each group drains its own M-tile;
    // MT = 1: one group drains, the other only observes the phase -- and must arrive too, or the issuer could start the
    // slot's next phase (and the one after) before the observer has looked, which a parity wait cannot tell apart
    for (int s = 0; s < (int)SL; ++s) { ptx::mbar_init(&bar_afull[s], 1); ptx::mbar_init(&bar_aempty[s], 8); }
    ptx::mbar_init(&bar_w, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(&s_tmem_base, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (warp == 3 && MODE == MARCH_3D_PLANES)
    for (int c = lane; c < 64 * COUT; c += 32) s_btab[c] = p.bias_tab ? p.bias_tab[c] : p.bias_c[c % COUT];
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp >= 4) {   // all accumulator slots start at zero: every UMMA accumulates
    const uint32_t row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp - 4) >> 2) * 256);
    for (int c = 0; c < 256; c += 16) ptx::tmem_st16_fill(row + c, 0u);
    ptx::tmem_st_wait();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();

  // barrier addresses in the shared window, computed once
  uint32_t a_full = ptx::smem_u32(bar_full), a_empty = ptx::smem_u32(bar_empty);
  uint32_t a_afull = ptx::smem_u32(bar_afull), a_aempty = ptx::smem_u32(bar_aempty);
  // (opaque to the compiler from here on: otherwise it re-derives each address -- S2UR SR_CgaCtaId + ULEA -- at every use)
  asm volatile("" : "+r"(a_full), "+r"(a_empty), "+r"(a_afull), "+r"(a_aempty));

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(&bar_w, (uint32_t)(p.nblk * G::WBLK));
      for (int b = 0; b < p.nblk; ++b)
        ptx::tma_load_2d(sW + (size_t)b * G::WBLK, &p.tmB, &bar_w, 0, b * 3 * COUT);
      int stage = 0;
      uint32_t phase = 0;
      for (long long k = blockIdx.x; k < p.total_strips; k += gridDim.x) {
        Strip s;
        decode_strip<MODE, MT>(p, k, s);
        const int i_lo = max(s.ma - 1, 0), i_hi = min(s.mb, p.L - 1);
        for (int i = i_lo; i <= i_hi; ++i)
          for (int src = 0; src < p.nsrc; ++src)
            for (int c = 0; c < p.chunks; ++c) {
              ptx::mbar_wait_a(a_empty + 8u * (uint32_t)stage, phase ^ 1u);
              ptx::mbar_arrive_expect_tx_a(a_full + 8u * (uint32_t)stage, (uint32_t)(G::NBOX * G::BOX_BYTES));
              uint8_t* dst = sA + (size_t)stage * G::STAGE_BYTES;
              if (MODE == MARCH_2D_ROWS)
                for (int t = 0; t < MT; ++t)
                  ptx::tma_load_4d(dst + t * G::BOX_STRIDE, &p.tmA[src], &bar_full[stage], c * KC, s.x0 + t * 128 - 1, i, s.img);
              else
                ptx::tma_load_4d(dst, &p.tmA[src], &bar_full[stage], c * KC, s.x0 - DIL3D, s.y0 - DIL3D, i);
              if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
      }
    }
  } else if (warp == 1) {
    // ================================ UMMA issuer =================================
    // The whole warp runs this loop on warp-uniform values (so descriptors live in uniform
    // registers); one elected lane issues the tcgen05 instructions.
    constexpr uint32_t A_HI = ptx::smem_desc_hi(G::SBO_A, G::LAYOUT), B_HI = ptx::smem_desc_hi(G::SBO_B, G::LAYOUT);
    constexpr uint32_t IDESC0 = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 4) << 24);
    const uint32_t sW_lo = ptx::smem_desc_lo(ptx::smem_u32(sW)), sA_lo = ptx::smem_desc_lo(ptx::smem_u32(sA));
    constexpr uint32_t tile_cols = (uint32_t)(S_CT * COUT);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t umask = 0;      // bit s: parity of the number of rows that have used logical slot s
    const uint32_t org = (uint32_t)p.row_origin;
    ptx::mbar_wait(&bar_w, 0);
    for (long long k = blockIdx.x; k < p.total_strips; k += gridDim.x) {
      Strip s;
      decode_strip<MODE, MT>(p, k, s);
      const int i_lo = max(s.ma - 1, 0), i_hi = min(s.mb, p.L - 1);
      int r_touched = s.ma;
      // running slots (no division in the row loop): sl_t = slot of row r_touched, sl_p = slot of row i - 1
      uint32_t sl_t = ((uint32_t)s.ma + org) % SL;
      uint32_t sl_p = ((uint32_t)i_lo + org + SL - 1u) % SL;
      for (int i = i_lo; i <= i_hi; ++i) {
        const int r_lo = max(s.ma, i - 1), r_hi = min(s.mb - 1, i + 1);
        const int n = r_hi - r_lo + 1;
        for (; r_touched <= r_hi; ++r_touched) {    // rows touched for the first time: slot must be drained
          ptx::mbar_wait_a(a_aempty + 8u * sl_t, ((umask >> sl_t) & 1u) ^ 1u);
          umask ^= 1u << sl_t;
          sl_t = (sl_t + 1u == SL) ? 0u : sl_t + 1u;
        }
        // the window starts at the home slot of row i-1 and runs on into the phantom slots: no wrap, and a
        // row's partial sums land in the same places wherever the strip starts
        uint32_t s_lo = sl_p + (uint32_t)(r_lo - (i - 1));
        if (!PHANTOM && s_lo >= SL) s_lo -= SL;
        const int n1 = PHANTOM ? n : min(n, (int)(SL - s_lo)), n2 = n - n1;     // plain ring: the wrap splits the columns
        const uint32_t id1 = IDESC0 | ((uint32_t)(n1 * COUT >> 3) << 17), id2 = IDESC0 | ((uint32_t)(n2 * COUT >> 3) << 17);
        const uint32_t boff1 = (uint32_t)(r_lo - (i - 1)) * (G::SLOT_BYTES >> 4);
        const uint32_t boff2 = boff1 + (uint32_t)n1 * (G::SLOT_BYTES >> 4);
        const uint32_t d1 = tmem_base + s_lo * COUT, d2 = tmem_base;
        const uint32_t sl_i = (sl_p + 1u == SL) ? 0u : sl_p + 1u;     // slot of row i
        for (int src = 0; src < p.nsrc; ++src)
          for (int c = 0; c < p.chunks; ++c) {
            ptx::mbar_wait_a(a_full + 8u * (uint32_t)stage, phase);
            ptx::tc_fence_after();
            const uint32_t a_lo = sA_lo + (uint32_t)(stage * (G::STAGE_BYTES >> 4));
            const uint32_t w_lo = sW_lo + (uint32_t)((src * p.chunks + c) * G::T * (G::WBLK >> 4));
            if (ptx::elect_one()) {
#pragma unroll
              for (int t = 0; t < MT; ++t)
#pragma unroll
                for (int j = 0; j < G::T; ++j)
#pragma unroll
                  for (int kk = 0; kk < G::K16; ++kk)
                  {
                    ptx::umma_bf16_lohi(d1 + t * tile_cols, a_lo + (G::aoff(t, j, kk) >> 4), A_HI,
                                        w_lo + boff1 + ((j * G::WBLK + kk * 32) >> 4), B_HI, id1);
                    if (!PHANTOM && n2)
                      ptx::umma_bf16_lohi(d2 + t * tile_cols, a_lo + (G::aoff(t, j, kk) >> 4), A_HI,
                                          w_lo + boff2 + ((j * G::WBLK + kk * 32) >> 4), B_HI, id2);
                  }
              ptx::umma_commit_a(a_empty + 8u * (uint32_t)stage);
              if (src == p.nsrc - 1 && c == p.chunks - 1) {
                // last operand block of input row i: the rows that have now seen all three of their input rows
                // (same election as the UMMAs: one ELECT / reconvergence per row instead of two)
                if (i - 1 >= s.ma) ptx::umma_commit_a(a_afull + 8u * sl_p);
                if (i == p.L - 1 && i < s.mb) ptx::umma_commit_a(a_afull + 8u * sl_i);
              }
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        sl_p = sl_i;
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int quad = warp & 3, eg = (warp - 4) >> 2;
    const int m = quad * 32 + lane;
    const bool fuse_hm = MODE == MARCH_3D_PLANES && p.hm_out != nullptr;
    uint32_t q0 = 0, emask = 0;
    const uint32_t org = (uint32_t)p.row_origin;
    for (long long k = blockIdx.x; k < p.total_strips; k += gridDim.x) {
      Strip s;
      decode_strip<MODE, MT>(p, k, s);
      float hmA = 0.f, hmB = 0.f;     // fused hm head: partial sums of hm[r-1] and hm[r] (this thread's pixel)
      const bool pool = MODE == MARCH_2D_ROWS && p.pool_out != nullptr;
      uint32_t prow[MODE == MARCH_2D_ROWS ? COUT / 2 : 1];   // fused pool: the even row of the pair (packed bf16)
      for (int r = s.ma; r < s.mb; ++r) {
        const uint32_t q = q0 + (uint32_t)(r - s.ma);       // running row count: only spreads rows over the two groups
        const uint32_t slot = ((uint32_t)r + org) % SL, par = (emask >> slot) & 1u;
        emask ^= 1u << slot;
#pragma unroll
        for (int t = 0; t < MT; ++t) {
          // one epilogue group per (row, M-tile); with the fused pool a group keeps both rows of a pair
          // (MT = 2: group = M-tile; MT = 1: groups alternate row PAIRS; strips start on even rows)
          // A warp waits for a phase of a slot's barrier EXACTLY ONCE and arrives on the slot-free barrier afterwards:
          //  * it must observe every phase (a parity wait only tells the current phase from the one before it; a group
          //    that skipped a phase could run two phases ahead across a strip boundary and take the stale parity for
          //    "done") -- MT = 1: the group that does not drain a row still waits, then arrives as an observer;
          //  * it must never wait for a phase AGAIN after its arrival: the issuer may then reuse the slot, and a warp that
          //    is held up for two phases between its arrival and a second wait takes the parity of the phase two ahead
          //    for the one it wants and blocks on a phase that needs its own next arrival -- a dead-lock seen about once
          //    in ten slab-mode forwards when MT = 2 groups also "observed" the other group's tile (found with a wait
          //    recorder).  MT = 2: each group drains its own tile of EVERY row, so it skips the
          //    other tile without waiting.
          if (MT == 2 && t != eg) continue;
          ptx::mbar_wait_a(a_afull + 8u * slot, par);
          if ((pool && MT == 1) ? (((r >> 1) & 1) != eg) : ((int)((q * (uint32_t)MT + (uint32_t)t) & 1u) != eg)) {
            if (MT == 1) { __syncwarp(); if (lane == 0) ptx::mbar_arrive_a(a_aempty + 8u * slot); }   // observed: see bar_aempty's init
            continue;
          }
          ptx::tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((t * S_CT + (int)slot) * COUT);
          uint32_t v[COUT];
          __syncwarp();
#pragma unroll
          for (int c0 = 0; c0 < COUT; c0 += 16) ptx::tmem_ld16(taddr + c0, v + c0);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c0 = 0; c0 < COUT; c0 += 16) ptx::tmem_st16_fill(taddr + c0, 0u);
          if (PHANTOM && slot < 2u) {   // home slots 0 / 1: the other part of the sum sits in the phantom slots SL / SL+1
            const uint32_t paddr = taddr + SL * COUT;
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
          if (lane == 0) ptx::mbar_arrive_a(a_aempty + 8u * slot);

          int x, y, z;
          if (MODE == MARCH_2D_ROWS) { x = s.x0 + t * 128 + m; y = r; z = s.img; }
          else { x = s.x0 + t * 8 + (m & 7); y = s.y0 + (m >> 3); z = r; }
          const bool valid = x < p.W && y < p.H;
          int cls = 63;                       // 3-D: which taps are inside the volume (63 = all of them)
          if (MODE == MARCH_3D_PLANES) {
            const int cz = (z >= 1) + 2 * (z + 1 < p.NIMG);
            const int cy = (y >= DIL3D) + 2 * (y + DIL3D < p.H);
            const int cx = (x >= DIL3D) + 2 * (x + DIL3D < p.W);
            cls = (cz * 4 + cy) * 4 + cx;
          }
          const float* brow = s_btab + cls * COUT;
          float d0 = 0.f, d1 = 0.f, d2 = 0.f;
          uint4* dst = reinterpret_cast<uint4*>(p.out + (((size_t)z * p.H + y) * p.W + x) * COUT);
          uint4 wprev = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int c0 = 0; c0 < COUT; c0 += 8) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float bv = (MODE == MARCH_3D_PLANES && cls != 63) ? brow[c0 + i] : p.bias_c[c0 + i];
              f[i] = __uint_as_float(v[c0 + i]) + bv;
              if (p.relu) f[i] = fmaxf(f[i], 0.f);
            }
            if (fuse_hm) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                d0 = fmaf(f[i], p.hmw_c[(c0 + i) % 32], d0);
                d1 = fmaf(f[i], p.hmw_c[32 + (c0 + i) % 32], d1);
                d2 = fmaf(f[i], p.hmw_c[64 + (c0 + i) % 32], d2);
              }
            }
            uint4 w;
            w.x = pack_bf16x2(f[0], f[1]); w.y = pack_bf16x2(f[2], f[3]);
            w.z = pack_bf16x2(f[4], f[5]); w.w = pack_bf16x2(f[6], f[7]);
            if (c0 & 8) { if (valid && p.out) ptx::st_global_256(dst + c0 / 8 - 1, wprev, w); }
            else wprev = w;
            if (MODE == MARCH_2D_ROWS && pool) {
              // max with the even row of the pair (values are post-ReLU, so a missing pixel counts as 0)
              auto mx = [](uint32_t a, uint32_t b) {
                __nv_bfloat162 r2 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
                return *reinterpret_cast<uint32_t*>(&r2);
              };
              if (!valid) w = make_uint4(0u, 0u, 0u, 0u);
              if (r & 1) { w.x = mx(w.x, prow[c0 / 2]); w.y = mx(w.y, prow[c0 / 2 + 1]); w.z = mx(w.z, prow[c0 / 2 + 2]); w.w = mx(w.w, prow[c0 / 2 + 3]); }
              prow[c0 / 2] = w.x; prow[c0 / 2 + 1] = w.y; prow[c0 / 2 + 2] = w.z; prow[c0 / 2 + 3] = w.w;
            }
          }
          if (MODE == MARCH_2D_ROWS && pool && ((r & 1) || r == p.L - 1)) {
            // prow = max over the row pair; now the x pair (lanes m, m^1) and the even lane stores
            auto mx = [](uint32_t a, uint32_t b) {
              __nv_bfloat162 r2 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
              return *reinterpret_cast<uint32_t*>(&r2);
            };
            const int Hp = (p.H + 1) >> 1, Wp = (p.W + 1) >> 1;
            uint4* pd = reinterpret_cast<uint4*>(p.pool_out + (((size_t)z * Hp + (r >> 1)) * Wp + (x >> 1)) * COUT);
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
          if (fuse_hm) {
            // hm[z] = w0.f[z-1] + w1.f[z] + w2.f[z+1]; rows outside the volume contribute zero
            auto emit = [&](int zz, float val) {
              if (valid && zz >= s.ha && zz < s.hb) {
                if (p.hm_sigmoid) val = fminf(fmaxf(1.0f / (1.0f + expf(-val)), 1e-4f), 1.0f - 1e-4f);
                p.hm_out[((size_t)zz * p.H + y) * p.W + x] = val;
              }
            };
            emit(r - 1, hmA + d2);
            hmA = hmB + d1;
            hmB = d0;
            if (r == p.L - 1) emit(r, hmA);
          }
        }
      }
      q0 += (uint32_t)(s.mb - s.ma);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

int smem_limit() { return 227 * 1024; }

}  // namespace

bool march_supported(int mode, int C, int nsrc, int Cout) {
  if (Cout != 32 && Cout != 64) return false;
  if (C != 16 && C != 32 && C != 64) return false;
  if (Cout == 64 && C == 16) return false;
  if (nsrc < 1 || nsrc > 2) return false;
  if (mode == MARCH_3D_PLANES && (C != 32 || nsrc != 1 || Cout != 32)) return false;
  const int T = mode == MARCH_2D_ROWS ? 3 : 9;
  const size_t wbytes = align_up((size_t)nsrc * T * 3 * Cout * C * 2, 1024);
  const size_t stage = mode == MARCH_2D_ROWS ? (Cout == 32 ? 2 : 1) * align_up((size_t)130 * std::min(C, 64) * 2, 1024)
                                             : (size_t)24 * 24 * 64;
  return wbytes + 3 * stage + 4096 <= (size_t)smem_limit();   // weights + >= 3 activation stages
}

std::vector<uint16_t> march_pack_weights(int mode, const float* w, int Cout, int nsrc, int C, const double* scale) {
  const int KC = std::min(C, 64), chunks = C / KC, T = mode == MARCH_2D_ROWS ? 3 : 9;
  const int Cin = nsrc * C, ktot = 3 * T;     // kernel volume: 9 (ky,kx) or 27 (kz,ky,kx)
  std::vector<uint16_t> out((size_t)nsrc * chunks * T * 3 * Cout * KC);
  for (int s = 0; s < nsrc; ++s)
    for (int c = 0; c < chunks; ++c)
      for (int j = 0; j < T; ++j) {
        const size_t b = ((size_t)s * chunks + c) * T + j;
        for (int slot = 0; slot < 3; ++slot)
          for (int co = 0; co < Cout; ++co)
            for (int k = 0; k < KC; ++k) {
              const int ci = s * C + c * KC + k;
              const int km = 2 - slot;                // kernel index along the march axis
              const int kidx = km * T + j;            // (ky, kx) or (kz, ky*3+kx): march axis is outermost
              const double v = (double)w[((size_t)co * Cin + ci) * ktot + kidx] * (scale ? scale[co] : 1.0);
              out[((b * 3 + slot) * Cout + co) * KC + k] = f2bf_host((float)v);
            }
      }
  return out;
}

namespace {

template <int COUT, int KC, int MODE, int MT>
int launch_inst(MarchParams& p, const MarchLaunch& L, cudaStream_t stream) {
  using G = Geo<COUT, KC, MODE, MT>;
  auto kern = conv_march_kernel<COUT, KC, MODE, MT>;
  static DeviceOnce attr_once;
  static int static_smem = 0;
  if (attr_once.first()) {
    cudaFuncAttributes fa;
    CETPICK_CUDA(cudaFuncGetAttributes(&fa, kern));
    CETPICK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      smem_limit() - (int)fa.sharedSizeBytes));
    static_smem = (int)fa.sharedSizeBytes;
  }
  p.nblk = L.nsrc * p.chunks * G::T;
  p.S = std::min(MAX_SLOTS, TMEM_COLS / (MT * COUT));
  p.SL = COUT == 32 ? p.S - 2 : p.S;
  p.row_origin = MODE == MARCH_3D_PLANES ? (int)(((long long)L.z_origin % p.SL + p.SL) % p.SL) : 0;
  const size_t wtot = align_up((size_t)p.nblk * G::WBLK, 1024);
  const size_t avail = (size_t)smem_limit() - static_smem - 1024 - wtot;
  p.stages = (int)std::min<size_t>(MAX_STAGES, avail / G::STAGE_BYTES);
  if (p.stages < 2) return CETPICK_ERR_UNSUPPORTED;
  const size_t smem = 1024 + wtot + (size_t)p.stages * G::STAGE_BYTES;

  long long base_strips;
  if (MODE == MARCH_2D_ROWS) {
    p.L = L.H;
    p.nxb = ceil_div(L.W, 128 * MT);
    base_strips = (long long)p.nxb * L.NIMG;
  } else {
    p.L = L.NIMG;
    p.nxb = ceil_div(L.W, 8 * MT);
    base_strips = (long long)p.nxb * ceil_div(L.H, 16);
  }
  // strips along the march axis: enough strips to balance the persistent grid, few enough that the
  // two extra input rows per strip stay cheap
  const int sms = num_sms();
  int best_n = 1;
  double best_eff = -1.0;
  const bool even_rows = MODE == MARCH_2D_ROWS && p.pool_out != nullptr;   // fused pool: row pairs stay in one strip
  int best_R = p.L + (even_rows ? (p.L & 1) : 0);
  for (int n = 1; n <= std::max(1, p.L / 8); ++n) {
    int R = ceil_div(p.L, n);
    if (even_rows) R += R & 1;
    const int nn = ceil_div(p.L, R);
    const long long strips = base_strips * nn;
    const double waves = (double)ceil_div<long long>(strips, sms);
    const double eff = ((double)strips / (waves * sms)) * ((double)R / (R + 2));
    if (eff > best_eff + 1e-9) { best_eff = eff; best_n = nn; best_R = R; }
  }
  p.nchunk = best_n;
  p.R = best_R;
  p.total_strips = base_strips * p.nchunk;

  int rc;
  for (int s = 0; s < L.nsrc; ++s) {
    const uint64_t C = (uint64_t)L.C;
    const uint64_t dims[4] = {C, (uint64_t)L.W, (uint64_t)L.H, (uint64_t)L.NIMG};
    const uint64_t strides[3] = {C * 2, C * 2 * L.W, C * 2 * (uint64_t)L.W * L.H};
    const uint32_t box[4] = {(uint32_t)KC, (uint32_t)G::BX, (uint32_t)G::BY, 1};
    if ((rc = tmap_encode_bf16(&p.tmA[s], L.src[s], 4, dims, strides, box, KC))) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)KC, (uint64_t)p.nblk * 3 * COUT};
    const uint64_t strides[1] = {(uint64_t)G::PIX};
    const uint32_t box[2] = {(uint32_t)KC, (uint32_t)(3 * COUT)};
    if ((rc = tmap_encode_bf16(&p.tmB, L.wpk, 2, dims, strides, box, KC))) return rc;
  }
  const int grid = (int)std::min<long long>(p.total_strips, sms);
  kern<<<grid, MARCH_THREADS, smem, stream>>>(p);
  CETPICK_LAUNCH_CHECK();
  return CETPICK_OK;
}

}  // namespace

int conv_march_launch(const MarchLaunch& L, cudaStream_t stream) {
  if (!march_supported(L.mode, L.C, L.nsrc, L.Cout)) return CETPICK_ERR_UNSUPPORTED;
  if (!L.src[0] || (L.nsrc > 1 && !L.src[1]) || !L.wpk || (!L.out && !(L.mode == MARCH_3D_PLANES && L.hm_out)) || L.NIMG <= 0 || L.H <= 0 || L.W <= 0)
    return CETPICK_ERR_BAD_ARG;
  if (L.mode == MARCH_3D_PLANES && L.dil != DIL3D) return CETPICK_ERR_UNSUPPORTED;

  MarchParams p;
  memset(&p, 0, sizeof(p));
  p.nsrc = L.nsrc;
  const int KC = std::min(L.C, 64);
  p.chunks = L.C / KC;
  p.NIMG = L.NIMG; p.H = L.H; p.W = L.W;
  p.relu = L.relu; p.bias = L.bias; p.out = static_cast<__nv_bfloat16*>(L.out);
  p.pool_out = static_cast<__nv_bfloat16*>(L.pool_out);
  if (L.pool_out && (L.mode != MARCH_2D_ROWS || !L.relu)) return CETPICK_ERR_BAD_ARG;
  if ((L.bias != nullptr) != (L.bias_host != nullptr) || (L.bias_tab != nullptr) != (L.bias_tab_host != nullptr) ||
      (L.hm_w != nullptr) != (L.hm_w_host != nullptr))
    return CETPICK_ERR_BAD_ARG;             // every epilogue constant comes with its host copy
  for (int c = 0; c < L.Cout; ++c)
    p.bias_c[c] = L.bias_tab_host ? L.bias_tab_host[63 * L.Cout + c] : L.bias_host ? L.bias_host[c] : 0.f;
  if (L.hm_w_host) memcpy(p.hmw_c, L.hm_w_host, 96 * sizeof(float));
  if (L.mode == MARCH_3D_PLANES) {
    p.bias_tab = L.bias_tab; p.hm_w = L.hm_w; p.hm_out = L.hm_out; p.hm_sigmoid = L.hm_sigmoid;
    if ((L.hm_out != nullptr) != (L.hm_w != nullptr)) return CETPICK_ERR_BAD_ARG;
  }

  if (L.mode == MARCH_3D_PLANES) return launch_inst<32, 32, MARCH_3D_PLANES, 2>(p, L, stream);
  const bool wide = L.W > 128;   // two M-tiles per step amortise the per-step barrier traffic
  if (L.Cout == 32) {
    if (KC == 16) return wide ? launch_inst<32, 16, MARCH_2D_ROWS, 2>(p, L, stream) : launch_inst<32, 16, MARCH_2D_ROWS, 1>(p, L, stream);
    if (KC == 32) return wide ? launch_inst<32, 32, MARCH_2D_ROWS, 2>(p, L, stream) : launch_inst<32, 32, MARCH_2D_ROWS, 1>(p, L, stream);
    return wide ? launch_inst<32, 64, MARCH_2D_ROWS, 2>(p, L, stream) : launch_inst<32, 64, MARCH_2D_ROWS, 1>(p, L, stream);
  }
  if (KC == 16) return CETPICK_ERR_UNSUPPORTED;
  if (KC == 32) return launch_inst<64, 32, MARCH_2D_ROWS, 1>(p, L, stream);
  return launch_inst<64, 64, MARCH_2D_ROWS, 1>(p, L, stream);
}

}  // namespace cetpick

using namespace cetpick;

#ifdef CETPICK_TEST_HOOKS   // test / tuning hooks: built into libcetpick_test_sm100a.so only (include/cetpick_test.h)
// Test hook: one marching convolution from a PyTorch-layout fp32 host weight (packs, uploads,
// launches, synchronises, frees) -- tests/test_gpu_conv.py.
extern "C" int cetpick_conv_march_pool_bf16(int mode, int dil, int nsrc, const void* src0, const void* src1, int C,
                                            int NIMG, int H, int W, const float* w_host, int Cout,
                                            const float* bias, int relu, void* out, void* pool_out, void* stream);

extern "C" int cetpick_conv_march_bf16(int mode, int dil, int nsrc, const void* src0, const void* src1, int C,
                                       int NIMG, int H, int W, const float* w_host, int Cout,
                                       const float* bias, int relu, void* out, void* stream) {
  return cetpick_conv_march_pool_bf16(mode, dil, nsrc, src0, src1, C, NIMG, H, W, w_host, Cout, bias, relu, out,
                                      nullptr, stream);
}

extern "C" int cetpick_conv_march_pool_bf16(int mode, int dil, int nsrc, const void* src0, const void* src1, int C,
                                            int NIMG, int H, int W, const float* w_host, int Cout,
                                            const float* bias, int relu, void* out, void* pool_out, void* stream) {
  g_launches = 0;
  if (!w_host) return CETPICK_ERR_BAD_ARG;
  if (!march_supported(mode, C, nsrc, Cout)) return CETPICK_ERR_UNSUPPORTED;
  std::vector<uint16_t> pk = march_pack_weights(mode, w_host, Cout, nsrc, C, nullptr);
  void* d = nullptr;
  CETPICK_CUDA(cudaMalloc(&d, pk.size() * 2));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = CETPICK_OK;
  if (cudaMemcpyAsync(d, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice, st) != cudaSuccess) rc = CETPICK_ERR_CUDA;
  std::vector<float> bh;
  if (rc == CETPICK_OK && bias) {
    bh.resize(Cout);
    if (cudaMemcpy(bh.data(), bias, (size_t)Cout * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = CETPICK_ERR_CUDA;
  }
  if (rc == CETPICK_OK) {
    MarchLaunch L;
    L.mode = mode; L.dil = dil; L.nsrc = nsrc; L.src[0] = src0; L.src[1] = src1; L.C = C;
    L.NIMG = NIMG; L.H = H; L.W = W; L.wpk = d; L.Cout = Cout; L.bias = bias; L.relu = relu; L.out = out;
    L.pool_out = pool_out;
    L.bias_host = bias ? bh.data() : nullptr;
    rc = conv_march_launch(L, st);
  }
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (rc == CETPICK_OK && e != cudaSuccess) return cuda_fail(e, "conv_march");
  return rc;
}

#endif  // CETPICK_TEST_HOOKS
