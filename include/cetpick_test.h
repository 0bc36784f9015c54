/*
 * cetpick_test.h -- test and tuning hooks of libcetpick_test_sm100a.so.
 *
 * NOT part of the drop-in boundary (include/cetpick.h): these entry points run ONE kernel of the hot path from
 * PyTorch-layout host weights (they pack, upload, launch and synchronise), or probe the hardware, so that tests/
 * can check every kernel in isolation against torch fp32 and tuning scripts can time instruction shapes.
 * libcetpick_test_sm100a.so = the product objects compiled with -DCETPICK_TEST_HOOKS; the product library
 * libcetpick_sm100a.so does not export any of them.
 */
#ifndef CETPICK_TEST_H
#define CETPICK_TEST_H

#include "cetpick.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Self-test of the tcgen05 implicit-GEMM building block: C[M,N] = A[M,K] * B[N,K]^T (bf16 in,
 * fp32 out) on device buffers; used by tests to validate descriptors in isolation. */
int cetpick_selftest_gemm_bf16(const void* A, const void* B, float* C, int M, int N, int K,
                               void* stream);

/* Test hook: ONE convolution through the tcgen05 implicit-GEMM kernel with caller-packed bf16
 * weights [k-block][Ntot][KC] (k-block = (source, tap, channel chunk)); taps = ntaps x (dz,dy,dx).
 * epi: 0 bf16 NHWC, 1 ConvTranspose 2x2 scatter, 2 fp32 row-major, 3 fp32 L2-normalised NCDHW. */
int cetpick_conv_bf16(int nsrc, const void* src0, int C0, const void* src1, int C1, int NIMG,
                      int H, int W, const void* wpk, int KC, int ntaps, const int* taps, int Ntot,
                      const float* bias, int relu, int epi, void* out, int out_cstride,
                      int Ho, int Wo, int Cout, void* stream);

/* Test hook: ONE convolution through the marching tcgen05 kernel (csrc/conv_march.cu) used for the
 * Cout 32/64 layers.  mode 0: Conv2d 3x3 pad 1 over NIMG images; mode 1: Conv3d 3x3x3 dilation
 * (1,dil,dil) pad (1,dil,dil) with NIMG = depth.  src*: bf16 device [NIMG][H][W][C] (nsrc = 2 is the
 * channel concat of two sources); w_host: fp32 HOST weight in PyTorch layout (Cout, nsrc*C, 3, 3[, 3]);
 * out: bf16 device [NIMG][H][W][Cout].  Packs, uploads, launches and synchronises. */
int cetpick_conv_march_bf16(int mode, int dil, int nsrc, const void* src0, const void* src1, int C,
                            int NIMG, int H, int W, const float* w_host, int Cout,
                            const float* bias, int relu, void* out, void* stream);

/* Same as cetpick_conv_march_bf16 (mode 0, relu != 0) with the fused MaxPool2d(2, ceil_mode=True) of
 * cet_pick/models/networks/unet.py:225,237-238: pool_out = bf16 device [NIMG][(H+1)/2][(W+1)/2][Cout] or null. */
int cetpick_conv_march_pool_bf16(int mode, int dil, int nsrc, const void* src0, const void* src1, int C,
                                 int NIMG, int H, int W, const float* w_host, int Cout,
                                 const float* bias, int relu, void* out, void* pool_out, void* stream);

/* Test hook: ConvTranspose2d(Cin,Cout,2,stride 2)+bias+ReLU through csrc/conv_up.cu.  src: bf16 device
 * [NIMG][h][w][Cin]; w_host: fp32 HOST weight in PyTorch layout (Cin,Cout,2,2); bias_host: fp32 HOST
 * [Cout]; out: bf16 device [NIMG][Ho][Wo][Cout] with Ho <= 2h, Wo <= 2w (autocrop).  Synchronises. */
int cetpick_upconv_bf16(const void* src, int Cin, int NIMG, int h, int w, const float* w_host,
                        const float* bias_host, int Cout, void* out, int Ho, int Wo, void* stream);

/* Test hook: ONE 3x3 Conv2d (pad 1) + bias (+ReLU) through the halo-tile tcgen05 kernel (csrc/conv_halo.cu)
 * used for the wide trunk levels (C per source a multiple of 64, Cout a multiple of 128).  src*: bf16 device
 * [NIMG][H][W][C]; w_host: fp32 HOST weight (Cout, nsrc*C, 3, 3); bias_host: fp32 HOST [Cout];
 * out: bf16 device [NIMG][H][W][Cout].  Packs, uploads, launches and synchronises. */
int cetpick_conv_halo_bf16(int nsrc, const void* src0, const void* src1, int C, int NIMG, int H, int W,
                           const float* w_host, const float* bias_host, int Cout, int relu, void* out,
                           void* stream);

/* Test hook: the detector stem Conv2d(1,16,7,stride 2,pad 3)+BN+ReLU through the tensor-core march of
 * csrc/conv_stem.cu (replaces cet_pick/models/networks/unet_small.py:35-37,72-74).  in: fp32 device (D,H,W)
 * with W % 4 == 0; w_host: fp32 HOST weight (16,1,7,7); scale_host/shift_host: folded BatchNorm (HOST, [16],
 * nullable); out: bf16 device (D,(H-1)/2+1,(W-1)/2+1,16).  Packs, uploads, launches and synchronises. */
int cetpick_conv_stem_bf16(const float* in, int D, int H, int W, const float* w_host,
                           const float* scale_host, const float* shift_host, void* out, void* stream);

/* Test hook: the fused full-resolution block of csrc/conv_block.cu: conv3x3 + bias + ReLU -> conv3x3 + bias + ReLU
 * (-> MaxPool2d(2, ceil)) with the intermediate map kept in shared memory (cet_pick/models/networks/unet.py:198-249,
 * 375-399).  src*: bf16 device [NIMG][H][W][C1] (C1 = 16, or 32 with nsrc 1 / 2); w1_host (32, nsrc*C1, 3, 3),
 * w2_host (32, 32, 3, 3), bias*_host [32]: fp32 HOST; out: bf16 device [NIMG][H][W][32]; pool_out: bf16 device
 * [NIMG][(H+1)/2][(W+1)/2][32] or null.  W <= 1024 (one cluster of <= 8 CTAs spans a row).  Packs, uploads, launches and synchronises. */
int cetpick_conv_block_bf16(int nsrc, const void* src0, const void* src1, int C1, int NIMG, int H, int W,
                            const float* w1_host, const float* bias1_host, const float* w2_host,
                            const float* bias2_host, void* out, void* pool_out, void* stream);

/* Bring-up aid for csrc/conv_block.cu: registers (first call) and returns a 256-word HOST buffer the kernel can write;
 * word 0 = number of mbarrier waits that timed out, 4-word records from word 4: (cta << 8 | warp, wait site, a, b). */
int cetpick_block_debug_buffer(uint32_t** host_buf);

/* Test hook: ONE convolution through the small-map implicit-GEMM kernel (csrc/conv_small.cu).  src: bf16 device
 * [B][Z][Hin][Win][C] (C a multiple of 64); w_host: fp32 HOST weight (Cout, C, ntaps); taps: ntaps x (dz,dy,dx) input
 * offsets relative to stride * output position; bias: fp32 device [Cout] or null; residual: bf16 device
 * [B][Z][Ho][Wo][Cout] or null; out: bf16 (out_f32 = 0) or fp32 device [B][Z][Ho][Wo][Cout].  Wo*Ho must divide 128. */
int cetpick_conv_small_bf16(const void* src, int C, int B, int Z, int Hin, int Win, int stride, int Ho, int Wo,
                            const float* w_host, int Cout, int ntaps, const int* taps, const float* bias,
                            const void* residual, int relu, int out_f32, void* out, void* stream);

/* Hardware probe (test hook): D[128][32] = A_big[rows] * B^T where the A descriptor starts r0 rows
 * into a TMA-written swizzled tile, with 8-row groups sbo_bytes apart and the given base_offset. */
int cetpick_probe_umma(const void* A_big, int R, const void* B, int KC, int r0, int sbo_bytes,
                       int base_offset, float* out, void* stream);

/* Hardware probe (tuning hook): cycles per tcgen05.mma (M=128, N, K=16, bf16) issued back to back from
 * shared-memory operands; KC selects the swizzle width, sbo_a the A 8-row group stride, the A start
 * address cycles through ntap offsets a_step bytes apart and the accumulator through ndst TMEM
 * regions.  out_cycles: `grid` floats (device). */
int cetpick_probe_mma_rate(int N, int KC, int sbo_a, int a_step, int ntap, int ndst, int iters,
                           float* out_cycles, int grid, void* stream);

/* Same measurement for a CTA pair: tcgen05.mma.cta_group::2 (M = 256 over two SMs).  out_cycles: `pairs` floats. */
int cetpick_probe_mma_rate2(int N, int KC, int sbo_a, int a_step, int ntap, int ndst, int iters,
                            float* out_cycles, int pairs, void* stream);

/* (hits, misses) of the process-wide tensor-map cache: a forward repeated on the same plan, buffers and shape encodes no
 * new CUtensorMap (csrc/conv_tc.cu). */
int cetpick_tmap_cache_stats(int64_t* hits, int64_t* misses);

/* Stage timing of the decode (scripts/decode_stages.py): cetpick_decode_f32 returns after stage n of its launch
 * sequence (1 init, 2 sample select, 3 sieve, 4 gated fall-backs + EQ pass, 5 tail kernel); 0 = all. */
int cetpick_decode_set_stop_stage(int n);
/* calls of cetpick_decode_f32 in this process that were served by a cached CUDA graph (one cudaGraphLaunch) */
int64_t cetpick_decode_graph_hits(void);

#ifdef __cplusplus
}
#endif
#endif /* CETPICK_TEST_H */
