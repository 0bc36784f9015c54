/*
 * cetpick.h -- C ABI of libcetpick_sm100a.so, the B200-native localisation hot path of
 * MiLoPYP / cet_pick (refinement-step inference: detector forward + heat-map decode).
 *
 * This is the drop-in boundary.  The reference is pure Python/PyTorch and
 * has no FFI of its own; each entry point below replaces the PyTorch operator sequence of the
 * reference function it cites, and is what a ctypes binding inside the reference would call
 * (see INTEGRATION.md).  Conventions:
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless named *_host;
 *   - the caller owns every input / output / workspace buffer; the library owns only plan objects;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises
 *     the host except cetpick_unet_finalize (one-off weight upload) and cetpick_decode_status;
 *   - return value 0 = CETPICK_OK, negative = error (cetpick_strerror).
 *   - there is no CPU fallback: without a CUDA device every compute call returns
 *     CETPICK_ERR_CUDA.
 */
#ifndef CETPICK_H
#define CETPICK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CETPICK_ABI_VERSION 1

enum {
  CETPICK_OK = 0,
  CETPICK_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, K > D*H*W, even NMS kernel ... */
  CETPICK_ERR_UNSUPPORTED = -2,  /* valid in the reference but not built          */
  CETPICK_ERR_WORKSPACE = -3,    /* workspace pointer null / too small / misaligned               */
  CETPICK_ERR_CUDA = -4,         /* CUDA runtime / driver error (no device, launch failure)       */
  CETPICK_ERR_STATE = -5,        /* plan not finalized / missing parameter                        */
  CETPICK_ERR_SHAPE = -6         /* parameter tensor has the wrong number of elements             */
};

int cetpick_version(void);
const char* cetpick_strerror(int code);
/* last CUDA error string seen by the calling thread's most recent failing call ("" if none) */
const char* cetpick_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------
 * Heat-map decode.  Replaces cet_pick/models/decode.py:
 *   tomo_decode  (decode.py:123-155)  = _nms / (_nms_xy -> _nms_z) -> _topk -> _convert_1d_to_3d
 *                                        -> (+0.25 | + reg gather) -> cat
 *   _nms (:27-33), _nms_xy (:11-17), _nms_z (:19-25), _topk (:82-92), _convert_1d_to_3d (:35-41),
 *   and models/utils.py:171-193 (_transpose_and_gather_feat) for `reg`.
 * Tie order (unspecified by torch.topk) is fixed to (score descending, linear index ascending).
 * ------------------------------------------------------------------------------------------ */

/* nms_mode values */
#define CETPICK_NMS_NONE  0   /* plain top-K of the map (decode.py:82-92 `_topk`)                  */
#define CETPICK_NMS_3D    1   /* (3,k,k) max-pool NMS (decode.py:27-33 `_nms`)                      */
#define CETPICK_NMS_FIBER 2   /* (1,k,k) then (k,1,1) on the suppressed map (decode.py:126-128)     */

/* bytes of device workspace cetpick_decode_f32 needs for one (D,H,W) map and K picks */
int cetpick_decode_workspace_bytes(int64_t D, int64_t H, int64_t W, int K, size_t* bytes);

/*
 * heat : (B,1,D,H,W) float32, contiguous.      reg : NULL or (B,2,D,H,W) float32.
 * dets : (B,K,5) float32 out, rows [x+0.25 | x+reg0, y+0.25 | y+reg1, z, score, score] with the
 *        reference's fp32 index arithmetic (wrong-by-design for linear indices >= 2^24).
 * inds : NULL or (B,K) int64 out, linear indices (decode.py:87).
 * kernel_xy : odd, 1..7 (opt.nms).  nms_mode : CETPICK_NMS_*.  FIBER needs kernel_xy == 3.
 * ws   : >= cetpick_decode_workspace_bytes bytes, 256-byte aligned; reused across the batch.
 * Repeated calls with identical arguments (the detector loop) are replayed as one cached CUDA graph from the third
 * call on (at most 16 argument sets are kept per process; CETPICK_DECODE_GRAPH=0 disables it).  A call made while
 * `stream` is itself being captured simply adds its launches to the caller's capture.
 */
int cetpick_decode_f32(const float* heat, int64_t B, int64_t D, int64_t H, int64_t W,
                       int kernel_xy, int K, int nms_mode, const float* reg,
                       float* dets, int64_t* inds, void* ws, size_t ws_bytes, void* stream);

/* Status of the most recent decode that used `ws` (synchronises `stream`).
 * flags bit0: heat contained NaN (reference raises ValueError later, tomo_det.py:64-65; picks are
 *             unspecified); bit1: the sampled threshold overflowed the candidate buffer and the
 *             exact full-volume select ran (slower, still exact).
 * n_candidates: entries the final select saw. */
int cetpick_decode_status(const void* ws, void* stream, int* flags, int64_t* n_candidates);

/* Diagnostics: the first 24 words of the decode state block in `ws` (t0key, select prefix, rank
 * left, flags, candidate count, n_gt, need_fallback, eq_need, eq_zc, ...).  Synchronises. */
int cetpick_decode_debug_state(const void* ws, void* stream, uint32_t* out24);

/* Pick rows of given (score, linear index) pairs with the reference's index arithmetic (decode.py:35-41 in fp32,
 * :141-154): dets[i] = [x + 0.25, y + 0.25, z, score, score].  Used to assemble the merged pick list of a z-sharded
 * volume (per-rank top-K -> all_gather -> merge-select), where the indices are global. */
int cetpick_rows_from_indices_f32(const float* scores, const int64_t* inds, int64_t n, int64_t D, int64_t H, int64_t W,
                                  float* dets, void* stream);

/* Full NMS map (decode.py:11-33): out = heat * (maxpool(heat) == heat).  mode: CETPICK_NMS_3D
 * ((3,k,k)), 3 = xy only ((1,k,k), `_nms_xy`), 4 = z only ((k,1,1), `_nms_z`). */
#define CETPICK_NMS_XY 3
#define CETPICK_NMS_Z  4
int cetpick_nms_f32(const float* heat, float* out, int64_t B, int64_t D, int64_t H, int64_t W,
                    int kernel, int mode, void* stream);

/* Greedy distance-threshold suppression = cet_pick/models/decode.py:42-79 `non_maximum_suppression_3d`
 * (through `tomo_decode_classify`, decode.py:108-120; caller detectors/tomo_det_classify.py:112,146).
 * heat: fp32 device (D,H,W).  Voxels with heat > threshold are visited in (score desc, index asc) order;
 * a visited voxel not yet suppressed is a pick and suppresses the flat-index deltas of the ball of radius
 * scale*d/2 (wrap-around across rows kept as in the reference).  scores: fp32 device [max_out];
 * coords: int32 device [max_out][3] = (x, y, z); *n_out = number of picks (may exceed max_out: only the
 * first max_out are written).  max_candidates bounds the voxels above threshold the workspace can hold
 * (CETPICK_ERR_WORKSPACE if exceeded).  Synchronises the stream (the result length is data dependent). */
int cetpick_greedy_nms_workspace_bytes(int64_t D, int64_t H, int64_t W, int64_t max_candidates, size_t* bytes);
int cetpick_greedy_nms_f32(const float* heat, int64_t D, int64_t H, int64_t W, double d, double scale,
                           double threshold, int64_t max_candidates, float* scores, int32_t* coords,
                           int64_t max_out, int64_t* n_out, int* rounds_out, void* ws, size_t ws_bytes,
                           void* stream);

/* ---- exploration-step candidate generator (first half): cet_pick/utils/image.py:42-105,138-183 ---- */

/* out = heat * (max_pool3d(heat, (kz,ky,kx), stride 1, same padding) == heat): image.py `_nms_xy` (1,k,k),
 * `_nms_z` (k,1,1), `_nms` (k,k,k).  dtype: 0 float32, 1 float64; (B, D, H, W) volumes; odd window sizes. */
int cetpick_nms_window(const void* heat, void* out, int dtype, int64_t B, int64_t D, int64_t H, int64_t W,
                       int kz, int ky, int kx, void* stream);
/* float64 twin of cetpick_greedy_nms_f32 (image.py:42-79 on the float64 DoG map): same arguments and conventions;
 * equal scores are visited in ascending index order (stable sort); scores come back as float32 like the reference. */
int cetpick_greedy_nms_f64_workspace_bytes(int64_t D, int64_t H, int64_t W, int64_t max_candidates, size_t* bytes);
int cetpick_greedy_nms_f64(const double* vol, int64_t D, int64_t H, int64_t W, double d, double scale,
                           double threshold, int64_t max_candidates, float* scores, int32_t* coords,
                           int64_t max_out, int64_t* n_out, int* rounds_out, void* ws, size_t ws_bytes,
                           void* stream);

/* Candidate patches = `extract_subvols` of cet_pick/datasets/tomo_pre_proj_angle_select_new3d_vol.py:117-128: the
 * z-slab [z - sub_z/2, z + sub_z/2] of the (sub_y x sub_x) window around each candidate is summed, min-max
 * normalised in float64 and written as float32 [n][2*(sub_y/2)][2*(sub_x/2)].  coords: int32 device [n][3] =
 * (x, y, z); the caller guarantees the windows lie inside the volume (x, y, and z - sub_z/2 >= 0). */
int cetpick_extract_subvols_f64(const double* vol, int64_t D, int64_t H, int64_t W, const int32_t* coords,
                                int64_t n, int sub_z, int sub_y, int sub_x, float* out, void* stream);

/* ---- pre-processing in front of the path: cet_pick/utils/loader.py:16-25 (quantize), :27-88
 * (load_rec), :90-121 (preprocess), all float64 like the reference ------------------------------------------------ */

/* load_rec's slice loop (loader.py:44-58,70-84): out[j][a][b] = src[a*sa + b*sb + z*sz] with z = j, or the max over
 * z in {2j, 2j+1} (z < Z) when pair_max (`--compress`).  src_dtype: 0 float32, 1 int16, 2 uint16, 3 int8, 4 float64
 * (MRC modes 2/1/6/0); strides in ELEMENTS (the caller folds the `order` axis swaps of loader.py:32-36 into them). */
int cetpick_pre_gather_f64(const void* src, int src_dtype, int64_t A, int64_t B, int64_t J, int64_t sa,
                           int64_t sb, int64_t sz, int64_t Z, int pair_max, double* out, void* stream);
/* stats[0] = mean, stats[1] = population std (np.mean / np.std, loader.py:59-60,85-86,104,118); two-stage fixed-grid
 * reduction (reproducible); ws >= cetpick_pre_stats_workspace_bytes. */
int cetpick_pre_stats_workspace_bytes(size_t* bytes);
int cetpick_pre_mean_std_f64(const double* x, int64_t n, double* stats, void* ws, size_t ws_bytes, void* stream);
/* x = (x - stats[0]) / stats[1] in place. */
int cetpick_pre_zscore_f64(double* x, int64_t n, const double* stats, void* stream);
/* One axis of scipy.ndimage.gaussian_filter(mode='reflect') (loader.py:103) on a C-contiguous (n0,n1,n2) volume;
 * weights_host = the 2*radius+1 taps of scipy's _gaussian_kernel1d (HOST); in != out. */
int cetpick_pre_gauss1d_f64(const double* in, double* out, int64_t n0, int64_t n1, int64_t n2, int axis,
                            const double* weights_host, int radius, void* stream);
/* quantize (loader.py:16-25): q = uint8(round_half_even(clip(255*(x - mi)/(ma - mi), 0, 255))). */
int cetpick_pre_quantize_u8(const double* x, int64_t n, double mi, double ma, unsigned char* q, void* stream);
/* (q - min q) / (max q - min q) (loader.py:106,120) as float64 (out_f64 = 1) or float32; minmax = device int[2]. */
int cetpick_pre_minmax_normalize(const unsigned char* q, int64_t n, int* minmax, void* out, int out_f64,
                                 void* stream);

/* models/utils.py:167-169 `_sigmoid`: x <- clamp(sigmoid(x), 1e-4, 1-1e-4), in place. */
int cetpick_sigmoid_clamp_f32(float* x, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Detector forward.  Replaces TomoConvUNet.forward (cet_pick/models/networks/unet_small.py:63-97)
 * with its UNet trunk (models/networks/unet.py:198-249, 319-399, 861-886), as built by
 * create_model('unet_N', heads, head_conv) (models/model.py:65-70).
 * Arithmetic: BF16 operands on tcgen05 tensor cores, FP32 accumulation, BatchNorm (eval) folded
 * into the convolution weights/bias, ReLU / bias / concat / pixel-shuffle fused in the epilogues.
 * ------------------------------------------------------------------------------------------ */
typedef struct cetpick_unet cetpick_unet;

/* n_blocks: 4 for unet_4, 5 for unet_5 (2..6).  head_conv: channels of feature_head (32).
 * proj_channels: classes of the 'proj' head (0 = head absent). */
int cetpick_unet_create(cetpick_unet** plan, int n_blocks, int head_conv, int proj_channels);
void cetpick_unet_destroy(cetpick_unet* plan);

/* Hand over one tensor of the reference state_dict, by its key, as host
 * float32 in PyTorch's native layout.  Keys ending in num_batches_tracked are ignored. */
int cetpick_unet_set_param(cetpick_unet* plan, const char* key, const float* data_host,
                           int64_t numel);

/* Operand precision, to be chosen before cetpick_unet_finalize: 0 = BF16 (default; the specialised kernels, heat-map
 * within 1e-2 of the fp32 reference), 1 = TF32 (fp32 activations holding TF32 values, tcgen05 kind::tf32 through the
 * generic implicit-GEMM kernel, heat-map within 1e-4; about half the tensor peak and twice the activation bytes). */
int cetpick_unet_set_precision(cetpick_unet* plan, int mode);

/* Fold BatchNorm, repack to the tensor-core layouts and upload.  Synchronous. */
int cetpick_unet_finalize(cetpick_unet* plan);

int cetpick_unet_workspace_bytes(const cetpick_unet* plan, int64_t D, int64_t H, int64_t W,
                                 int want_proj, size_t* bytes);

/*
 * tomo : (D,H,W) float32 device (the reference's (1,D,H,W) input with b == 1).
 * hm   : (D,h,w) float32 out, h = floor((H-1)/2)+1, w likewise (the 'hm' head, (1,1,D,h,w)).
 *        apply_sigmoid != 0 fuses models/utils.py:167-169 `_sigmoid` into the head epilogue.
 * proj : NULL or (C,D,h,w) float32 out, the L2-normalised 'proj' head ((1,C,D,h,w)).
 */
int cetpick_unet_forward(cetpick_unet* plan, const float* tomo, int64_t D, int64_t H, int64_t W,
                         float* hm, int apply_sigmoid, float* proj,
                         void* ws, size_t ws_bytes, void* stream);

/* Same forward from the QUANTISED tomogram.  cet_pick/utils/loader.py:90-121 `preprocess` ends with a 256-level
 * uint8 quantisation and a min-max normalisation, so the float32 volume the reference feeds the detector holds at
 * most 256 distinct values; shipping the levels is lossless and a quarter of the bytes across PCIe.
 * tomo_q: (D,H,W) uint8 device, rows 16-byte aligned (W % 16 == 0, else CETPICK_ERR_UNSUPPORTED);
 * level_values_host: HOST float32[256], the value the reference's float32 input holds for each level
 * (level_values_host[0] must be 0: level 0 doubles as the convolution's zero padding).  The stem converts
 * level -> bf16 operand exactly as the float32 entry point converts value -> bf16, so both give identical results. */
int cetpick_unet_forward_u8(cetpick_unet* plan, const uint8_t* tomo_q, const float* level_values_host,
                            int64_t D, int64_t H, int64_t W, float* hm, int apply_sigmoid, float* proj,
                            void* ws, size_t ws_bytes, void* stream);

/* One z-slab of a larger volume (the sliding-slab scheduler and the z-sharded multi-GPU forward): planes
 * [z_origin, z_origin + D) of the tomogram, given as float32 (tomo) OR as levels (tomo_q + level_values_host), the
 * other pointer NULL.  The 2-D trunk is per slice and the 3-D head reaches +-3 planes, so a slab forwarded with a
 * 3-plane recompute halo reproduces the whole-volume heat-map on its core planes; z_origin makes the accumulation
 * order of the head depend on the ABSOLUTE plane index only, so the reproduction is bit-identical. */
int cetpick_unet_forward_slab(cetpick_unet* plan, const float* tomo, const uint8_t* tomo_q,
                              const float* level_values_host, int64_t D, int64_t H, int64_t W, int64_t z_origin,
                              float* hm, int apply_sigmoid, float* proj, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Exploration-step embedding network.  Replaces TomoResClassifier.forward_test
 * (cet_pick/models/networks/simsiam_model.py:325-366; arch simsiam_18 / simsiam3d_18), called per batch of candidate
 * sub-volumes by simsiam_test_hm_3d.py:169.  Same arithmetic conventions as the detector: BF16 operands on tcgen05
 * tensor cores, FP32 accumulation, every eval-mode BatchNorm folded, residual adds / ReLU in the epilogues.
 * ------------------------------------------------------------------------------------------ */
typedef struct cetpick_simsiam cetpick_simsiam;

/* blocks1..3: BasicBlocks per stage (2,2,2 for *_18; 3,4,6 for *_34).  has_proj / has_pred: which heads exist. */
int cetpick_simsiam_create(cetpick_simsiam** plan, int blocks1, int blocks2, int blocks3, int has_proj, int has_pred);
/* 2-D exploration variant: TomoResClassifier2D.forward_test (cet_pick/models/networks/simsiam_model_2d.py:617-774, arch
 * simsiam2d_18; conv1 3x3 stride 1, no max-pool, AdaptiveAvgPool2d, fc 256 -> out_dim, heads of width out_dim).
 * out_dim = the reference's head_conv (128 by default for this task, opts.py:207-209; 32 by default in the factory's
 * signature), 1 ... 256; widths that are not multiples of 64 are computed zero-padded. */
int cetpick_simsiam_create_2d(cetpick_simsiam** plan, int blocks1, int blocks2, int blocks3, int out_dim,
                              int has_proj, int has_pred);
void cetpick_simsiam_destroy(cetpick_simsiam* plan);
/* one tensor of the reference state_dict by its key, host float32, PyTorch layout */
int cetpick_simsiam_set_param(cetpick_simsiam* plan, const char* key, const float* data_host, int64_t numel);
int cetpick_simsiam_finalize(cetpick_simsiam* plan);
int cetpick_simsiam_workspace_bytes(const cetpick_simsiam* plan, int64_t B, int64_t D, int64_t H, int64_t W, size_t* bytes);
/* x: (B,D,H,W) float32 device sub-volumes, H = W = 32 or 16, any depth D <= 4096 (D = 32 at H = 32 fills the Conv3d
 * layer's 128-row tile exactly; the reference's exploration dataset feeds D = 1 slab sums); other H, W return
 * CETPICK_ERR_UNSUPPORTED.
 * A 2-D plan takes (B,1,H,W) patches, D = 1, H = W in {8, 16, 32, 64}.
 * proj / pred: (B,256) float32 device out ((B,out_dim) for a 2-D plan), either may be NULL. */
int cetpick_simsiam_forward(cetpick_simsiam* plan, const float* x, int64_t B, int64_t D, int64_t H, int64_t W,
                            float* proj, float* pred, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Refinement training step (BASELINE.json configs[4]): losses, optimiser, and -- further down -- the
 * training-mode layers (cetpick_train_conv_f32 ...) that cet_pick_b200/trains/engine.py strings into the forward and
 * backward of the detector.
 * ------------------------------------------------------------------------------------------ */
int cetpick_train_workspace_bytes(size_t* bytes);
/* cet_pick/models/loss.py:255-325 `PULoss(tau)(pred, gt)`, pred = `_sigmoid(logits)` (models/utils.py:167-169) when
 * apply_sigmoid -- the training loss of trains/tomo_cr_semi_trainer.py:43-60 without `--contrastive`.  gt: 1 labelled
 * positive, (-1,1) soft positive, -1 unlabelled.  out4 (device): loss, positive risk, negative risk, #positives (the
 * reference raises when that is 0).  grad (device [n], nullable): grad_scale * d loss / d logits.  ws: 256-byte aligned,
 * >= cetpick_train_workspace_bytes.  Reductions are two-stage with a fixed grid: reproducible. */
int cetpick_pu_loss_f32(const float* logits, const float* gt, int64_t n, int apply_sigmoid, double tau, double beta,
                        float* out4, float* grad, float grad_scale, void* ws, size_t ws_bytes, void* stream);
/* loss.py:701-715 `ConsistencyLoss` = mse_loss(a, b); out1 (device); grad_a (nullable) = grad_scale * d loss / d a. */
int cetpick_mse_loss_f32(const float* a, const float* b, int64_t n, float* out1, float* grad_a, float grad_scale,
                         void* ws, size_t ws_bytes, void* stream);
/* One torch.optim.Adam step (main.py:55: Adam(model.parameters(), lr), betas (0.9, 0.999), eps 1e-8) fused over a flat
 * fp32 bucket of n parameters; step counts from 1; gradients are multiplied by grad_scale first (1 / world size after a
 * summed all-reduce of the same bucket). */
int cetpick_adam_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, double lr,
                          double beta1, double beta2, double eps, double weight_decay, int64_t step, double grad_scale,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Training-mode layers, fp32 (csrc/train_net.cu): what `loss.backward()` of trains/base_trainer.py:484-489 runs through
 * for models/networks/unet_small.py:63-97 and unet.py:198-249,319-399.  Tensors have dense rows and explicit (slice,
 * channel) element strides: (D,C,h,w) of the 2-D trunk and (C,D,h,w) of the 3-D head are the same buffer, a channel
 * concat is a channel offset.  Gradients of parameters ACCUMULATE (+=) into their buffers like torch's `.grad`.
 * ------------------------------------------------------------------------------------------ */
typedef struct cetpick_conv_geom {
  int N, Cin, Cout;            /* slices (2-D: batch; 3-D: z), channels                                             */
  int H, W, Ho, Wo;            /* input / output rows and columns                                                   */
  long long xs_n, xs_c;        /* input element strides of slice and channel                                        */
  long long ys_n, ys_c;        /* output (forward) / output-gradient (wgrad) strides                                */
  int kz, ky, kx;              /* taps; kz > 1 = 3-D conv over the slice axis                                       */
  int dz, dy, dx;              /* dilation                                                                          */
  int pz, py, px;              /* zero padding                                                                      */
  int stride;                  /* in-plane stride (1; 2 for the stem and for the transposed conv's data gradient)   */
  int zdepth;                  /* slices per crop: z taps never cross a multiple of zdepth (N % zdepth == 0)        */
} cetpick_conv_geom;

/* y = [relu]([y +] conv(x, w) + bias); flags: bit 0 accumulate into y, bit 1 ReLU.  w: [Cout][Cin][kz][ky][kx] (torch
 * layout), bias nullable.  Also the DATA GRADIENT of a
 * stride-1 conv: x = dy, w = cetpick_train_flip_weights_f32(w) ([Cin][Cout][reversed taps]), pad' = (k-1)*dil - pad, and of
 * ConvTranspose2d(2, stride 2): x = dy, w = the transposed conv's own [Cin][Cout][2][2] weights, 2x2 taps, stride 2. */
int cetpick_train_conv_f32(const float* x, const float* w, const float* bias, float* y, const cetpick_conv_geom* g,
                           int flags, void* stream);
int cetpick_train_flip_weights_f32(const float* w, float* wt, int Cout, int Cin, int taps, void* stream);
/* dw[Cout][Cin][taps] += sum over (slice, pixel) of dy * shifted x  (tap sets 3x3, 3x3x3, 7x7, 1x1, 3x1x1, 2x2; others
 * CETPICK_ERR_UNSUPPORTED).  Sums in fp32 with atomics: order, hence the last bits, vary from run to run (like cuDNN's
 * default weight-gradient algorithms). */
int cetpick_train_conv_wgrad_f32(const float* x, const float* dy, float* dw, const cetpick_conv_geom* g, void* stream);
/* 1 (default): the wide layers contract on the tensor cores with operands rounded to TF32 and fp32 accumulation -- the
 * arithmetic of the reference's own training run (PyTorch's default torch.backends.cudnn.allow_tf32 = True); 0: fp32 FMA. */
int cetpick_train_set_tf32(int on);
/* ConvTranspose2d(Cin, Cout, 2, stride 2) + bias cropped to (Ho, Wo) <= (2H, 2W) (unet.py:285-292,375-380); w [Cin][Cout][2][2]. */
int cetpick_train_upconv_f32(const float* x, const float* w, const float* bias, float* y, const cetpick_conv_geom* g, void* stream);
/* workspace of the reductions below for up to C_max channels (256-byte aligned) */
int cetpick_train_net_workspace_bytes(int C_max, size_t* bytes);
/* BatchNorm2d in training mode (+ ReLU when relu): batch mean / biased variance over (N, HW) per channel -> save_mean,
 * save_invstd; running statistics updated with `momentum` (unbiased variance), nullable. */
int cetpick_train_bn_f32(const float* x, long long xs_n, long long xs_c, float* y, long long ys_n, long long ys_c,
                         const float* gamma, const float* beta, float* running_mean, float* running_var, float* save_mean,
                         float* save_invstd, int N, int C, int HW, float eps, float momentum, int relu, void* ws, size_t ws_bytes,
                         void* stream);
/* Backward of the above: dy is the gradient w.r.t. the (ReLU'd) output y; dx w.r.t. the BatchNorm input x; dgamma / dbeta
 * accumulate (nullable).  y and dy share strides. */
int cetpick_train_bn_bwd_f32(const float* x, long long xs_n, long long xs_c, const float* y, const float* dy, long long ys_n,
                             long long ys_c, float* dx, long long dxs_n, long long dxs_c, const float* gamma,
                             const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, int N, int C, int HW,
                             int relu, void* ws, size_t ws_bytes, void* stream);
/* out[c] (+)= sum over (N, HW) of x: bias gradients */
int cetpick_train_channel_sum_f32(const float* x, long long xs_n, long long xs_c, float* out, int N, int C, int HW, int accumulate,
                                  void* ws, size_t ws_bytes, void* stream);
/* MaxPool2d(2, ceil_mode=True) (unet.py:225) and its backward (the first maximum of a window takes the gradient) */
int cetpick_train_pool_f32(const float* x, long long xs_n, long long xs_c, float* y, long long ys_n, long long ys_c, int N, int C,
                           int H, int W, void* stream);
int cetpick_train_pool_bwd_f32(const float* x, long long xs_n, long long xs_c, const float* dy, long long dys_n, long long dys_c,
                               float* dx, long long dxs_n, long long dxs_c, int N, int C, int H, int W, int accumulate, void* stream);
/* dx = dy * (y > 0), contiguous */
int cetpick_train_relu_bwd_f32(const float* y, const float* dy, float* dx, size_t n, void* stream);

/* Number of kernels the most recent cetpick_unet_forward / cetpick_decode_f32 on this thread
 * enqueued (bench.py's gpu_launches). */
int64_t cetpick_last_launch_count(void);

/* Per-launch timing of cetpick_unet_forward with CUDA events on the launching stream (bench.py's roofline), kept in
 * the plan.  enable(plan, 1), run a forward of that plan, then read(): n entries of (milliseconds, algorithmic FLOPs,
 * 32-byte name) in launch order.  read() synchronises; profiling adds one event per launch. */
int cetpick_unet_profile_enable(cetpick_unet* plan, int on);
int cetpick_unet_profile_read(cetpick_unet* plan, int max_entries, int* n, float* ms, double* flops, char* names32);

#ifdef __cplusplus
}
#endif
#endif /* CETPICK_H */
