"""ctypes binding of libcetpick_sm100a.so (include/cetpick.h).  No CPU fallback: if the library
cannot be loaded, or a compute call is made without a CUDA device, this raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CETPICK_LIB") or os.path.join(_PKG, "libcetpick_sm100a.so")   # explicit override for A/B builds
TEST_LIB_PATH = os.environ.get("CETPICK_TEST_LIB") or os.path.join(_PKG, "libcetpick_test_sm100a.so")

OK = 0
ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA, ERR_STATE, ERR_SHAPE = -1, -2, -3, -4, -5, -6
NMS_NONE, NMS_3D, NMS_FIBER, NMS_XY, NMS_Z = 0, 1, 2, 3, 4
EPI_BF16_NHWC, EPI_UPCONV_2X2, EPI_F32_ROWMAJOR, EPI_F32_L2NORM_NCDHW = 0, 1, 2, 3
MARCH_2D_ROWS, MARCH_3D_PLANES = 0, 1

_i64, _int, _vp, _sz, _ll = C.c_int64, C.c_int, C.c_void_p, C.c_size_t, C.c_longlong

# every symbol include/cetpick.h declares: (restype, argtypes)
SIGNATURES = {
    "cetpick_version": (_int, []),
    "cetpick_strerror": (C.c_char_p, [_int]),
    "cetpick_last_cuda_error": (C.c_char_p, []),
    "cetpick_decode_workspace_bytes": (_int, [_i64, _i64, _i64, _int, C.POINTER(_sz)]),
    "cetpick_decode_f32": (_int, [_vp, _i64, _i64, _i64, _i64, _int, _int, _int, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cetpick_decode_status": (_int, [_vp, _vp, C.POINTER(_int), C.POINTER(_i64)]),
    "cetpick_decode_debug_state": (_int, [_vp, _vp, _vp]),
    "cetpick_rows_from_indices_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp]),
    "cetpick_nms_f32": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _vp]),
    "cetpick_greedy_nms_workspace_bytes": (_int, [_i64, _i64, _i64, _i64, C.POINTER(_sz)]),
    "cetpick_greedy_nms_f32": (_int, [_vp, _i64, _i64, _i64, C.c_double, C.c_double, C.c_double, _i64, _vp, _vp, _i64,
                                      C.POINTER(_i64), C.POINTER(_int), _vp, _sz, _vp]),
    "cetpick_nms_window": (_int, [_vp, _vp, _int, _i64, _i64, _i64, _i64, _int, _int, _int, _vp]),
    "cetpick_greedy_nms_f64_workspace_bytes": (_int, [_i64, _i64, _i64, _i64, C.POINTER(_sz)]),
    "cetpick_greedy_nms_f64": (_int, [_vp, _i64, _i64, _i64, C.c_double, C.c_double, C.c_double, _i64, _vp, _vp, _i64,
                                      C.POINTER(_i64), C.POINTER(_int), _vp, _sz, _vp]),
    "cetpick_extract_subvols_f64": (_int, [_vp, _i64, _i64, _i64, _vp, _i64, _int, _int, _int, _vp, _vp]),
    "cetpick_pre_gather_f64": (_int, [_vp, _int, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _int, _vp, _vp]),
    "cetpick_pre_stats_workspace_bytes": (_int, [C.POINTER(_sz)]),
    "cetpick_pre_mean_std_f64": (_int, [_vp, _i64, _vp, _vp, _sz, _vp]),
    "cetpick_pre_zscore_f64": (_int, [_vp, _i64, _vp, _vp]),
    "cetpick_pre_gauss1d_f64": (_int, [_vp, _vp, _i64, _i64, _i64, _int, _vp, _int, _vp]),
    "cetpick_pre_quantize_u8": (_int, [_vp, _i64, C.c_double, C.c_double, _vp, _vp]),
    "cetpick_pre_minmax_normalize": (_int, [_vp, _i64, _vp, _vp, _int, _vp]),
    "cetpick_sigmoid_clamp_f32": (_int, [_vp, _i64, _vp]),
    "cetpick_unet_create": (_int, [C.POINTER(_vp), _int, _int, _int]),
    "cetpick_unet_destroy": (None, [_vp]),
    "cetpick_unet_set_param": (_int, [_vp, C.c_char_p, _vp, _i64]),
    "cetpick_unet_set_precision": (_int, [_vp, _int]),
    "cetpick_unet_finalize": (_int, [_vp]),
    "cetpick_unet_workspace_bytes": (_int, [_vp, _i64, _i64, _i64, _int, C.POINTER(_sz)]),
    "cetpick_unet_forward": (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _int, _vp, _vp, _sz, _vp]),
    "cetpick_unet_forward_u8": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _int, _vp, _vp, _sz, _vp]),
    "cetpick_unet_forward_slab": (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _int, _vp, _vp, _sz, _vp]),
    "cetpick_simsiam_create": (_int, [C.POINTER(_vp), _int, _int, _int, _int, _int]),
    "cetpick_simsiam_create_2d": (_int, [C.POINTER(_vp), _int, _int, _int, _int, _int, _int]),
    "cetpick_simsiam_destroy": (None, [_vp]),
    "cetpick_simsiam_set_param": (_int, [_vp, C.c_char_p, _vp, _i64]),
    "cetpick_simsiam_finalize": (_int, [_vp]),
    "cetpick_simsiam_workspace_bytes": (_int, [_vp, _i64, _i64, _i64, _i64, C.POINTER(_sz)]),
    "cetpick_simsiam_forward": (_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "cetpick_train_workspace_bytes": (_int, [C.POINTER(_sz)]),
    "cetpick_pu_loss_f32": (_int, [_vp, _vp, _i64, _int, C.c_double, C.c_double, _vp, _vp, C.c_float, _vp, _sz, _vp]),
    "cetpick_mse_loss_f32": (_int, [_vp, _vp, _i64, _vp, _vp, C.c_float, _vp, _sz, _vp]),
    "cetpick_adam_step_f32": (_int, [_vp, _vp, _vp, _vp, _i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                     _i64, C.c_double, _vp]),
    "cetpick_train_conv_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "cetpick_train_flip_weights_f32": (_int, [_vp, _vp, _int, _int, _int, _vp]),
    "cetpick_train_conv_wgrad_f32": (_int, [_vp, _vp, _vp, _vp, _vp]),
    "cetpick_train_set_tf32": (_int, [_int]),
    "cetpick_train_upconv_f32": (_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "cetpick_train_net_workspace_bytes": (_int, [_int, C.POINTER(_sz)]),
    "cetpick_train_bn_f32": (_int, [_vp, _ll, _ll, _vp, _ll, _ll, _vp, _vp, _vp, _vp, _vp, _vp, _int, _int, _int, C.c_float,
                                    C.c_float, _int, _vp, _sz, _vp]),
    "cetpick_train_bn_bwd_f32": (_int, [_vp, _ll, _ll, _vp, _vp, _ll, _ll, _vp, _ll, _ll, _vp, _vp, _vp, _vp, _vp, _int, _int,
                                        _int, _int, _vp, _sz, _vp]),
    "cetpick_train_channel_sum_f32": (_int, [_vp, _ll, _ll, _vp, _int, _int, _int, _int, _vp, _sz, _vp]),
    "cetpick_train_pool_f32": (_int, [_vp, _ll, _ll, _vp, _ll, _ll, _int, _int, _int, _int, _vp]),
    "cetpick_train_pool_bwd_f32": (_int, [_vp, _ll, _ll, _vp, _ll, _ll, _vp, _ll, _ll, _int, _int, _int, _int, _int, _vp]),
    "cetpick_train_relu_bwd_f32": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "cetpick_last_launch_count": (_i64, []),
    "cetpick_unet_profile_enable": (_int, [_vp, _int]),
    "cetpick_unet_profile_read": (_int, [_vp, _int, C.POINTER(_int), _vp, _vp, _vp]),
}

# test / tuning hooks of include/cetpick_test.h (libcetpick_test_sm100a.so only)
TEST_SIGNATURES = {
    "cetpick_conv_small_bf16": (_int, [_vp, _int, _int, _int, _int, _int, _int, _int, _int, _vp, _int, _int, _vp, _vp,
                                       _vp, _int, _int, _vp, _vp]),
    "cetpick_selftest_gemm_bf16": (_int, [_vp, _vp, _vp, _int, _int, _int, _vp]),
    "cetpick_probe_umma": (_int, [_vp, _int, _vp, _int, _int, _int, _int, _vp, _vp]),
    "cetpick_conv_march_bf16": (_int, [_int, _int, _int, _vp, _vp, _int, _int, _int, _int, _vp, _int, _vp, _int,
                                       _vp, _vp]),
    "cetpick_conv_march_pool_bf16": (_int, [_int, _int, _int, _vp, _vp, _int, _int, _int, _int, _vp, _int, _vp, _int,
                                            _vp, _vp, _vp]),
    "cetpick_conv_block_bf16": (_int, [_int, _vp, _vp, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cetpick_block_debug_buffer": (_int, [_vp]),
    "cetpick_upconv_bf16": (_int, [_vp, _int, _int, _int, _int, _vp, _vp, _int, _vp, _int, _int, _vp]),
    "cetpick_conv_halo_bf16": (_int, [_int, _vp, _vp, _int, _int, _int, _int, _vp, _vp, _int, _int, _vp, _vp]),
    "cetpick_conv_stem_bf16": (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp]),
    "cetpick_decode_set_stop_stage": (_int, [_int]),
    "cetpick_decode_graph_hits": (_i64, []),
    "cetpick_tmap_cache_stats": (_int, [C.POINTER(_i64), C.POINTER(_i64)]),
    "cetpick_probe_mma_rate": (_int, [_int, _int, _int, _int, _int, _int, _int, _vp, _int, _vp]),
    "cetpick_probe_mma_rate2": (_int, [_int, _int, _int, _int, _int, _int, _int, _vp, _int, _vp]),
    "cetpick_conv_bf16": (_int, [_int, _vp, _int, _vp, _int, _int, _int, _int, _vp, _int, _int, _vp, _int,
                                 _vp, _int, _int, _vp, _int, _int, _int, _int, _vp]),
}

_lib = None
_lock = threading.Lock()


class CetpickError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where}: {detail}" if detail else where)


def lib() -> C.CDLL:
    """Load the shared library once; raise loudly when it is missing (no silent fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} is missing: build it with `python -m cet_pick_b200.build` "
                        "(cet_pick_b200 has no CPU/PyTorch fallback)")
                h = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(h, name)          # AttributeError if the symbol is not exported
                    fn.restype, fn.argtypes = res, args
                _lib = h
    return _lib


_test_lib = None


def test_lib() -> C.CDLL:
    """The test twin of the product library: the same objects plus the per-kernel hooks and hardware probes of
    include/cetpick_test.h.  Used by tests/ and scripts/ only; nothing in the product package calls it."""
    global _test_lib
    if _test_lib is None:
        with _lock:
            if _test_lib is None:
                if not os.path.exists(TEST_LIB_PATH):
                    raise ImportError(f"{TEST_LIB_PATH} is missing: build it with `python -m cet_pick_b200.build`")
                h = C.CDLL(TEST_LIB_PATH)
                for name, (res, args) in {**SIGNATURES, **TEST_SIGNATURES}.items():
                    fn = getattr(h, name)
                    fn.restype, fn.argtypes = res, args
                _test_lib = h
    return _test_lib


def check(code: int, where: str):
    if code == OK:
        return
    L = lib()
    msg = L.cetpick_strerror(code).decode()
    if code == ERR_CUDA:
        msg += " (" + L.cetpick_last_cuda_error().decode() + ")"
    exc = ValueError if code in (ERR_BAD_ARG, ERR_SHAPE) else \
        NotImplementedError if code == ERR_UNSUPPORTED else RuntimeError
    raise exc(f"{where}: {msg}")


def require_cuda(t, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: tensor is on {t.device}; cet_pick_b200 runs on CUDA (sm_100a) only "
                           "and has no CPU fallback")
    import torch
    if t.device.index != torch.cuda.current_device():
        # plans, workspaces and the launch stream belong to the current device (one process per GPU)
        raise RuntimeError(f"{what}: tensor is on {t.device} but the current CUDA device is "
                           f"cuda:{torch.cuda.current_device()}; call torch.cuda.set_device first")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def profile_forward(model, fn):
    """Run fn() (one or more forwards of `model`, a TomoConvUNet) with per-launch CUDA-event timing kept in the model's
    plan and return [(name, ms, flops)] of the LAST forward."""
    L = lib()
    plan = model.plan()
    L.cetpick_unet_profile_enable(plan, 1)
    try:
        fn()
        n = C.c_int(0)
        ms = (C.c_float * 256)()
        fl = (C.c_double * 256)()
        names = C.create_string_buffer(256 * 32)
        check(L.cetpick_unet_profile_read(plan, 256, C.byref(n), ms, fl, names), "cetpick_unet_profile_read")
        return [(names.raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode(), float(ms[i]), float(fl[i]))
                for i in range(n.value)]
    finally:
        L.cetpick_unet_profile_enable(plan, 0)
