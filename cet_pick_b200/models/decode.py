"""Mirror of cet_pick/models/decode.py: same names, arguments and return shapes; the arithmetic
runs in libcetpick_sm100a.so (csrc/decode.cu).  Tie order is (score desc, linear index asc)."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib

_ws_cache = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _check_heat(heat: torch.Tensor, what: str):
    _lib.require_cuda(heat, what)
    if heat.dim() != 5 or heat.dtype != torch.float32:
        raise ValueError(f"{what}: expected a (B,C,D,H,W) float32 tensor, got {tuple(heat.shape)} {heat.dtype}")
    return heat.contiguous()


def _nms_generic(heat, kernel, mode, what):
    heat = _check_heat(heat, what)
    B, Cc, D, H, W = heat.shape
    out = torch.empty_like(heat)
    _lib.check(_lib.lib().cetpick_nms_f32(heat.data_ptr(), out.data_ptr(), B * Cc, D, H, W, int(kernel), mode,
                                          _lib.stream_ptr()), what)
    return out


def _nms(heat, kernel=3):
    """decode.py:27-33: heat * (max_pool3d(heat,(3,k,k)) == heat)."""
    return _nms_generic(heat, kernel, _lib.NMS_3D, "_nms")


def _nms_xy(heat, kernel=3):
    """decode.py:11-17."""
    return _nms_generic(heat, kernel, _lib.NMS_XY, "_nms_xy")


def _nms_z(heat, kernel=3):
    """decode.py:19-25."""
    return _nms_generic(heat, kernel, _lib.NMS_Z, "_nms_z")


def _decode_call(heat, kernel, K, nms_mode, reg, want_inds, what):
    heat = _check_heat(heat, what)
    B, Cc, D, H, W = heat.shape
    if Cc != 1:
        raise ValueError(f"{what}: the reference's .view(batch, K) needs one heat-map channel, got {Cc}")
    K = int(K)
    if reg is not None:
        _lib.require_cuda(reg, what)
        if tuple(reg.shape) != (B, 2, D, H, W):
            raise ValueError(f"{what}: reg must be (B,2,D,H,W)")
        reg = reg.float().contiguous()
    L = _lib.lib()
    nbytes = C.c_size_t(0)
    _lib.check(L.cetpick_decode_workspace_bytes(D, H, W, K, C.byref(nbytes)), what)
    ws = _workspace(heat.device, nbytes.value)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    dets = torch.empty((B, K, 5), dtype=torch.float32, device=heat.device)
    inds = torch.empty((B, K), dtype=torch.int64, device=heat.device) if want_inds else None
    _lib.check(L.cetpick_decode_f32(heat.data_ptr(), B, D, H, W, int(kernel), K, nms_mode,
                                    reg.data_ptr() if reg is not None else None, dets.data_ptr(),
                                    inds.data_ptr() if inds is not None else None, ws_ptr,
                                    ws.numel() - (ws_ptr - ws.data_ptr()), _lib.stream_ptr()), what)
    return dets, inds


def _topk(scores, K=900):
    """decode.py:82-92: top-K of the flattened map -> (scores (B,1,K), zs, ys, xs, inds (B,K)).
    zs/xs int32, ys float32 exactly like _convert_1d_to_3d (decode.py:35-41)."""
    dets, inds = _decode_call(scores, 1, K, _lib.NMS_NONE, None, True, "_topk")
    B = dets.shape[0]
    xs = (dets[:, :, 0] - 0.25).to(torch.int32)
    ys = dets[:, :, 1] - 0.25
    zs = dets[:, :, 2].to(torch.int32)
    return dets[:, :, 3].reshape(B, 1, K).contiguous(), zs, ys, xs, inds


def tomo_decode(heat, kernel=3, reg=None, K=900, if_fiber=False):
    """decode.py:123-155: (B,1,D,H,W) -> (B,K,5) rows [x+0.25|x+reg0, y+0.25|y+reg1, z, score, score]."""
    if if_fiber and int(kernel) != 3:
        # decode.py:126-128 with a (1,k,k) then (k,1,1) window: the fused scan keeps 3 planes in flight, so wider
        # fiber windows run as the two suppression passes (csrc/decode.cu nms_full_kernel) + a plain top-K
        if int(kernel) < 1 or int(kernel) % 2 == 0:
            raise ValueError("tomo_decode: the NMS kernel must be odd (the reference's hmax == heat breaks otherwise)")
        heat = _nms_z(_nms_xy(heat, kernel), kernel)
        dets, _ = _decode_call(heat, 1, K, _lib.NMS_NONE, reg, False, "tomo_decode")
        return dets
    mode = _lib.NMS_FIBER if if_fiber else _lib.NMS_3D
    dets, _ = _decode_call(heat, kernel, K, mode, reg, False, "tomo_decode")
    return dets


def non_maximum_suppression_3d(x, d, scale=1.0, threshold=float("-inf"), max_candidates=None):
    """decode.py:42-79: greedy distance-threshold suppression of a (D,H,W) score volume.  Accepts a numpy
    array or a tensor (moved to the current CUDA device); returns numpy `(scores[j], coords[j,3])` with
    coords = (x, y, z) int32, like the reference.  The loop runs on the device (csrc/greedy_nms.cu)."""
    import numpy as np
    t = torch.as_tensor(x)
    if t.dim() != 3:
        raise ValueError(f"non_maximum_suppression_3d: expected a (D,H,W) volume, got {tuple(t.shape)}")
    if not t.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("non_maximum_suppression_3d: no CUDA device (there is no CPU fallback)")
        t = t.cuda()
    t = t.float().contiguous()
    D, H, W = t.shape
    n = D * H * W
    L = _lib.lib()
    thr = float(threshold)
    cap = int(max_candidates) if max_candidates else int(min(n, max(1 << 20, int((t.double() > thr).sum().item()))))
    nbytes = C.c_size_t(0)
    _lib.check(L.cetpick_greedy_nms_workspace_bytes(D, H, W, cap, C.byref(nbytes)), "non_maximum_suppression_3d")
    ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=t.device)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    max_out = cap
    scores = torch.empty(max_out, dtype=torch.float32, device=t.device)
    coords = torch.empty((max_out, 3), dtype=torch.int32, device=t.device)
    n_out, rounds = C.c_int64(0), C.c_int(0)
    _lib.check(L.cetpick_greedy_nms_f32(t.data_ptr(), D, H, W, float(d), float(scale), thr, cap, scores.data_ptr(),
                                        coords.data_ptr(), max_out, C.byref(n_out), C.byref(rounds), ws_ptr,
                                        ws.numel() - (ws_ptr - ws.data_ptr()), _lib.stream_ptr()),
               "non_maximum_suppression_3d")
    j = min(n_out.value, max_out)
    non_maximum_suppression_3d.last_rounds = rounds.value
    return scores[:j].cpu().numpy(), coords[:j].cpu().numpy()


def tomo_decode_classify(heat, r, threshold):
    """decode.py:108-120: (C=1,D,H,W) heat -> (n,4) float32 CPU tensor rows [x, y, z, score] (the
    reference builds it from numpy on the host; callers index it with numpy, tomo_det_classify.py:184-186)."""
    heat = heat.unsqueeze(0)
    if heat.dim() != 5:
        raise ValueError(f"tomo_decode_classify: expected a (C,D,H,W) tensor, got {tuple(heat.shape[1:])}")
    vol = heat.squeeze()
    scores, coords = non_maximum_suppression_3d(vol, r, threshold=threshold)
    scores = torch.from_numpy(scores).unsqueeze(1)
    coords = torch.from_numpy(coords)
    return torch.cat([coords, scores], dim=1)


def decode_status(device=None):
    """(flags, n_candidates) of the last decode on the current stream (synchronises)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    ws = _ws_cache.get((device.index, torch.cuda.current_stream(device).cuda_stream))
    if ws is None:
        return 0, 0
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    flags, n = C.c_int(0), C.c_int64(0)
    _lib.check(_lib.lib().cetpick_decode_status(ws_ptr, _lib.stream_ptr(), C.byref(flags), C.byref(n)),
               "decode_status")
    return flags.value, n.value


def decode_debug_state(device=None):
    """First 24 words of the device-side decode state (diagnostics)."""
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    ws = _ws_cache.get((device.index, torch.cuda.current_stream(device).cuda_stream))
    if ws is None:
        return None
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    out = (C.c_uint32 * 24)()
    _lib.check(_lib.lib().cetpick_decode_debug_state(ws_ptr, _lib.stream_ptr(), out), "decode_debug_state")
    names = ["t0key", "sel_prefix", "sel_kleft", "flags", "cand_count", "n_gt", "need_fallback", "eq_need",
             "eq_zc", "done_ctr", "csel_kleft", "out_count", "csel_prefix_lo", "csel_prefix_hi", "kth_comp_lo",
             "kth_comp_hi", "n_final", "csel_done", "n_real", "t_run", "hit_total", "dense", "need_dense", "csel_wl"]
    return {n: int(out[i]) for i, n in enumerate(names)}
