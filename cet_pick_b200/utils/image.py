"""Mirror of the candidate-generation functions of cet_pick/utils/image.py used by the exploration step
(`_nms_xy` :81-87, `_nms_z` :89-95, `_nms` :97-105, `non_maximum_suppression_3d` :42-79,
`get_potential_coords_pyramid` :138-183): same names, arguments and results; the difference-of-Gaussians volume
stays on the device in float64 like the reference's numpy/scipy/torch-double chain, every stage is a kernel of
csrc/preproc.cu (Gaussians) or csrc/explore.cu (NMS maps, greedy distance suppression)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from .loader import _to_device, gaussian_filter


def _nms_window(heat, kz, ky, kx, what):
    _lib.require_cuda(heat, what)
    if heat.dim() != 5 or heat.dtype not in (torch.float32, torch.float64):
        raise ValueError(f"{what}: expected a (B,C,D,H,W) float32/float64 tensor, got {tuple(heat.shape)} {heat.dtype}")
    heat = heat.contiguous()
    B, Cc, D, H, W = heat.shape
    out = torch.empty_like(heat)
    _lib.check(_lib.lib().cetpick_nms_window(heat.data_ptr(), out.data_ptr(), 0 if heat.dtype == torch.float32 else 1,
                                             B * Cc, D, H, W, int(kz), int(ky), int(kx), _lib.stream_ptr()), what)
    return out


def _nms_xy(heat, kernel=3):
    """image.py:81-87."""
    return _nms_window(heat, 1, kernel, kernel, "_nms_xy")


def _nms_z(heat, kernel=3):
    """image.py:89-95."""
    return _nms_window(heat, kernel, 1, 1, "_nms_z")


def _nms(heat, kernel=3):
    """image.py:97-105: a (k,k,k) window, unlike models/decode.py's (3,k,k)."""
    return _nms_window(heat, kernel, kernel, kernel, "_nms")


def non_maximum_suppression_3d(x, d, scale=1.0, threshold=float("-inf"), max_candidates=None):
    """image.py:42-79 for a float64 (or float32) volume: -> numpy (scores float32 [j], coords int32 [j,3] = x,y,z)."""
    t, _ = _to_device(x)
    if t.dim() != 3:
        raise ValueError(f"non_maximum_suppression_3d: expected a (D,H,W) volume, got {tuple(t.shape)}")
    if t.dtype != torch.float64:
        from ..models.decode import non_maximum_suppression_3d as nms32
        return nms32(t, d, scale=scale, threshold=threshold, max_candidates=max_candidates)
    t = t.contiguous()
    D, H, W = t.shape
    n = D * H * W
    thr = float(threshold)
    cap = int(max_candidates) if max_candidates else int(min(n, max(1 << 20, int((t > thr).sum().item()))))
    L = _lib.lib()
    nbytes = C.c_size_t(0)
    _lib.check(L.cetpick_greedy_nms_f64_workspace_bytes(D, H, W, cap, C.byref(nbytes)), "non_maximum_suppression_3d")
    ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=t.device)
    ws_ptr = (ws.data_ptr() + 255) // 256 * 256
    scores = torch.empty(cap, dtype=torch.float32, device=t.device)
    coords = torch.empty((cap, 3), dtype=torch.int32, device=t.device)
    n_out, rounds = C.c_int64(0), C.c_int(0)
    _lib.check(L.cetpick_greedy_nms_f64(t.data_ptr(), D, H, W, float(d), float(scale), thr, cap, scores.data_ptr(),
                                        coords.data_ptr(), cap, C.byref(n_out), C.byref(rounds), ws_ptr,
                                        ws.numel() - (ws_ptr - ws.data_ptr()), _lib.stream_ptr()),
               "non_maximum_suppression_3d")
    j = min(n_out.value, cap)
    return scores[:j].cpu().numpy(), coords[:j].cpu().numpy()


def get_potential_coords_pyramid(rec, sigmas=[2, 4], num_pyramid=3, kernel=3):
    """image.py:138-183: difference-of-Gaussians candidate generator of the exploration step.
    rec: (z, r, c) volume (numpy or tensor; processed in float64).  -> (scores float32 [n], coords int32 [n,3])."""
    rec, _ = _to_device(rec)
    rec = rec.to(torch.float64)
    z, r, c = rec.shape
    bound_x, bound_y = 30, 30
    if r > 512 and c > 512:
        bound_x, bound_y = bound_x * 2, bound_y * 2
    num_pyramid = len(sigmas)
    ims = [gaussian_filter(rec, sigmas[i]) for i in range(num_pyramid)]
    nms_all = None
    for i in range(num_pyramid - 1):
        diff = ims[i + 1] - ims[i]
        diff[:10, :, :] = 0
        diff[-10:, :, :] = 0
        diff[:, :bound_x, :] = 0
        diff[:, -bound_x:, :] = 0
        diff[:, :, :bound_y] = 0
        diff[:, :, -bound_y:] = 0
        nms_xy = _nms_xy(diff[None, None], kernel=kernel)[0, 0]
        nms_all = nms_xy if nms_all is None else torch.maximum(nms_all, nms_xy)     # np.max over the stack
    pos = nms_all[nms_all > 0]
    mean_nms = pos.mean().item()
    std_nms_half = pos.std().item()                 # unbiased, like torch.Tensor.std
    cutoff_score = mean_nms + std_nms_half * 0.5
    return non_maximum_suppression_3d(nms_all, 14, threshold=cutoff_score)


def extract_subvols(v, tomo_coords, subvol_size):
    """datasets/tomo_pre_proj_angle_select_new3d_vol.py:117-128, batched: for every candidate (x, y, z) the z-slab
    sum of the window, min-max normalised -> float32 CUDA tensor (n, 1, 2*(sub_y//2), 2*(sub_x//2)) (the reference
    returns one (1, sub_y, sub_x) tensor per call).  v: (D,H,W) volume, processed in float64."""
    v, _ = _to_device(v)
    v = v.to(torch.float64).contiguous()
    D, H, W = v.shape
    c = torch.as_tensor(np.asarray(tomo_coords).reshape(-1, 3), dtype=torch.int32)
    sub_z, sub_y, sub_x = (int(s) for s in subvol_size)
    hz, hy, hx = sub_z // 2, sub_y // 2, sub_x // 2
    if len(c):
        x, y, z = c[:, 0], c[:, 1], c[:, 2]
        ok = (x - hx >= 0) & (x + hx <= W) & (y - hy >= 0) & (y + hy <= H) & (z - hz >= 0) & (z < D)
        if not bool(ok.all()):
            raise ValueError("extract_subvols: a window leaves the volume (the reference filters such candidates, :207)")
    out = torch.empty((len(c), 1, 2 * hy, 2 * hx), dtype=torch.float32, device=v.device)
    cd = c.to(v.device).contiguous()
    _lib.check(_lib.lib().cetpick_extract_subvols_f64(v.data_ptr(), D, H, W, cd.data_ptr(), len(c), sub_z, sub_y, sub_x,
                                                      out.data_ptr(), _lib.stream_ptr()), "extract_subvols")
    return out
