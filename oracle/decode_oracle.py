"""numpy restatement of cet_pick/models/decode.py (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference lines it follows.  Tie order, which torch.topk leaves
unspecified, is fixed here to (score descending, linear index ascending); -0.0 == +0.0.
"""
from __future__ import annotations

import numpy as np


def _max_pool3d_same(heat: np.ndarray, kz: int, ky: int, kx: int) -> np.ndarray:
    """F.max_pool3d(heat, (kz,ky,kx), stride=1, padding=((kz-1)//2, ...)) with -inf padding.

    heat: (..., D, H, W) float32.  Odd kernel sizes only (even sizes change the output shape
    and make the reference's `hmax == heat` fail, decode.py:30-32).
    """
    pz, py, px = (kz - 1) // 2, (ky - 1) // 2, (kx - 1) // 2
    D, H, W = heat.shape[-3:]
    pad = [(0, 0)] * (heat.ndim - 3) + [(pz, pz), (py, py), (px, px)]
    p = np.pad(heat, pad, mode="constant", constant_values=-np.inf)
    out = np.full_like(heat, -np.inf)
    for dz in range(kz):
        for dy in range(ky):
            for dx in range(kx):
                out = np.maximum(out, p[..., dz:dz + D, dy:dy + H, dx:dx + W])
    return out


def nms(heat: np.ndarray, kernel: int = 3) -> np.ndarray:
    """decode.py:27-33 `_nms`: z extent is always 3, xy extent is `kernel`."""
    hmax = _max_pool3d_same(heat, 3, kernel, kernel)
    return heat * (hmax == heat).astype(np.float32)


def nms_xy(heat: np.ndarray, kernel: int = 3) -> np.ndarray:
    """decode.py:11-17 `_nms_xy`."""
    hmax = _max_pool3d_same(heat, 1, kernel, kernel)
    return heat * (hmax == heat).astype(np.float32)


def nms_z(heat: np.ndarray, kernel: int = 3) -> np.ndarray:
    """decode.py:19-25 `_nms_z`."""
    hmax = _max_pool3d_same(heat, kernel, 1, 1)
    return heat * (hmax == heat).astype(np.float32)


def convert_1d_to_3d(inds: np.ndarray, d: int, h: int, w: int):
    """decode.py:35-41 `_convert_1d_to_3d`, including its fp32 rounding for inds >= 2**24.

    z = int32(floor(float32(ind) / float32(h*w))); t = int32(ind) - z*h*w;
    y = floor(float32(t) / float32(w))  (stays float32);  x = t mod w (sign of divisor).
    """
    inds = np.asarray(inds, dtype=np.int64)
    z = np.floor(inds.astype(np.float32) / np.float32(h * w)).astype(np.int32)
    t = inds.astype(np.int32) - z * np.int32(h * w)
    y = np.floor(t.astype(np.float32) / np.float32(w)).astype(np.float32)
    x = np.mod(t, np.int32(w)).astype(np.int32)
    return z, y, x


def topk_canonical(flat: np.ndarray, K: int):
    """torch.topk(flat, K) (decode.py:84) with ties resolved by ascending index."""
    order = np.argsort(-flat, kind="stable")[:K]
    return flat[order], order.astype(np.int64)


def topk(scores: np.ndarray, K: int = 900):
    """decode.py:82-92 `_topk` on (B,C,D,H,W); C must be 1 like the reference's .view(batch,K)."""
    B, C, D, H, W = scores.shape
    assert C == 1
    flat = scores.reshape(B, -1)
    ts = np.empty((B, K), np.float32)
    ti = np.empty((B, K), np.int64)
    for b in range(B):
        ts[b], ti[b] = topk_canonical(flat[b], K)
    zs, ys, xs = convert_1d_to_3d(ti, D, H, W)
    return ts, zs, ys, xs, ti


def transpose_and_gather_feat(feat: np.ndarray, ind: np.ndarray) -> np.ndarray:
    """models/utils.py:171-193: (B,C,D,H,W) -> (B,DHW,C) then gather rows `ind` (B,K)."""
    B, C = feat.shape[:2]
    f = np.moveaxis(feat, 1, -1).reshape(B, -1, C)
    return np.stack([f[b, ind[b]] for b in range(B)], 0)


def tomo_decode(heat: np.ndarray, kernel: int = 3, reg=None, K: int = 900,
                if_fiber: bool = False) -> np.ndarray:
    """decode.py:123-155 `tomo_decode`: (B,1,D,H,W) float32 -> (B,K,5) float32
    rows [x+0.25, y+0.25, z, score, score] (or x+reg0, y+reg1 with `reg`)."""
    heat = np.asarray(heat, dtype=np.float32)
    B = heat.shape[0]
    if if_fiber:
        h = nms_z(nms_xy(heat, kernel), kernel)
    else:
        h = nms(heat, kernel)
    scores, zs, ys, xs, inds = topk(h, K)
    if reg is not None:
        r = transpose_and_gather_feat(np.asarray(reg, np.float32), inds)
        xs = xs.astype(np.float32) + r[:, :, 0]
        ys = ys + r[:, :, 1]
    else:
        xs = (xs.astype(np.float32) + np.float32(0.25)).astype(np.float32)
        ys = (ys + np.float32(0.25)).astype(np.float32)
    det = np.stack([xs.astype(np.float32), ys.astype(np.float32), zs.astype(np.float32),
                    scores, scores], axis=2)
    return det.reshape(B, K, 5).astype(np.float32)


def sigmoid_clamp(x: np.ndarray) -> np.ndarray:
    """models/utils.py:167-169 `_sigmoid` (out of place here): clamp(sigmoid(x), 1e-4, 1-1e-4)."""
    y = (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)
    return np.clip(y, np.float32(1e-4), np.float32(1 - 1e-4))


def tomo_post_process(dets: np.ndarray, z_dim_tot: int = 128):
    """utils/post_process.py:11-25: bucket rows by exact z == j; only the last batch element
    is returned (ret.append sits outside the batch loop)."""
    top_preds = {}
    for i in range(dets.shape[0]):
        top_preds = {}
        z = dets[i, :, 2]
        for j in range(z_dim_tot):
            m = z == j
            if m.sum() > 0:
                top_preds[j] = dets[i, m, :].astype(np.float32).tolist()
    return [top_preds]


def greedy_distance_nms(x: np.ndarray, d: float, scale: float = 1.0, threshold: float = -np.inf):
    """decode.py:42-79 `non_maximum_suppression_3d` with canonical visit order
    (score desc, index asc); flat-index deltas (wrap across rows) kept as in the reference."""
    r = scale * d / 2
    width = int(np.ceil(r))
    A = np.arange(-width, width + 1)
    ii, jj, kk = np.meshgrid(A, A, A)
    mask = (ii ** 2 + jj ** 2 + kk ** 2) <= r * r
    deltas = ii[mask] * (x.shape[1] * x.shape[2]) + jj[mask] * x.shape[2] + kk[mask]
    flat = x.ravel()
    order = np.argsort(-flat, kind="stable")
    sup = set()
    scores, coords = [], []
    for i in order:
        if flat[i] <= threshold:
            break
        if int(i) not in sup:
            zz, yy, xx = np.unravel_index(i, x.shape)
            scores.append(flat[i])
            coords.append((xx, yy, zz))
            for dl in deltas:
                sup.add(int(i + dl))
    return (np.asarray(scores, np.float32),
            np.asarray(coords, np.int32).reshape(-1, 3))
