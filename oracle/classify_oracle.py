"""TEST INFRASTRUCTURE (CPU restatement, numpy): the semiclass tile scheduler of
cet_pick/detectors/tomo_det_classify.py.  Only tests/ may import this.  Pinned by
tests/golden/patch_dataset.npz and classify_{tiled,whole}.npz, which were produced by the unmodified
reference (tests/golden/make_golden_classify.py)."""
from __future__ import annotations

import numpy as np

from .decode_oracle import greedy_distance_nms, sigmoid_clamp


def patch(tomo: np.ndarray, n: int, psz: int, psxy: int, pdz: int, pdxy: int):
    """tomo_det_classify.py:37-75 `PatchDataset.__getitem__`: (index (3,), zero-padded tile)."""
    nz, ny, nx = tomo.shape
    grid = (int(np.ceil(nz / psz)), int(np.ceil(ny / psxy)), int(np.ceil(nx / psxy)))
    i, j, k = np.unravel_index(n, grid)
    i, j, k = psz * int(i), psxy * int(j), psxy * int(k)
    x = np.zeros((psz + 2 * pdz, psxy + 2 * pdxy, psxy + 2 * pdxy), dtype=np.float32)
    si, ei = max(0, i - pdz), min(nz, i + psz + pdz)
    sj, ej = max(0, j - pdxy), min(ny, j + psxy + pdxy)
    sk, ek = max(0, k - pdxy), min(nx, k + psxy + pdxy)
    sic, sjc, skc = pdz - i + si, pdxy - j + sj, pdxy - k + sk
    x[sic:sic + ei - si, sjc:sjc + ej - sj, skc:skc + ek - sk] = tomo[si:ei, sj:ej, sk:ek]
    return np.array((i, j, k)), x, grid


def stub_model(x: np.ndarray) -> np.ndarray:
    """numpy twin of synthdata.fullres_stub_model (separately rounded fp32 operations)."""
    xp = np.pad(x, 1)
    y = (x * np.float32(6.0) - np.float32(3.0)).astype(np.float32)
    y = (y + np.float32(0.5) * xp[:-2, 1:-1, 1:-1]).astype(np.float32)
    y = (y + np.float32(0.25) * xp[1:-1, 2:, 1:-1]).astype(np.float32)
    return y


def process(vol: np.ndarray, nms: float, out_thresh: float):
    """tomo_det_classify.py:82-156 with the stub model: heat-map (D,H,W) and picks (n,4) [x,y,z,score]."""
    D, H, W = vol.shape
    if D <= 128 and H <= 128:          # :85 reads shape[1], shape[2] of the (1,D,H,W) batch
        hm = sigmoid_clamp(stub_model(vol))
    else:
        psz, psxy, pdz, pdxy = 32, 96, 16, 24
        hm = np.zeros_like(vol)
        n_tiles = int(np.ceil(D / psz)) * int(np.ceil(H / psxy)) * int(np.ceil(W / psxy))
        for n in range(n_tiles):
            (i, j, k), x, _ = patch(vol, n, psz, psxy, pdz, pdxy)
            xb = sigmoid_clamp(stub_model(x))
            pz, py, px = hm[i:i + psz, j:j + psxy, k:k + psxy].shape
            hm[i:i + psz, j:j + psxy, k:k + psxy] = xb[pdz:pdz + pz, pdxy:pdxy + py, pdxy:pdxy + px]
    hm[:, :30, :] = 0
    hm[:, -30:, :] = 0
    hm[:, :, :30] = 0
    hm[:, :, -30:] = 0
    sc, co = greedy_distance_nms(hm, nms, threshold=out_thresh)
    return hm, np.concatenate([co.astype(np.float32), sc[:, None]], axis=1)
