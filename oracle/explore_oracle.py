"""TEST INFRASTRUCTURE (CPU restatement, numpy / scipy / torch-CPU): the exploration-step candidate generator of
cet_pick/utils/image.py (`_nms_xy` :81-87, `get_potential_coords_pyramid` :138-183).  Only tests/ may import this.
Pinned by tests/golden/explore_pyramid.npz, produced by the unmodified reference (make_golden_explore.py)."""
from __future__ import annotations

import numpy as np
import torch
from scipy.ndimage import gaussian_filter

from .decode_oracle import greedy_distance_nms


def nms_window(heat: np.ndarray, kz: int, ky: int, kx: int) -> np.ndarray:
    """heat * (max_pool3d(heat, (kz,ky,kx), 1, same) == heat) for a (D,H,W) float32/float64 array."""
    t = torch.from_numpy(np.ascontiguousarray(heat))[None, None]
    hmax = torch.nn.functional.max_pool3d(t, (kz, ky, kx), stride=1, padding=((kz - 1) // 2, (ky - 1) // 2, (kx - 1) // 2))
    return (t * (hmax == t).float())[0, 0].numpy()


def get_potential_coords_pyramid(rec: np.ndarray, sigmas=(2, 4), kernel=3):
    """image.py:138-183."""
    z, r, c = rec.shape
    bx = by = 60 if (r > 512 and c > 512) else 30
    ims = [gaussian_filter(rec, s) for s in sigmas]
    alls = []
    for i in range(len(sigmas) - 1):
        diff = ims[i + 1] - ims[i]
        diff[:10] = 0
        diff[-10:] = 0
        diff[:, :bx] = 0
        diff[:, -bx:] = 0
        diff[:, :, :by] = 0
        diff[:, :, -by:] = 0
        alls.append(nms_window(diff, 1, kernel, kernel))
    nms = np.max(np.stack(alls, axis=0), axis=0)
    t = torch.as_tensor(nms)
    pos = t[torch.where(t > 0)]
    cutoff = pos.mean().item() + pos.std().item() * 0.5
    return greedy_distance_nms(nms, 14, threshold=cutoff), nms, cutoff
