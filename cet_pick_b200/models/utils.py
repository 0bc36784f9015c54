"""Mirror of cet_pick/models/utils.py for the helpers on the hot path (:167-193)."""
from __future__ import annotations

import torch

from .. import _lib


def _sigmoid(x: torch.Tensor) -> torch.Tensor:
    """models/utils.py:167-169: clamp(x.sigmoid_(), 1e-4, 1-1e-4); mutates and returns `x`."""
    _lib.require_cuda(x, "_sigmoid")
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("_sigmoid expects a contiguous float32 tensor")
    _lib.check(_lib.lib().cetpick_sigmoid_clamp_f32(x.data_ptr(), x.numel(), _lib.stream_ptr()), "_sigmoid")
    return x


def _gather_feat(feat, ind, mask=None):
    """models/utils.py:171-181: rows `ind` (N, K) of feat (N, M, C) -> (N, K, C); with a boolean `mask` (N, K)
    only the selected rows, flattened to (n_selected, C)."""
    picked = torch.take_along_dim(feat, ind.unsqueeze(-1), dim=1)
    return picked if mask is None else picked[mask.bool()]


def _transpose_and_gather_feat(feat, ind):
    """models/utils.py:186-193: (N, C, D, H, W) feature map sampled at flat voxel indices `ind` (N, K) -> (N, K, C).
    tomo_decode itself does not need this (the CUDA pick writer gathers `reg` at the linear index); kept for API
    parity."""
    n, c = feat.shape[:2]
    return _gather_feat(feat.reshape(n, c, -1).transpose(1, 2), ind)
