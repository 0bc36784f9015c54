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
    """models/utils.py:171-181 (index glue; plain tensor indexing, no arithmetic)."""
    dim = feat.size(2)
    ind = ind.unsqueeze(2).expand(ind.size(0), ind.size(1), dim)
    feat = feat.gather(1, ind)
    if mask is not None:
        mask = mask.unsqueeze(2).expand_as(feat)
        feat = feat[mask].view(-1, dim)
    return feat


def _transpose_and_gather_feat(feat, ind):
    """models/utils.py:186-193.  tomo_decode does not use this (the CUDA pick writer gathers `reg`
    directly at the linear index); kept for API parity."""
    feat = feat.permute(0, 2, 3, 4, 1).contiguous()
    feat = feat.view(feat.size(0), -1, feat.size(4))
    return _gather_feat(feat, ind)
