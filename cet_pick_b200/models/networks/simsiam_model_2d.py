"""Host-side mirror of cet_pick/models/networks/simsiam_model_2d.py:617-774 (`TomoResClassifier2D`, arch
`simsiam2d_18`): the exploration-step embedding network for 2-D patches.  Against the 3-D classifier
(simsiam_model.py) conv1 is a 3x3 stride-1 convolution without a max-pool, there is no Conv3d feature layer, the pool
is AdaptiveAvgPool2d and `fc` / the heads have width `head_conv`.  The nn.Module holds the reference's parameter names,
shapes and registration order; `forward_test` runs in libcetpick_sm100a.so (csrc/simsiam.cu `two_d` plans)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from ... import _lib
from .simsiam_model import BN_MOMENTUM, _RESNET_SPEC, TomoResClassifier, _fill_fc_weights


class TomoResClassifier2D(TomoResClassifier):
    """simsiam_model_2d.py:617-664.  forward_test(x: (B, 1, H, W)) -> {'proj': (B, head_conv), 'pred': (B, head_conv)}."""

    def __init__(self, layers, heads, head_conv):
        nn.Module.__init__(self)
        if not 1 <= head_conv <= 256:
            raise NotImplementedError(f"TomoResClassifier2D head_conv={head_conv}: widths 1 ... 256 are built (the "
                                      "reference's default for the exploration task is 128, opts.py:207-209)")
        self.heads = heads
        self.layers_spec = list(layers[:3])
        self.inplanes = 64
        self.conv1 = nn.Conv2d(1, 64, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.layer1 = self._make_layer(64, layers[0])
        self.layer2 = self._make_layer(128, layers[1], stride=2)
        self.layer3 = self._make_layer(256, layers[2], stride=2)
        self.out_dim = d = head_conv
        self.fc = nn.Linear(256, d)
        _fill_fc_weights(self.fc)
        for head in self.heads:
            if "proj" in head:
                fc = nn.Sequential(nn.Linear(d, d, bias=False), nn.BatchNorm1d(d), nn.Identity(),
                                   nn.Linear(d, d, bias=False), nn.BatchNorm1d(d), nn.Identity(),
                                   nn.Linear(d, d, bias=False), nn.BatchNorm1d(d, affine=False))
            elif "pred" in head:
                fc = nn.Sequential(nn.Linear(d, d, bias=False), nn.BatchNorm1d(d), nn.Identity(), nn.Linear(d, d))
            else:
                raise NotImplementedError(f"TomoResClassifier2D head {head!r} (the reference builds 'proj' / 'pred')")
            _fill_fc_weights(fc)
            setattr(self, head, fc)
        self._plan = None
        self._plan_key = None
        self._ws = None

    def _create_plan(self, L, h, has_proj, has_pred):
        _lib.check(L.cetpick_simsiam_create_2d(C.byref(h), *self.layers_spec, self.out_dim, has_proj, has_pred),
                   "cetpick_simsiam_create_2d")

    def forward_test(self, x1):
        """simsiam_model_2d.py:751-774: x1 (B, 1, H, W) patches (a 5-D tensor loses its dim 1 first)."""
        if x1.dim() > 4:
            x1 = x1.squeeze(dim=1)
        if x1.dim() != 4 or x1.shape[1] != 1:
            raise RuntimeError(f"TomoResClassifier2D.forward_test expects (B, 1, H, W) patches, got {tuple(x1.shape)}")
        return super().forward_test(x1)


def get_simsiam2d_net_small(num_layers, heads, head_conv=32, last_k=0, local_path=None):
    """simsiam_model_2d.py:928-932 (`init_weights` loads ImageNet ResNet weights from `local_path` when given; here a
    checkpoint is always loaded afterwards, so only files that exist are applied)."""
    if num_layers not in _RESNET_SPEC:
        raise NotImplementedError(f"simsiam2d_{num_layers}: only BasicBlock ResNets (18, 34) are built")
    model = TomoResClassifier2D(_RESNET_SPEC[num_layers], heads, head_conv=head_conv)
    if local_path:
        sd = torch.load(local_path, map_location="cpu")
        own = model.state_dict()
        model.load_state_dict({k: v for k, v in sd.items() if k in own and own[k].shape == v.shape}, strict=False)
    return model
