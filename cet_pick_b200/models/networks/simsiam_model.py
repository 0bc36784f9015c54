"""Host-side mirror of cet_pick/models/networks/simsiam_model.py (`TomoResClassifier`, arch `simsiam_18` /
`simsiam3d_18`), the exploration-step embedding network simsiam_test_hm_3d.py:136-195 runs over the candidate
sub-volumes.  The nn.Module holds the reference's parameter names, shapes and registration order (state_dicts are
interchangeable); `forward_test` runs entirely in libcetpick_sm100a.so (csrc/simsiam.cu, csrc/conv_small.cu).
There is no PyTorch/CPU forward path; the two-view training forward (:368-439) is outside this build."""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from ... import _lib

BN_MOMENTUM = 0.1


class _BasicBlockParams(nn.Module):
    """simsiam_model.py:44-73 BasicBlock (expansion 1): conv3x3(stride)-BN-ReLU-conv3x3-BN (+ shortcut), ReLU."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        if downsample is not None:
            self.downsample = downsample


def _fill_fc_weights(layers):
    """simsiam_model.py:141-156: N(0, 0.001) weights, zero biases for conv / linear layers"""
    for m in layers.modules():
        if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.Linear)):
            nn.init.normal_(m.weight, std=0.001)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)


class TomoResClassifier(nn.Module):
    """simsiam_model.py:159-235.  forward_test(x: (B, D, H, W) or (B, 1, D, H, W)) -> {'proj': (B,256), 'pred': (B,256)}."""

    def __init__(self, layers, heads, head_conv=0):
        super().__init__()
        self.heads = heads
        self.layers_spec = list(layers[:3])
        self.inplanes = 64
        self.conv1 = nn.Conv2d(1, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.layer1 = self._make_layer(64, layers[0])
        self.layer2 = self._make_layer(128, layers[1], stride=2)
        self.layer3 = self._make_layer(256, layers[2], stride=2)
        self.feature_3d = nn.Sequential(nn.Conv3d(256, 256, kernel_size=3, padding=1, bias=False),
                                        nn.BatchNorm3d(256, momentum=BN_MOMENTUM), nn.Identity())
        _fill_fc_weights(self.feature_3d)
        self.fc = nn.Linear(256, 256)
        _fill_fc_weights(self.fc)
        for head in self.heads:
            if "proj" in head:
                fc = nn.Sequential(nn.Linear(256, 256, bias=False), nn.BatchNorm1d(256), nn.Identity(),
                                   nn.Linear(256, 256, bias=False), nn.BatchNorm1d(256), nn.Identity(),
                                   nn.Linear(256, 256, bias=False), nn.BatchNorm1d(256, affine=False))
            elif "pred" in head:
                fc = nn.Sequential(nn.Linear(256, 256, bias=False), nn.BatchNorm1d(256), nn.Identity(),
                                   nn.Linear(256, 256))
            else:
                raise NotImplementedError(f"TomoResClassifier head {head!r} (the reference builds 'proj' / 'pred')")
            _fill_fc_weights(fc)
            setattr(self, head, fc)
        self._plan = None
        self._plan_key = None
        self._ws = None

    def _make_layer(self, planes, blocks, stride=1):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False))
        layers = [_BasicBlockParams(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(_BasicBlockParams(self.inplanes, planes))
        return nn.Sequential(*layers)

    # ------------------------------------------------------------------ plan management
    def _destroy_plan(self):
        if self._plan is not None:
            _lib.lib().cetpick_simsiam_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self._destroy_plan()
        except Exception:
            pass

    def plan(self):
        key = tuple((k, v.data_ptr(), v._version) for k, v in self.state_dict().items())
        if self._plan is not None and key == self._plan_key:
            return self._plan
        self._destroy_plan()
        L = _lib.lib()
        h = C.c_void_p()
        has_proj = any("proj" in k for k in self.heads)
        has_pred = any("pred" in k for k in self.heads)
        self._create_plan(L, h, int(has_proj), int(has_pred))
        for k, v in self.state_dict().items():
            if k.endswith("num_batches_tracked"):
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            _lib.check(L.cetpick_simsiam_set_param(h, k.encode(), t.data_ptr(), t.numel()), f"set_param({k})")
        _lib.check(L.cetpick_simsiam_finalize(h), "cetpick_simsiam_finalize")
        self._plan, self._plan_key = h, key
        return h

    out_dim = 256

    def _create_plan(self, L, h, has_proj, has_pred):
        _lib.check(L.cetpick_simsiam_create(C.byref(h), *self.layers_spec, has_proj, has_pred), "cetpick_simsiam_create")

    # ------------------------------------------------------------------ forward
    def forward_test(self, x1):
        """simsiam_model.py:325-366 (eval-mode statistics)."""
        _lib.require_cuda(x1, "TomoResClassifier.forward_test")
        if x1.dim() > 4:
            x1 = x1.squeeze(dim=1)
        b, d, h, w = x1.shape
        x1 = x1.to(torch.float32).contiguous()
        plan = self.plan()
        L = _lib.lib()
        nbytes = C.c_size_t(0)
        _lib.check(L.cetpick_simsiam_workspace_bytes(plan, b, d, h, w, C.byref(nbytes)), "cetpick_simsiam_workspace_bytes")
        if self._ws is None or self._ws.numel() < nbytes.value or self._ws.device != x1.device:
            self._ws = None
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x1.device)
        ret = {}
        proj = pred = None
        for head in self.heads:
            if "proj" in head:
                proj = ret[head] = torch.empty((b, self.out_dim), dtype=torch.float32, device=x1.device)
            if "pred" in head:
                pred = ret[head] = torch.empty((b, self.out_dim), dtype=torch.float32, device=x1.device)
        _lib.check(L.cetpick_simsiam_forward(plan, x1.data_ptr(), b, d, h, w,
                                             proj.data_ptr() if proj is not None else None,
                                             pred.data_ptr() if pred is not None else None,
                                             self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr()),
                   "cetpick_simsiam_forward")
        self.last_launches = L.cetpick_last_launch_count()
        return ret

    def forward(self, x1, x2):
        raise NotImplementedError("TomoResClassifier.forward (two-view SimSiam training, simsiam_model.py:368-439) is "
                                  "outside cet_pick_b200; inference uses forward_test")


_RESNET_SPEC = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3]}


def get_simsiam_net_small(num_layers, heads, head_conv=32, last_k=0, local_path=None):
    """simsiam_model.py:517-523.  The reference then overwrites the trunk with ImageNet ResNet weights from `local_path`
    (init_weights, :464-509); here a checkpoint is always loaded afterwards (load_model), so only `local_path` files
    that exist are applied."""
    if num_layers not in _RESNET_SPEC:
        raise NotImplementedError(f"simsiam_{num_layers}: only BasicBlock ResNets (18, 34) are built")
    model = TomoResClassifier(_RESNET_SPEC[num_layers], heads, head_conv=0)
    if local_path:
        sd = torch.load(local_path, map_location="cpu")
        if "conv1.weight" in sd and sd["conv1.weight"].shape[1] == 3:
            sd["conv1.weight"] = sd["conv1.weight"].sum(dim=1, keepdim=True)      # _load_pretrained(inchans=1)
        own = model.state_dict()
        model.load_state_dict({k: v for k, v in sd.items() if k in own and own[k].shape == v.shape}, strict=False)
    return model
