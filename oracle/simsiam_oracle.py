"""TEST INFRASTRUCTURE (CPU restatement, functional torch fp32) of the exploration-step embedding network
`TomoResClassifier.forward_test` (cet_pick/models/networks/simsiam_model.py:325-366; blocks :44-73, stages :256-271,
3-D feature layer and heads :181-215), i.e. BASELINE.json configs[3] (`simsiam3d_18`).  Groundwork for SURVEY 8f-3's
second half.  Pinned by tests/golden/simsiam3d_small.npz (unmodified reference).  `forward_test_2d` restates the 2-D
exploration variant `TomoResClassifier2D.forward_test` (cet_pick/models/networks/simsiam_model_2d.py:751-774; layers
:617-664, arch `simsiam2d_18`), pinned by tests/golden/simsiam2d_small.npz."""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-5


def _bn(x, sd, p, affine=True):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd.get(p + ".weight") if affine else None,
                        sd.get(p + ".bias") if affine else None, False, 0.0, EPS)


def _block(x, sd, p, stride):
    """BasicBlock (:44-73): conv3x3(stride)-BN-ReLU-conv3x3-BN, + (1x1 strided conv of x | x), ReLU."""
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1), sd, p + ".bn1"))
    out = _bn(F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1), sd, p + ".bn2")
    if (p + ".downsample.0.weight") in sd:
        x = F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0)      # no BN on the shortcut (:259-263)
    return F.relu(out + x)


def forward_test(x, sd, layers=(2, 2, 2), heads=("proj", "pred")):
    """x: (B, D, H, W) sub-volumes -> {'proj': (B,256), 'pred': (B,256)} (eval-mode statistics)."""
    if x.dim() > 4:
        x = x.squeeze(1)
    b, d, h, w = x.shape
    y = x.reshape(-1, 1, h, w)                                   # every slice through the 2-D trunk (:333-337)
    y = F.relu(_bn(F.conv2d(y, sd["conv1.weight"], None, 2, 3), sd, "bn1"))
    y = F.max_pool2d(y, 3, 2, 1)
    for li, nblk in enumerate(layers, start=1):
        for k in range(nblk):
            y = _block(y, sd, f"layer{li}.{k}", 2 if (li > 1 and k == 0) else 1)
    _, ch, hh, ww = y.shape
    y = y.reshape(b, d, ch, hh, ww).permute(0, 2, 1, 3, 4)       # (B, C, D, h, w) (:348-353)
    y = F.relu(_bn(F.conv3d(y, sd["feature_3d.0.weight"], None, 1, 1), sd, "feature_3d.1"))
    y = y.mean(dim=(2, 3, 4))                                    # AdaptiveAvgPool3d((1,1,1)) + flatten
    y = F.linear(y, sd["fc.weight"], sd["fc.bias"])
    out = {}
    z = None
    if "proj" in heads:
        z = F.relu(_bn(F.linear(y, sd["proj.0.weight"]), sd, "proj.1"))
        z = F.relu(_bn(F.linear(z, sd["proj.3.weight"]), sd, "proj.4"))
        z = _bn(F.linear(z, sd["proj.6.weight"]), sd, "proj.7", affine=False)
        out["proj"] = z
    if "pred" in heads:
        p = F.relu(_bn(F.linear(z, sd["pred.0.weight"]), sd, "pred.1"))
        out["pred"] = F.linear(p, sd["pred.3.weight"], sd["pred.3.bias"])
    return out


def forward_test_2d(x, sd, layers=(2, 2, 2), heads=("proj", "pred")):
    """x: (B, 1, H, W) patches -> {'proj': (B,out_dim), 'pred': (B,out_dim)} (simsiam_model_2d.py:751-774)."""
    if x.dim() > 4:
        x = x.squeeze(1)
    y = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], None, 1, 1), sd, "bn1"))        # 3x3 stride 1, no max-pool (:625-628)
    for li, nblk in enumerate(layers, start=1):
        for k in range(nblk):
            y = _block(y, sd, f"layer{li}.{k}", 2 if (li > 1 and k == 0) else 1)
    y = y.mean(dim=(2, 3))                                                         # AdaptiveAvgPool2d((1,1)) + flatten
    y = F.linear(y, sd["fc.weight"], sd["fc.bias"])
    out = {}
    z = None
    if "proj" in heads:
        z = F.relu(_bn(F.linear(y, sd["proj.0.weight"]), sd, "proj.1"))
        z = F.relu(_bn(F.linear(z, sd["proj.3.weight"]), sd, "proj.4"))
        z = _bn(F.linear(z, sd["proj.6.weight"]), sd, "proj.7", affine=False)
        out["proj"] = z
    if "pred" in heads:
        p = F.relu(_bn(F.linear(z, sd["pred.0.weight"]), sd, "pred.1"))
        out["pred"] = F.linear(p, sd["pred.3.weight"], sd["pred.3.bias"])
    return out
