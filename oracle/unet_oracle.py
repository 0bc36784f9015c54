"""torch-fp32 restatement of the default detector forward (TEST INFRASTRUCTURE, see
oracle/__init__.py).  Floating-point path => a plain PyTorch fp32 reference is the oracle.

Follows cet_pick/models/networks/unet_small.py:63-97 (TomoConvUNet.forward) and
cet_pick/models/networks/unet.py:198-249 (DownConv), :252-316 (autocrop), :319-399 (UpConv),
:861-886 (UNet.forward).  It consumes the reference state_dict keys directly (SURVEY.md
Appendix A) and contains no nn.Module, so it does not depend on the reference at run time.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-5  # nn.BatchNorm2d default, eval mode (running statistics)


def _bn(x, sd, p, train=False):
    """eval: running statistics (inference path).  train: batch statistics, running statistics updated IN PLACE with
    momentum 0.1 (nn.BatchNorm2d defaults) -- the training step of trains/base_trainer.py:135-155."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], train, 0.1 if train else 0.0, EPS)


def n_blocks_of(sd) -> int:
    n = 0
    while f"unet.down_convs.{n}.conv1.weight" in sd:
        n += 1
    return n


def unet_trunk(x, sd, train=False):
    """x: (D,16,h,w) -> (D,32,h,w); UNet(16, 32, n_blocks, dim=2, 'concat', 'transpose', 'same')."""
    nb = n_blocks_of(sd)
    enc = []
    for i in range(nb):
        p = f"unet.down_convs.{i}"
        y = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], padding=1), sd, p + ".norm0", train))
        y = F.relu(_bn(F.conv2d(y, sd[p + ".conv2.weight"], padding=1), sd, p + ".norm1", train))
        enc.append(y)
        x = F.max_pool2d(y, 2, ceil_mode=True) if i < nb - 1 else y      # unet.py:225
    for i in range(nb - 1):
        p = f"unet.up_convs.{i}"
        skip = enc[-(i + 2)]
        up = F.conv_transpose2d(x, sd[p + ".upconv.weight"], sd[p + ".upconv.bias"], stride=2)
        # autocrop step 1 (unet.py:285-292): crop the upsampled map where the skip is odd-sized
        up = up[:, :, :skip.shape[2], :skip.shape[3]]
        up = F.relu(_bn(up, sd, p + ".norm0", train))
        m = torch.cat((up, skip), 1)                                       # unet.py:390
        y = F.relu(_bn(F.conv2d(m, sd[p + ".conv1.weight"], padding=1), sd, p + ".norm1", train))
        x = F.relu(_bn(F.conv2d(y, sd[p + ".conv2.weight"], padding=1), sd, p + ".norm2", train))
    return F.conv2d(x, sd["unet.conv_final.weight"], sd["unet.conv_final.bias"])


def forward(x, sd, want_proj: bool = True, train: bool = False):
    """x: (1,D,H,W) float32 -> {'hm': (1,1,D,h,w), 'proj': (1,C,D,h,w)} raw (pre-sigmoid) outputs."""
    if x.dim() > 4:
        x = x.squeeze()
    b, d, h, w = x.shape
    if b > 1:                                                              # unet_small.py:67-69,79-81
        x = x.reshape((-1, h, w)).unsqueeze(1)
    else:
        x = x.permute(1, 0, 2, 3)
    x = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), sd, "bn1", train))
    x = unet_trunk(x, sd, train)
    if b > 1:
        x = x.reshape((b, d) + tuple(x.shape[1:])).permute(0, 2, 1, 3, 4)
    else:
        x = x.permute(1, 0, 2, 3).unsqueeze(0)
    x = F.relu(F.conv3d(x, sd["feature_head.0.weight"], padding=(1, 4, 4), dilation=(1, 4, 4)))
    x = F.relu(F.conv3d(x, sd["feature_head.2.weight"], padding=(1, 4, 4), dilation=(1, 4, 4)))
    ret = {"hm": F.conv3d(x, sd["hm.weight"], padding=(1, 0, 0))}
    if want_proj and "proj.weight" in sd:
        ret["proj"] = F.normalize(F.conv3d(x, sd["proj.weight"], padding=(1, 0, 0)), dim=1)
    return ret


def sigmoid_clamp(hm):
    """models/utils.py:167-169 `_sigmoid`."""
    return torch.clamp(torch.sigmoid(hm), min=1e-4, max=1 - 1e-4)
