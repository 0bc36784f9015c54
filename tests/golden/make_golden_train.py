"""Golden fixture for the training-step losses (cet_pick/models/loss.py), values and gradients from the REAL
reference on seeded inputs.

    python tests/golden/make_golden_train.py        # needs /root/reference (or $CET_PICK_REF)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import synthdata as synth                     # noqa: E402
from oracle import refbridge                        # noqa: E402

refbridge.install()
from cet_pick.models.loss import ConsistencyLoss, FocalLoss, PULoss        # noqa: E402

D, H, W = 6, 20, 24
u = synth.uniform_np(9, D * H * W).reshape(1, D, H, W)
pred0 = (0.02 + 0.96 * u).astype(np.float32)                               # sigmoid-clamped range
gt = -np.ones((1, D, H, W), np.float32)                                    # unlabelled everywhere ...
for (z, y, x) in [(2, 5, 6), (3, 12, 17), (4, 9, 10)]:                     # ... Gaussian bumps around 3 positives
    zz, yy, xx = np.mgrid[0:D, 0:H, 0:W]
    g = np.exp(-((zz - z) ** 2 + (yy - y) ** 2 + (xx - x) ** 2) / (2 * 1.5 ** 2)).astype(np.float32)
    m = g > 0.05
    gt[0][m] = np.maximum(np.where(gt[0][m] < 0, 0, gt[0][m]), g[m])
    gt[0, z, y, x] = 1.0
out = {"gt": gt, "seed": 9, "shape": np.array([D, H, W])}
for tag, fn in {"pu_tau01": lambda p, g_: PULoss(0.1)(p, g_), "pu_tau06": lambda p, g_: PULoss(0.6)(p, g_),
                "focal": lambda p, g_: FocalLoss()(p[0], g_[0])}.items():
    p = torch.from_numpy(pred0.copy()).requires_grad_(True)
    loss = fn(p, torch.from_numpy(gt))
    loss.backward()
    out[tag + "_loss"] = np.float32(loss.item())
    out[tag + "_grad"] = p.grad.numpy().copy()
    print(tag, loss.item(), float(np.abs(p.grad.numpy()).max()))
a = torch.from_numpy(pred0.copy()).requires_grad_(True)
b = torch.from_numpy((pred0[:, :, :, ::-1]).copy())
loss = ConsistencyLoss()(a, b)
loss.backward()
out["cons_loss"], out["cons_grad"] = np.float32(loss.item()), a.grad.numpy().copy()
np.savez_compressed(os.path.join(HERE, "train_losses.npz"), **out)
print("wrote train_losses")
