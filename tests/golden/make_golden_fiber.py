"""Supplementary decode fixtures from the REAL reference: `--fiber` with NMS windows 5 and 7
(cet_pick/models/decode.py:126-128: (1,k,k) suppression, then (k,1,1) suppression of the result).

    python tests/golden/make_golden_fiber.py      # needs /root/reference (or $CET_PICK_REF)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import synthdata as synth          # noqa: E402
from oracle import refbridge             # noqa: E402

d = refbridge.decode_module()
for k, shape, seed, K in ((5, (9, 21, 27), 14, 25), (7, (11, 24, 30), 15, 18)):
    hm = synth.heatmap_tiefree_np(*shape, seed)[None, None]
    out = d.tomo_decode(torch.from_numpy(hm), kernel=k, reg=None, K=K, if_fiber=True).numpy()
    # rows past the last surviving peak are zero-score filler whose indices torch.topk leaves unspecified
    n_pos = int((out[0, :, 3] > 0).sum())
    np.savez_compressed(os.path.join(HERE, f"decode_fiber_k{k}.npz"), dets=out, kernel=k, K=K, fiber=1,
                        shape=shape, seed=seed, n_pos=n_pos)
    print("wrote decode_fiber_k%d" % k, out.shape, "positive rows", n_pos)
