"""Mirror of cet_pick/utils/post_process.py:11-25 (`tomo_post_process`): host-side grouping of the
(1,K,5) pick rows by integer z.  The fiber / spike graph post-processing (:31-106) is out of scope
(CPU, <= K points, option-gated)."""
from __future__ import annotations

import numpy as np


def tomo_post_process(dets, z_dim_tot=128):
    """dets: (B, K, 5) numpy.  Returns [ {z: [[x, y, z, score, score], ...]} ] for the LAST batch
    element only, rows in top-K order, exactly like the reference (its ret.append is outside the
    batch loop)."""
    top_preds = {}
    for i in range(dets.shape[0]):
        top_preds = {}
        z = dets[i, :, 2]
        order = np.argsort(z, kind="stable")          # rows of one plane stay in top-K order (stable)
        zs = z[order]
        ok = (zs >= 0) & (zs < z_dim_tot) & (zs == np.floor(zs))     # integer plane index in [0, z_dim_tot)
        order, zs = order[ok], zs[ok]
        if zs.size == 0:
            continue
        rows = dets[i, order, :].astype(np.float32).tolist()        # one conversion for all K rows (host-side cost of run())
        starts = np.flatnonzero(np.r_[True, zs[1:] != zs[:-1]]).tolist() + [int(zs.size)]
        keys = zs[starts[:-1]].astype(np.int64).tolist()
        for k, a, b in zip(keys, starts[:-1], starts[1:]):
            top_preds[k] = rows[a:b]
    return [top_preds]
