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
        order = np.argsort(z, kind="stable")
        zs = z[order]
        # rows whose z is an integer plane index in [0, z_dim_tot): one pass instead of one mask per plane
        for j in np.unique(zs):
            if j < 0 or j >= z_dim_tot or j != np.floor(j):
                continue
            lo, hi = np.searchsorted(zs, j, "left"), np.searchsorted(zs, j, "right")
            top_preds[int(j)] = dets[i, np.sort(order[lo:hi]), :].astype(np.float32).tolist()
    return [top_preds]
