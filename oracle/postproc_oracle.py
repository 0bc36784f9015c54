"""TEST INFRASTRUCTURE (never imported by the product): CPU restatement of the `--fiber` / `--spike` graph
post-processing of the pick list, cet_pick/utils/post_process.py:27-106, and of the tail of save_detection that calls
it (cet_pick/detectors/tomo_det.py:84-95, cet_pick/detectors/tomo_det_classify.py:196-214).

Third-party arithmetic: the connected components come from scikit_network==0.28.2 (requirements.txt:17), absent from
this image.  Its `sknetwork.topology.get_connected_components` hands a square adjacency to
`scipy.sparse.csgraph.connected_components(adjacency, connection='weak', return_labels=True)[1]`; the published
behaviour restated here: labels are 0, 1, 2, ... in the order in which a scan over the points 0 .. n-1 meets a point
that has no label yet, every point reachable from it getting the same label.  Pinned by tests/golden/postproc_*.npz,
which tests/golden/make_golden_postproc.py writes by running the reference's own functions with that scipy call in
place of the missing module.  The parabola fits are numpy.polyfit / numpy.polyval, as in the reference.
"""
from __future__ import annotations

import numpy as np


def connected_labels(points, distance_cutoff):
    """post_process.py:35-41 / :55-61: link i-j when sqrt(sum((p_i - p_j)^2)) <= cutoff, then label the components"""
    p = np.asarray(points)
    n = p.shape[0]
    nbrs = []
    for i in range(n):
        dist = np.sqrt(np.sum((p[i] - p) ** 2, 1))
        nbrs.append(np.where(dist <= distance_cutoff)[0])
    labels = np.full(n, -1, dtype=np.int64)
    nxt = 0
    for s in range(n):
        if labels[s] >= 0:
            continue
        labels[s] = nxt
        stack = [s]
        while stack:
            i = stack.pop()
            for j in nbrs[i]:
                if labels[j] < 0:
                    labels[j] = nxt
                    stack.append(int(j))
        nxt += 1
    return labels


def curvature(y, a, b, c):
    """post_process.py:27-29"""
    k = (2 * a) / ((1 + (2 * a * y + b) ** 2)) ** (2 / 3)
    return np.max(k)


def group_postprocess(dets_all, distance_cutoff=15, min_per_group=5):
    """post_process.py:31-50"""
    kept = []
    rows = np.asarray(dets_all)
    labels = connected_labels(rows[:, :3], distance_cutoff)
    for lb in np.unique(labels):
        members = rows[np.where(labels == lb)[0]]
        if members.shape[0] > min_per_group:
            for j in range(members.shape[0]):
                kept.append(members[j])
    return kept


def fiber_postprocess(dets, distance_cutoff=15, res_cutoff=30, curvature_cutoff=0.03, scale=2):
    """post_process.py:52-106"""
    out = []
    pts = np.asarray(dets)
    labels = connected_labels(pts, distance_cutoff)
    groups = []
    for lb in np.unique(labels):
        members = pts[np.where(labels == lb)[0]]
        if members.shape[0] > 6:
            groups.append(members)
    for g in groups:
        line = g.copy()
        line[:, [1, 0]] = line[:, [0, 1]]                       # columns (y, x, z): the abscissa is column 1
        span = np.max(line[:, 1]) - np.min(line[:, 1])
        n_fit, n_out = span // 2, span // scale
        t = np.linspace(np.min(line[:, 1]) - 1, np.max(line[:, 1]) + 1, int(n_fit))
        t_out = np.linspace(np.min(line[:, 1]) - 1, np.max(line[:, 1]) + 1, int(n_out))
        if t.shape[0] > 0:
            fit_a = np.polyfit(line[:, 1], line[:, 0], 2, full=True)
            fit_b = np.polyfit(line[:, 1], line[:, 2], 2, full=True)
            npts = line.shape[0]
            res_a = fit_a[1][0] / npts if fit_a[1].shape[0] > 0 else 10000
            res_b = fit_b[1][0] / npts if fit_b[1].shape[0] > 0 else 10000
            ka, kb = curvature(t, *fit_a[0]), curvature(t, *fit_b[0])
            emit = False
            if res_a + res_b < res_cutoff:
                emit = abs(ka) < curvature_cutoff and abs(kb) < curvature_cutoff
            elif res_a + res_b < res_cutoff * 3:
                emit = abs(ka) < curvature_cutoff / 10 and abs(kb) < curvature_cutoff / 10
            if emit:
                a_out, b_out = np.polyval(fit_a[0], t_out), np.polyval(fit_b[0], t_out)
                for j in range(a_out.shape[0]):
                    out.append([int(t_out[j]), int(b_out[j]), int(a_out[j])])
    return out


def synthetic_picks(seed, n_fibers=4, n_clusters=3, n_stray=25, extent=(400, 400, 120)):
    """A seeded pick list for the fixtures: gently curved filaments sampled every ~6 px with jitter, blobs of nearby
    picks, strays.  Returns integer rows [x, y, z] and a float32 score per row, shuffled."""
    rng = np.random.default_rng(seed)
    X, Y, Z = extent
    rows = []
    for f in range(n_fibers):
        x0, x1 = sorted(rng.uniform(30, X - 30, 2))
        if x1 - x0 < 60:
            x1 = min(X - 25, x0 + 60 + rng.uniform(0, 80))
        xs = np.arange(x0, x1, rng.uniform(4.0, 7.0))
        bend = rng.uniform(-1, 1) * (2e-3 if f % 2 == 0 else 2e-4)
        ys = rng.uniform(60, Y - 60) + rng.uniform(-0.3, 0.3) * (xs - x0) + bend * (xs - x0) ** 2
        zs = rng.uniform(30, Z - 30) + rng.uniform(-0.05, 0.05) * (xs - x0)
        jit = rng.normal(0, 0.8 if f < n_fibers - 1 else 3.5, (xs.size, 3))      # the last filament is a loose fit
        rows += [[x + j[0], y + j[1], z + j[2]] for x, y, z, j in zip(xs, ys, zs, jit)]
    for c in range(n_clusters):
        ctr = [rng.uniform(40, X - 40), rng.uniform(40, Y - 40), rng.uniform(20, Z - 20)]
        m = int(rng.integers(4, 12))
        rows += (np.asarray(ctr) + rng.normal(0, 5.0, (m, 3))).tolist()
    rows += np.stack([rng.uniform(25, X - 25, n_stray), rng.uniform(25, Y - 25, n_stray), rng.uniform(5, Z - 5, n_stray)], 1).tolist()
    pts = np.floor(np.asarray(rows)).astype(np.int64)
    pts = pts[rng.permutation(pts.shape[0])]
    scores = rng.uniform(0.3, 0.99, pts.shape[0]).astype(np.float32)
    return pts, scores
