"""Mirror of cet_pick/utils/post_process.py:11-25 (`tomo_post_process`): host-side grouping of the
(1,K,5) pick rows by integer z, and of :27-106 (`tomo_group_postprocess`, `tomo_fiber_postprocess`): the `--spike` /
`--fiber` graph post-processing of the <= K kept picks that save_detection calls (tomo_det.py:84-95)."""
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


# ------------------------------------------------------------------------------------------------------------------
# `--fiber` / `--spike` graph post-processing of the pick list (post_process.py:27-106).  Host-side work on at most K
# points (K = 900 by default), float64 / int64 numpy like the reference; the connected components come from
# scipy.sparse.csgraph, which is what the reference's sknetwork.topology.get_connected_components returns for a
# square adjacency (labels numbered in order of each component's lowest point index).
def _component_labels(points, distance_cutoff, block=512):
    """labels[i] of the graph that links two picks when their Euclidean distance is <= distance_cutoff (:35-41, :55-61).
    The reference fills a dense n x n matrix row by row; here the links are collected block-wise into a sparse one
    (same arithmetic per pair: sqrt of the integer / float sum of squares), so K = 10 000 picks need megabytes, not gigabytes."""
    from scipy import sparse
    from scipy.sparse.csgraph import connected_components
    p = np.asarray(points)
    n = p.shape[0]
    rows, cols = [], []
    for i0 in range(0, n, block):
        delta = p[i0:i0 + block, None, :] - p[None, :, :]
        r, c = np.nonzero(np.sqrt(np.sum(delta ** 2, axis=2)) <= distance_cutoff)     # the diagonal is always linked
        rows.append(r + i0)
        cols.append(c)
    rows, cols = np.concatenate(rows), np.concatenate(cols)
    adjacency = sparse.csr_matrix((np.ones(rows.size), (rows, cols)), shape=(n, n))
    return connected_components(adjacency, directed=False, return_labels=True)[1]


def k_x(y, a, b, c):
    """post_process.py:27-29: largest value over y of 2a / (1 + (2ay + b)^2)^(2/3), the curvature proxy of a fitted parabola"""
    return np.max((2 * a) / ((1 + (2 * a * y + b) ** 2)) ** (2 / 3))


def tomo_group_postprocess(dets_all, distance_cutoff=15, min_per_group=5):
    """post_process.py:31-50 (`--spike`): keep the picks whose connected group has MORE than min_per_group members.
    dets_all: rows [x, y, z(, score)]; returns the kept rows (numpy rows, all columns), grouped by component label."""
    rows = np.asarray(dets_all)
    if rows.size == 0:                                          # nothing passed the filters (the reference indexes an
        return []                                               # empty array here and raises IndexError)
    labels = _component_labels(rows[:, :3], distance_cutoff)
    sizes = np.bincount(labels)
    order = np.argsort(labels, kind="stable")                  # groups in label order, rows in input order inside one
    order = order[sizes[labels[order]] > min_per_group]
    return [rows[i] for i in order.tolist()]


def _parabola(u, v):
    """np.polyfit(u, v, 2, full=True) -> (coefficients, mean squared residual or 10000 when numpy returns none) (:78-90)"""
    fit = np.polyfit(u, v, 2, full=True)
    return fit[0], (fit[1][0] / u.shape[0] if fit[1].shape[0] > 0 else 10000)


def tomo_fiber_postprocess(dets, distance_cutoff=15, res_cutoff=30, curvature_cutoff=0.03, scale=2):
    """post_process.py:52-106 (`--fiber`): every connected group of more than 6 picks is fitted with two parabolas
    y(x), z(x); a group whose fit is tight and nearly straight is re-sampled every `scale` pixels along x.
    dets: rows [x, y, z]; returns rows [x, z, y] (ints) in the order save_detection prints them."""
    pts = np.asarray(dets)
    out = []
    if pts.size == 0:
        return out
    labels = _component_labels(pts, distance_cutoff)
    for lb in np.unique(labels).tolist():
        grp = pts[labels == lb]
        if grp.shape[0] <= 6:
            continue
        u, v, w = grp[:, 0], grp[:, 1], grp[:, 2]              # abscissa x; fitted: y(x), z(x)
        lo, hi = np.min(u), np.max(u)
        n_curv, n_out = int((hi - lo) // 2), int((hi - lo) // scale)
        if n_curv <= 0:
            continue
        grid = np.linspace(lo - 1, hi + 1, n_curv)             # one pixel beyond both ends (:71)
        grid_out = np.linspace(lo - 1, hi + 1, n_out)
        c_v, res_v = _parabola(u, v)
        c_w, res_w = _parabola(u, w)
        bend_v, bend_w = abs(k_x(grid, *c_v)), abs(k_x(grid, *c_w))
        res = res_v + res_w
        if res < res_cutoff:
            ok = bend_v < curvature_cutoff and bend_w < curvature_cutoff
        elif res < res_cutoff * 3:                              # looser fit: ten times straighter (:99-100)
            ok = bend_v < curvature_cutoff / 10 and bend_w < curvature_cutoff / 10
        else:
            ok = False
        if ok:
            v_out, w_out = np.polyval(c_v, grid_out), np.polyval(c_w, grid_out)
            out.extend([int(a), int(c), int(b)] for a, b, c in zip(grid_out, v_out, w_out))
    return out


def graph_pick_lines(opt, xs, ys, zs, scores, scale=2):
    """The `--fiber` / `--spike` tail of save_detection (tomo_det.py:84-95, tomo_det_classify.py:196-214): text lines for
    the picks (integer x, y, z and float score lists) that already passed the score / border filters."""
    lines = []
    if opt.fiber:
        cand = [[x, y, z] for x, y, z in zip(xs, ys, zs)]
        fitted = tomo_fiber_postprocess(cand, distance_cutoff=opt.distance_cutoff, res_cutoff=opt.r2_cutoff,
                                        curvature_cutoff=opt.curvature_cutoff, scale=scale)
        lines += [str(c[0]) + "\t" + str(c[1]) + "\t" + str(c[2]) for c in fitted]
    else:
        cand = [[x, y, z, s] for x, y, z, s in zip(xs, ys, zs, scores)]
    if opt.spike:
        for c in tomo_group_postprocess(cand, distance_cutoff=opt.distance_cutoff, min_per_group=5):
            ln = str(c[0]) + "\t" + str(c[2]) + "\t" + str(c[1])
            lines.append(ln + "\t" + str(c[3]) if opt.with_score else ln)
    return lines
