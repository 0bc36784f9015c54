"""Fixtures for the `--fiber` / `--spike` graph post-processing from the REAL reference
(cet_pick/utils/post_process.py:31-106 and the save_detection tails that call it).

scikit_network==0.28.2 (requirements.txt:17) is not in this image; its get_connected_components is a thin wrapper of
scipy.sparse.csgraph.connected_components(adjacency, connection='weak', return_labels=True)[1], which is installed in
its place before the reference module is imported.  Everything else is the reference's own code.

    python tests/golden/make_golden_postproc.py      # needs /root/reference (or $CET_PICK_REF)
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import postproc_oracle as po       # noqa: E402  (seeded pick lists only)
from oracle import refbridge                   # noqa: E402

refbridge.install()
topo = sys.modules["sknetwork.topology"]
topo.get_connected_components = lambda adj, connection="weak": \
    sparse.csgraph.connected_components(adj, connection=connection, return_labels=True)[1]
import cet_pick.utils.post_process as rpp                        # noqa: E402
from cet_pick.detectors.tomo_det import TomodetDetector        # noqa: E402
import cet_pick.detectors.tomo_det_classify as rcls              # noqa: E402
from cet_pick.detectors.tomo_det_classify import TomoClassdetDetector   # noqa: E402
if not hasattr(np, "float"):
    np.float = float                            # tomo_det_classify.py:191 uses the alias numpy 1.24 removed

for seed, kw in ((3, dict()), (4, dict(n_fibers=6, n_clusters=5, n_stray=60)), (5, dict(n_fibers=2, n_clusters=0, n_stray=5))):
    pts, scores = po.synthetic_picks(seed, **kw)
    out = {}
    for tag, fk in (("d15", dict(distance_cutoff=15, res_cutoff=30, curvature_cutoff=0.003, scale=2)),
                    ("d10s3", dict(distance_cutoff=10, res_cutoff=30, curvature_cutoff=0.03, scale=3.0)),
                    ("d25", dict(distance_cutoff=25, res_cutoff=8, curvature_cutoff=0.03, scale=2))):
        fib = rpp.tomo_fiber_postprocess(pts.tolist(), **fk)
        out["fiber_" + tag] = np.asarray(fib, dtype=np.int64).reshape(-1, 3)
        out["fiber_" + tag + "_args"] = np.asarray([fk["distance_cutoff"], fk["res_cutoff"], fk["curvature_cutoff"], fk["scale"]])
    for d in (15, 8):
        rows4 = [[int(p[0]), int(p[1]), int(p[2]), float(s)] for p, s in zip(pts, scores)]
        out["group4_d%d" % d] = np.asarray(rpp.tomo_group_postprocess(rows4, distance_cutoff=d, min_per_group=5)).reshape(-1, 4)
        out["group3_d%d" % d] = np.asarray(rpp.tomo_group_postprocess(pts.tolist(), distance_cutoff=d, min_per_group=5)).reshape(-1, 3)
    np.savez_compressed(os.path.join(HERE, "postproc_s%d.npz" % seed), pts=pts, scores=scores, **out)
    print("seed", seed, "points", pts.shape[0], {k: v.shape[0] for k, v in out.items() if not k.endswith("args")})

# file level: TomodetDetector.save_detection with --fiber (and --fiber --spike) on a dets dict built from a pick list
pts, scores = po.synthetic_picks(9, n_fibers=5, n_clusters=3, n_stray=30, extent=(300, 280, 60))
D, Hh, Wh = 64, 150, 160                      # heat-map (D, H', W'); picks live on the 2x grid (300 x 320)
dets = {}
for p, s in zip(pts, scores):
    dets.setdefault(int(p[2]), []).append([float(p[0]) + 0.5, float(p[1]) + 0.5, float(p[2]), float(s), float(s)])
dets = {k: dets[k] for k in sorted(dets)}
hm = np.zeros((1, 1, D, Hh, Wh), dtype=np.float32)
hm_cls = np.zeros((1, 1, D, 2 * Hh, 2 * Wh), dtype=np.float32)
files = {}
CASES = {"fiber": (TomodetDetector, dict(fiber=True)),
         "fiber_compress": (TomodetDetector, dict(fiber=True, compress=True, distance_cutoff=20.0, distance_scale=3.0)),
         "cls_fiber": (TomoClassdetDetector, dict(fiber=True)),
         "cls_spike": (TomoClassdetDetector, dict(spike=True)),
         "cls_spike_score": (TomoClassdetDetector, dict(spike=True, with_score=True, distance_cutoff=9.0)),
         "cls_fiber_spike": (TomoClassdetDetector, dict(fiber=True, spike=True)),
         "cls_plain": (TomoClassdetDetector, dict()),
         "cls_score": (TomoClassdetDetector, dict(with_score=True)),
         "cls_compress": (TomoClassdetDetector, dict(compress=True, out_thresh=0.6))}
# TomodetDetector with --spike cannot be pinned: tomo_det.py never imports tomo_group_postprocess (NameError at :90).
rows4 = np.concatenate([pts.astype(np.float32) + 0.5, scores[:, None]], 1).astype(np.float32)
for tag, (cls, kw) in CASES.items():
    det = cls.__new__(cls)
    det.opt = types.SimpleNamespace(down_ratio=2, out_thresh=0.25, cutoff_z=3, compress=False, fiber=False, spike=False,
                                    with_score=False, distance_cutoff=15.0, r2_cutoff=30.0, curvature_cutoff=0.03,
                                    distance_scale=2.0)
    for k, v in kw.items():
        setattr(det.opt, k, v)
    with tempfile.TemporaryDirectory() as td:
        if cls is TomodetDetector:
            det.save_detection(torch.from_numpy(hm), dets, td, None, name="tomoF")
        else:                                   # the classify detector filters against the heat-map size itself
            det.save_detection(torch.from_numpy(hm_cls), rows4, td, None, name="tomoF")
        import gc
        del det
        gc.collect()                            # the reference never closes its pick file
        files[tag] = open(os.path.join(td, "tomoF.txt")).read()
    print(tag, len(files[tag].splitlines()), "lines")
np.savez_compressed(os.path.join(HERE, "postproc_file.npz"), pts=pts, scores=scores, hm_shape=(D, Hh, Wh),
                    hm_cls_shape=hm_cls.shape[2:], **{"txt_" + k: v for k, v in files.items()})
