"""`--fiber` / `--spike` graph post-processing (cet_pick/utils/post_process.py:27-106 and the save_detection tails,
tomo_det.py:84-95, tomo_det_classify.py:196-214): the oracle restatement and the product against fixtures written by
the reference's own functions (tests/golden/make_golden_postproc.py).  Host-side numpy work: integer rows and text
lines must be identical."""
import types

import numpy as np
import pytest
import torch

from oracle import postproc_oracle as po
from cet_pick_b200.utils import post_process as pp

SEEDS = [3, 4, 5]
FIBER_TAGS = ["d15", "d10s3", "d25"]


def _fiber_args(g, tag):
    d, r, c, s = [float(v) for v in g["fiber_" + tag + "_args"]]
    return dict(distance_cutoff=d, res_cutoff=r, curvature_cutoff=c, scale=s)


@pytest.mark.parametrize("seed", SEEDS)
def test_fixture_inputs_are_reproducible(golden, seed):
    g = golden("postproc_s%d" % seed)
    kw = {3: {}, 4: dict(n_fibers=6, n_clusters=5, n_stray=60), 5: dict(n_fibers=2, n_clusters=0, n_stray=5)}[seed]
    pts, scores = po.synthetic_picks(seed, **kw)
    assert np.array_equal(pts, g["pts"]) and np.array_equal(scores, g["scores"])


@pytest.mark.parametrize("impl", ["oracle", "product"])
@pytest.mark.parametrize("seed", SEEDS)
def test_fiber_postprocess_equals_reference(golden, seed, impl):
    g = golden("postproc_s%d" % seed)
    fn = po.fiber_postprocess if impl == "oracle" else pp.tomo_fiber_postprocess
    for tag in FIBER_TAGS:
        out = np.asarray(fn(g["pts"].tolist(), **_fiber_args(g, tag)), dtype=np.int64).reshape(-1, 3)
        assert np.array_equal(out, g["fiber_" + tag]), (seed, tag)


@pytest.mark.parametrize("impl", ["oracle", "product"])
@pytest.mark.parametrize("seed", SEEDS)
def test_group_postprocess_equals_reference(golden, seed, impl):
    g = golden("postproc_s%d" % seed)
    fn = po.group_postprocess if impl == "oracle" else pp.tomo_group_postprocess
    rows4 = [[int(p[0]), int(p[1]), int(p[2]), float(s)] for p, s in zip(g["pts"], g["scores"])]
    for d in (15, 8):
        out4 = np.asarray(fn(rows4, distance_cutoff=d, min_per_group=5)).reshape(-1, 4)
        out3 = np.asarray(fn(g["pts"].tolist(), distance_cutoff=d, min_per_group=5)).reshape(-1, 3)
        assert out4.dtype == np.float64 and np.array_equal(out4, g["group4_d%d" % d])
        assert out3.dtype == np.int64 and np.array_equal(out3, g["group3_d%d" % d])


def test_component_labels_follow_lowest_index_order():
    """the labelling convention the oracle states for scipy / sknetwork: components numbered by their lowest member"""
    pts = np.array([[100, 0, 0], [0, 0, 0], [103, 0, 0], [50, 50, 0], [2, 1, 0], [106, 0, 0]])
    assert po.connected_labels(pts, 5).tolist() == [0, 1, 0, 2, 1, 0]
    assert pp._component_labels(pts, 5).tolist() == [0, 1, 0, 2, 1, 0]
    # chained reachability (0-2-5 are 3 apart pairwise-adjacent only through the middle one)
    assert po.connected_labels(pts, 3).tolist() == pp._component_labels(pts, 3).tolist() == [0, 1, 0, 2, 1, 0]


def test_empty_pick_list_gives_no_lines():
    assert pp.tomo_fiber_postprocess([]) == [] and pp.tomo_group_postprocess([]) == []


def _opt(**kw):
    o = types.SimpleNamespace(down_ratio=2, out_thresh=0.25, cutoff_z=3, compress=False, fiber=False, spike=False,
                              with_score=False, distance_cutoff=15.0, r2_cutoff=30.0, curvature_cutoff=0.03,
                              distance_scale=2.0, nms=3)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _dets_dict(g):
    dets = {}
    for p, s in zip(g["pts"], g["scores"]):
        dets.setdefault(int(p[2]), []).append([float(p[0]) + 0.5, float(p[1]) + 0.5, float(p[2]), float(s), float(s)])
    return {k: dets[k] for k in sorted(dets)}


@pytest.mark.parametrize("tag,kw", [("fiber", dict(fiber=True)),
                                    ("fiber_compress", dict(fiber=True, compress=True, distance_cutoff=20.0, distance_scale=3.0))])
def test_tomodet_save_detection_fiber_file(golden, tmp_path, tag, kw):
    """TomodetDetector.save_detection with --fiber writes the reference's file, byte for byte (host tensor path)."""
    from cet_pick_b200.detectors.tomo_det import TomodetDetector
    g = golden("postproc_file")
    det = TomodetDetector.__new__(TomodetDetector)
    det.opt = _opt(**kw)
    hm = torch.zeros((1, 1) + tuple(int(v) for v in g["hm_shape"]))
    det.save_detection(hm, _dets_dict(g), str(tmp_path), None, name="tomoF")
    assert open(tmp_path / "tomoF.txt").read() == str(g["txt_" + tag])


def test_tomodet_spike_follows_the_classify_writer(golden, tmp_path):
    """--spike on TomodetDetector cannot run in the reference (tomo_det.py:90 calls a function the module never
    imports); here it behaves like TomoClassdetDetector.save_detection's --spike branch."""
    from cet_pick_b200.detectors.tomo_det import TomodetDetector
    g = golden("postproc_file")
    det = TomodetDetector.__new__(TomodetDetector)
    det.opt = _opt(spike=True, with_score=True)
    D, Hh, Wh = (int(v) for v in g["hm_shape"])
    hm = torch.zeros((1, 1, D, Hh, Wh))
    dets = _dets_dict(g)
    det.save_detection(hm, dets, str(tmp_path), None, name="tomoF")
    rows = [[int(c[0]), int(c[1]), int(c[2]), float(c[3])] for v in dets.values() for c in v
            if c[3] > 0.25 and 3 <= int(c[2]) <= D - 3 and 20 < int(c[0]) < 2 * Wh - 20 and 20 < int(c[1]) < 2 * Hh - 20]
    want = "".join("%s\t%s\t%s\t%s\n" % (str(c[0]), str(c[2]), str(c[1]), str(c[3]))
                   for c in po.group_postprocess(rows, distance_cutoff=15.0, min_per_group=5))
    got = open(tmp_path / "tomoF.txt").read()
    assert got == want and len(got.splitlines()) > 50


@pytest.mark.parametrize("tag,kw", [("cls_fiber", dict(fiber=True)), ("cls_spike", dict(spike=True)),
                                    ("cls_spike_score", dict(spike=True, with_score=True, distance_cutoff=9.0)),
                                    ("cls_fiber_spike", dict(fiber=True, spike=True)),
                                    ("cls_plain", dict()), ("cls_score", dict(with_score=True)),
                                    ("cls_compress", dict(compress=True, out_thresh=0.6))])
def test_classdet_save_detection_graph_files(golden, tmp_path, tag, kw):
    from cet_pick_b200.detectors.tomo_det_classify import TomoClassdetDetector
    g = golden("postproc_file")
    det = TomoClassdetDetector.__new__(TomoClassdetDetector)
    det.opt = _opt(**kw)
    rows4 = np.concatenate([g["pts"].astype(np.float32) + 0.5, g["scores"][:, None]], 1).astype(np.float32)
    hm = torch.zeros((1, 1) + tuple(int(v) for v in g["hm_cls_shape"]))
    det.save_detection(hm, rows4, str(tmp_path), None, name="tomoF")
    assert open(tmp_path / "tomoF.txt").read() == str(g["txt_" + tag])
