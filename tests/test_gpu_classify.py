"""Semiclass path on the device: greedy distance suppression (csrc/greedy_nms.cu, decode.py:42-79) and the
tile scheduler mirror (detectors/tomo_det_classify.py) against reference-generated fixtures and the oracle."""
import hashlib
import types

import numpy as np
import pytest
import torch

import synthdata as synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def dec():
    from cet_pick_b200.models import decode
    return decode


@pytest.mark.parametrize("name", ["greedy_nms_d3", "greedy_nms_d5", "greedy_nms_d8_large"])
def test_greedy_nms_reference_golden(golden, dec, name):
    """bit-exact scores, identical coordinates, identical order (tie-free maps: the order is defined)."""
    g = golden(name)
    D, H, W = [int(v) for v in g["shape"]]
    x = synth.heatmap_tiefree_np(D, H, W, int(g["seed"]))
    sc, co = dec.non_maximum_suppression_3d(x, float(g["d"]), threshold=float(g["threshold"]))
    assert np.array_equal(bits(sc), bits(g["scores"])) and np.array_equal(co, g["coords"])
    assert co.dtype == np.int32 and sc.dtype == np.float32


@pytest.mark.parametrize("shape,d,scale,thr", [
    ((5, 9, 11), 3, 1.0, float("-inf")),        # every voxel is visited (the reference's default threshold)
    ((6, 20, 24), 4, 1.0, 0.6),
    ((4, 16, 16), 5, 0.5, 0.7),                  # scale shrinks the ball (decode.py:44)
    ((3, 7, 40), 1, 1.0, 0.5),                   # r = 0.5: only the voxel itself
    ((8, 30, 30), 7, 1.0, 0.8),
])
def test_greedy_nms_vs_oracle(dec, shape, d, scale, thr):
    from oracle import decode_oracle as do
    x = synth.heatmap_tiefree_np(*shape, 100 + d)
    x = ((x - x.min()) / (x.max() - x.min())).astype(np.float32)
    sc, co = dec.non_maximum_suppression_3d(x, d, scale=scale, threshold=thr)
    rs, rc = do.greedy_distance_nms(x, d, scale=scale, threshold=thr)
    assert np.array_equal(bits(sc), bits(rs)) and np.array_equal(co, rc)


def test_greedy_nms_plateaus_and_long_chains(dec):
    """ties (canonical index order, as in the oracle) and a monotone ramp whose dependency chain spans the
    whole row: the round-based resolution must still equal the sequential loop."""
    from oracle import decode_oracle as do
    x = np.zeros((3, 8, 200), dtype=np.float32)
    x[1, 4, :] = np.linspace(1.0, 2.0, 200, dtype=np.float32)       # ramp: each voxel depends on its right neighbour
    x[0, 2, 10:60] = 0.75                                            # plateau
    x[2, 6, 100:130] = 0.75
    sc, co = dec.non_maximum_suppression_3d(x, 3, threshold=0.5)
    rs, rc = do.greedy_distance_nms(x, 3, threshold=0.5)
    assert np.array_equal(bits(sc), bits(rs)) and np.array_equal(co, rc)
    assert dec.non_maximum_suppression_3d.last_rounds >= 100         # the ramp needs ~one round per pick


def test_greedy_nms_empty_and_errors(dec):
    x = synth.heatmap_tiefree_np(4, 8, 8, 1)
    sc, co = dec.non_maximum_suppression_3d(x, 3, threshold=2.0)      # nothing above the threshold
    assert sc.shape == (0,) and co.shape == (0, 3)
    with pytest.raises(ValueError):
        dec.non_maximum_suppression_3d(x[0], 3)
    with pytest.raises(RuntimeError):                                 # more candidates than the workspace was sized for
        dec.non_maximum_suppression_3d(x, 3, threshold=0.0, max_candidates=10)


def test_tomo_decode_classify_shape_and_dtype(dec):
    from oracle import decode_oracle as do
    x = synth.heatmap_tiefree_np(6, 12, 14, 8)
    thr = float(np.median(x))
    dets = dec.tomo_decode_classify(torch.from_numpy(x)[None].cuda(), 3, thr)
    rs, rc = do.greedy_distance_nms(x, 3, threshold=thr)
    assert dets.dtype == torch.float32 and dets.device.type == "cpu" and dets.shape == (len(rs), 4)
    assert np.array_equal(dets[:, :3].numpy(), rc.astype(np.float32)) and np.array_equal(bits(dets[:, 3].numpy()), bits(rs))


def test_patch_dataset_reference_golden(golden):
    from cet_pick_b200.detectors.tomo_det_classify import PatchDataset
    g = golden("patch_dataset")
    D, H, W = [int(v) for v in g["shape"]]
    vol = torch.from_numpy(synth.heatmap_tiefree_np(D, H, W, int(g["seed"]))).cuda()
    ds = PatchDataset(vol, *[int(v) for v in g["params"]])
    assert len(ds) == len(g["digest"]) and tuple(ds.shape) == tuple(g["grid"])
    for n in range(len(ds)):
        idx, x = ds[n]
        assert x.is_cuda and np.array_equal(idx, g["index"][n])
        dig = np.frombuffer(hashlib.sha1(x.cpu().numpy().tobytes()).digest()[:8], dtype=np.uint64)[0]
        assert dig == g["digest"][n]


@pytest.mark.parametrize("name", ["classify_tiled", "classify_whole"])
def test_classify_detector_reference_golden(golden, name, tmp_path):
    from cet_pick_b200.detectors.tomo_det_classify import TomoClassdetDetector
    g = golden(name)
    D, H, W = [int(v) for v in g["shape"]]
    opt = types.SimpleNamespace(gpus=[0], nms=int(g["nms"]), out_thresh=float(g["out_thresh"]), down_ratio=1,
                                cutoff_z=2, compress=False, with_score=True)
    det = TomoClassdetDetector(opt, model=synth.fullres_stub_model())
    x = torch.from_numpy(synth.heatmap_tiefree_np(D, H, W, int(g["seed"])))[None].cuda()
    output, dets, hm = det.process(x)
    assert output is None and tuple(hm.shape) == tuple(int(v) for v in g["hm_shape"])
    assert abs(hm.double().sum().item() - float(g["hm_sum"])) <= 1e-2
    key = lambda a: a[np.lexsort((a[:, 0], a[:, 1], a[:, 2]))]
    a, b = key(dets.numpy()), key(g["dets"])
    assert a.shape == b.shape and np.array_equal(a[:, :3], b[:, :3])      # same picks (expf vs Sleef: ulps in the score)
    assert np.abs(a[:, 3] - b[:, 3]).max() <= 1e-6
    dd, nm = det.post_process(dets, {"name": ["vol"]})
    det.save_detection(hm, dd, str(tmp_path), None, name=nm)
    lines = open(tmp_path / "vol.txt").read().splitlines()
    assert 0 < len(lines) <= len(a) and len(lines[0].split("\t")) == 4
