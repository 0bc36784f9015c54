"""Exploration-step candidate generator on the device (csrc/explore.cu + preproc.cu through utils/image.py) against
reference-generated fixtures and the oracle: float64 NMS maps, float64 greedy distance suppression, the
difference-of-Gaussians pyramid."""
import numpy as np
import pytest
import torch

import synthdata as synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def im():
    from cet_pick_b200.utils import image
    return image


@pytest.mark.parametrize("name", ["explore_pyramid", "explore_pyramid3"])
def test_pyramid_reference_golden(golden, im, name):
    """scores (float32 bits), coordinates and order identical to the unmodified reference."""
    g = golden(name)
    rec = synth.tomogram_np(*[int(v) for v in g["shape"]], int(g["seed"])).astype(np.float64)
    sc, co = im.get_potential_coords_pyramid(rec, sigmas=[float(s) for s in g["sigmas"]])
    assert sc.dtype == np.float32 and co.dtype == np.int32
    assert np.array_equal(co, g["coords"]) and np.array_equal(bits(sc), bits(g["scores"]))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("fn,win", [("_nms_xy", (1, 3, 3)), ("_nms_z", (5, 1, 1)), ("_nms", (3, 3, 3)), ("_nms_xy", (1, 5, 5))])
def test_nms_maps_vs_torch(im, dtype, fn, win):
    from oracle import explore_oracle as eo
    x = (synth.heatmap_tiefree_np(7, 19, 23, 3).astype(dtype) - dtype(0.5))
    x[2, 4:9, 5:12] = dtype(0.25)                                  # a plateau and signed values
    k = max(win)
    out = getattr(im, fn)(torch.from_numpy(x)[None, None].cuda(), kernel=k)[0, 0].cpu().numpy()
    ref = eo.nms_window(x, *win)
    assert out.dtype == dtype and np.array_equal(out.view(np.uint8), ref.astype(dtype).view(np.uint8))


@pytest.mark.parametrize("shape,d,thr", [((6, 20, 24), 4, 0.6), ((8, 30, 30), 14, 0.5), ((5, 9, 11), 3, float("-inf"))])
def test_greedy_nms_f64_vs_oracle(im, shape, d, thr):
    """float64 scores: order decided in float64, ties in index order (stable sort), float32 scores returned."""
    from oracle import decode_oracle as do
    x = synth.heatmap_tiefree_np(*shape, 50 + d).astype(np.float64)
    x = x + 1e-12 * np.arange(x.size, dtype=np.float64).reshape(shape)[::-1, ::-1, ::-1]   # distinct only in float64
    x[1, 3, 4] = x[2, 5, 6] = x[4, 7, 3] = 0.93                                      # exact ties
    sc, co = im.non_maximum_suppression_3d(x, d, threshold=thr)
    rs, rc = do.greedy_distance_nms(x, d, threshold=thr)
    assert np.array_equal(co, rc) and np.array_equal(bits(sc), bits(rs))


def test_extract_subvols_vs_reference_formula(im):
    """candidate patches: z-slab sum, min-max in float64, float32 out (new3d_vol.py:117-128 restated with numpy)."""
    rng = np.random.default_rng(4)
    v = rng.standard_normal((20, 60, 70))
    coords = np.array([[35, 30, 10], [20, 18, 1], [52, 41, 19], [16, 16, 5]], dtype=np.int32)     # z = 19: slab clipped at the top
    sub = [3, 32, 32]
    out = im.extract_subvols(v, coords, sub).cpu().numpy()
    assert out.shape == (4, 1, 32, 32) and out.dtype == np.float32
    for n, (x, y, z) in enumerate(coords):
        e = v[z - 1:z + 2, y - 16:y + 16, x - 16:x + 16].copy()
        e = np.sum(e, axis=0)
        e = (e - np.min(e)) / (np.max(e) - np.min(e))
        assert np.array_equal(out[n, 0], e.astype(np.float32))
    with pytest.raises(ValueError):
        im.extract_subvols(v, np.array([[5, 30, 10]]), sub)
