"""GPU pre-processing (csrc/preproc.cu through utils/loader.py) against the reference-generated fixtures and
the numpy/scipy oracle.  float64 on both sides; the only admissible difference is a reduction-order ulp in
mean/std, which can move a voxel that sits exactly on a level boundary (never seen on these inputs: the
256-level outputs are asserted identical)."""
import numpy as np
import pytest
import torch

import synthdata as synth

pytestmark = pytest.mark.gpu

PRE_CASES = ["preproc_xzy_c_g08", "preproc_xzy_g0", "preproc_zxy_c_g15", "preproc_xyz_c_odd", "preproc_yxz_g2"]


def pre_volume(shape, seed):
    v = synth.tomogram_np(*[int(s) for s in shape], int(seed))
    return (v * 37.5 - 11.0).astype(np.float32)


@pytest.fixture(scope="module")
def ld():
    from cet_pick_b200.utils import loader
    return loader


@pytest.mark.parametrize("name", PRE_CASES)
def test_reference_golden(golden, ld, name, tmp_path):
    from cet_pick_b200.utils import mrcio
    g = golden(name)
    path = str(tmp_path / "v.mrc")
    mrcio.write_mrc(path, pre_volume(g["shape"], g["seed"]))
    rec = ld.load_rec(path, order=str(g["order"]), compress=bool(g["compress"]))
    assert rec.is_cuda and rec.dtype == torch.float64 and tuple(rec.shape) == g["rec"].shape
    assert np.abs(rec.cpu().numpy() - g["rec"]).max() <= 1e-12
    im = ld.preprocess(rec, denoise=float(g["sigma"]))
    assert im.dtype == torch.float64
    assert np.array_equal(im.cpu().numpy(), g["im"])
    im32 = ld.load_tomos_from_list(["a"], [path], order=str(g["order"]), compress=bool(g["compress"]),
                                   denoise=float(g["sigma"]), dtype=torch.float32)["a"]
    assert np.array_equal(im32.cpu().numpy(), g["im"].astype(np.float32))


@pytest.mark.parametrize("shape,order,compress,sigma,dtype", [
    ((40, 64, 72), "xzy", True, 0.8, np.float32),
    ((33, 50, 47), "xzy", False, 1.3, np.int16),
    ((30, 41, 52), "zxy", True, 0.0, np.uint16),
    ((5, 3, 4), "xyz", True, 2.5, np.float32),       # lines shorter than the kernel radius: repeated reflection
    ((21, 34, 30), "yxz", False, 0.8, np.int8),
])
def test_vs_oracle(ld, shape, order, compress, sigma, dtype):
    from oracle import preproc_oracle as po
    v = pre_volume(shape, 7)
    if dtype != np.float32:
        info = np.iinfo(dtype)
        v = np.clip(np.round(v * 3), info.min, info.max).astype(dtype)
    rec = ld.load_rec(v, order=order, compress=compress)
    ref = po.load_rec(v, order=order, compress=compress)
    assert np.abs(rec.cpu().numpy() - ref).max() <= 1e-12
    im = ld.preprocess(rec, denoise=sigma).cpu().numpy()
    ref_im = po.preprocess(ref, denoise=sigma)
    lv = np.abs(im - ref_im) * 255
    assert (lv > 1e-9).mean() <= 1e-4 and lv.max() <= 1.0 + 1e-9      # at most a stray boundary voxel, one level
    # gaussian alone: same taps, same operation order as scipy's correlate1d
    if sigma > 0:
        from scipy.ndimage import gaussian_filter
        gf = ld.gaussian_filter(torch.from_numpy(ref).cuda(), sigma).cpu().numpy()
        assert np.abs(gf - gaussian_filter(ref, sigma)).max() <= 1e-14


def test_quantize_half_to_even_and_errors(ld):
    x = torch.tensor([[[-3.0, -2.5, 2.0, 3.0, -0.25, 0.0]]], dtype=torch.float64)
    q = ld.quantize(x.cuda(), mi=-2.5, ma=2).cpu().numpy().ravel()
    ref = np.round(np.clip(255 * (x.numpy().ravel() + 2.5) / 4.5, 0, 255)).astype(np.uint8)
    assert np.array_equal(q, ref)
    with pytest.raises(IndexError):
        ld.load_rec(np.zeros((3, 4, 4), np.float32), order="zxy", compress=True)


@pytest.mark.parametrize("name", ["preproc_tilt_zxy_g0", "preproc_tilt_zxy_g1", "preproc_tilt_xzy_c"])
def test_tilt_branch_reference_golden(golden, ld, name):
    """is_tilt=True: per-slice z-score / 2-D Gaussian / quantize / min-max.  The reference takes the slice
    statistics in float32 (the file's dtype), the device in float64: values agree to float32 rounding, and a
    voxel that sits on a level boundary may land one level away."""
    g = golden(name)
    v = pre_volume(g["shape"], g["seed"])
    rec = ld.load_rec(v, order=str(g["order"]), compress=bool(g["compress"]), is_tilt=True)
    assert rec.dtype == torch.float64 and tuple(rec.shape) == g["rec"].shape
    assert np.abs(rec.cpu().numpy() - g["rec"]).max() <= 5e-6
    im = ld.preprocess(torch.from_numpy(g["rec"]), denoise=float(g["sigma"]), is_tilt=True)   # same input as the reference
    assert im.dtype == torch.float32
    d = np.abs(im.cpu().numpy() - g["im"])
    assert (d > 1e-6).mean() <= 2e-3 and d.max() <= 1.0 / 100


@pytest.mark.parametrize("sigma", [0, 0.8])
def test_preprocess_levels_equals_preprocess(ld, sigma):
    """preprocess_levels = preprocess stopped before the min-max division: level_values[levels] is the float32 volume."""
    rng = np.random.default_rng(3)
    vol = rng.normal(size=(6, 40, 48))
    full = ld.preprocess(torch.from_numpy(vol), denoise=sigma, dtype=torch.float64).cpu().numpy().astype(np.float32)
    q, lv = ld.preprocess_levels(torch.from_numpy(vol), denoise=sigma)
    assert q.dtype == torch.uint8 and lv.dtype == np.float32 and lv[0] == 0
    assert np.array_equal(lv[q.cpu().numpy()], full)
