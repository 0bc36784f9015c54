"""Host-side pieces that need no GPU: MRC I/O, Gaussian taps, post-processing buckets, error paths."""
import os
import struct

import numpy as np
import pytest


def test_mrc_roundtrip_and_modes(tmp_path):
    from cet_pick_b200.utils import mrcio
    rng = np.random.default_rng(0)
    v = rng.standard_normal((5, 6, 7)).astype(np.float32)
    p = str(tmp_path / "a.mrc")
    mrcio.write_mrc(p, v)
    assert np.array_equal(mrcio.read_mrc(p), v)
    hdr = open(p, "rb").read(1024)
    assert struct.unpack_from("<4i", hdr, 0) == (7, 6, 5, 2) and hdr[208:212] == b"MAP "
    # integer modes as written by other packages: patch the mode word and the payload
    for mode, dt in ((0, np.int8), (1, np.int16), (6, np.uint16)):
        data = (rng.integers(0, 100, size=(3, 4, 5))).astype(dt)
        h = bytearray(1024)
        struct.pack_into("<4i", h, 0, 5, 4, 3, mode)
        q = str(tmp_path / f"m{mode}.mrc")
        open(q, "wb").write(bytes(h) + data.tobytes())
        out = mrcio.read_mrc(q)
        assert out.dtype == dt and np.array_equal(out, data)
    h = bytearray(1024)
    struct.pack_into("<4i", h, 0, 1, 1, 1, 4)
    open(str(tmp_path / "bad.mrc"), "wb").write(bytes(h) + b"\0" * 8)
    with pytest.raises(ValueError):
        mrcio.read_mrc(str(tmp_path / "bad.mrc"))


@pytest.mark.parametrize("sigma", [0.8, 1.0, 2.5, 5.0])
def test_gaussian_taps_equal_scipy(sigma):
    from scipy.ndimage._filters import _gaussian_kernel1d
    from cet_pick_b200.utils.loader import gaussian_kernel1d
    w, r = gaussian_kernel1d(sigma)
    assert r == int(4.0 * sigma + 0.5) and np.array_equal(w, _gaussian_kernel1d(sigma, 0, r))


def test_tomo_post_process_buckets_by_z():
    """utils/post_process.py:11-25: rows grouped by exact z, order inside a bucket preserved, last batch item wins."""
    from cet_pick_b200.utils.post_process import tomo_post_process
    from oracle import decode_oracle as do
    dets = np.array([[[1.5, 2.5, 3.0, 0.9, 0.9], [4.5, 5.5, 1.0, 0.8, 0.8], [6.5, 7.5, 3.0, 0.7, 0.7]],
                     [[0.5, 0.5, 0.0, 0.6, 0.6], [2.5, 2.5, 2.0, 0.5, 0.5], [3.5, 3.5, 2.0, 0.4, 0.4]]], dtype=np.float32)
    ours = tomo_post_process(dets.copy(), z_dim_tot=4)
    ref = do.tomo_post_process(dets.copy(), z_dim_tot=4)
    assert ours == ref and sorted(ours[0].keys()) == [0, 2]
    assert [r[0] for r in ours[0][2]] == [2.5, 3.5]


def test_loader_and_image_have_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from cet_pick_b200.utils import image, loader
    with pytest.raises(RuntimeError):
        loader.preprocess(np.zeros((2, 4, 4), np.float32))
    with pytest.raises(RuntimeError):
        image.get_potential_coords_pyramid(np.zeros((40, 80, 80)))


def test_load_model_contract(tmp_path, capsys):
    """models/model.py:195-251: 'module.' prefix stripped, shape mismatches and missing keys keep the model's own
    tensors, extra keys dropped, optimizer resume replays the lr schedule."""
    import torch
    import synthdata as synth
    from cet_pick_b200.models.model import create_model, load_model, save_model
    m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
    sd = synth.unet_state_dict_torch(317, 4)
    own_hm = m.state_dict()["hm.weight"].clone()
    bad = {("module." + k): v for k, v in sd.items()}
    bad["module.hm.weight"] = torch.zeros(2, 32, 3, 1, 1)             # wrong shape -> skipped
    del bad["module.conv1.weight"]                                     # missing -> kept
    bad["module.extra.weight"] = torch.zeros(3)                        # unknown -> dropped
    own_c1 = m.state_dict()["conv1.weight"].clone()
    p = str(tmp_path / "c.pth")
    torch.save({"epoch": 7, "state_dict": bad}, p)
    m2 = load_model(m, p)
    out = capsys.readouterr().out
    assert "Skip loading parameter hm.weight" in out and "No param conv1.weight" in out and "Drop parameter extra.weight" in out
    got = m2.state_dict()
    assert torch.equal(got["hm.weight"], own_hm) and torch.equal(got["conv1.weight"], own_c1)
    assert torch.equal(got["unet.down_convs.0.conv1.weight"], sd["unet.down_convs.0.conv1.weight"])
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    save_model(p, 450, m, opt)
    m3, opt3, ep = load_model(m, p, optimizer=torch.optim.Adam(m.parameters(), lr=1.0), resume=True, lr=1e-3,
                              lr_step=[200, 400, 600])
    assert ep == 450 and abs(opt3.param_groups[0]["lr"] - 1e-5) < 1e-12


def test_exploration_driver_host_logic():
    """cet_pick_b200/simsiam_test_hm_3d.py: the patch normalisation is torchvision's ToPILImage -> ToTensor -> Normalize of
    PrefetchDatasetProj (simsiam_test_hm_3d.py:44-51) bit for bit; the border filter is the dataset's (:207)."""
    import torch
    import torchvision.transforms as T
    from cet_pick_b200 import simsiam_test_hm_3d as drv
    g = torch.Generator().manual_seed(0)
    p = torch.rand(7, 1, 32, 32, generator=g)
    p[0, 0, 0, 0], p[0, 0, 0, 1] = 1.0, 0.0               # the min-max normalised patches contain both ends
    mean, std = p.mean(), p.std()
    tr = T.Compose([T.ToPILImage(), T.ToTensor(), T.Normalize((mean), (std))])
    ref = torch.stack([tr(x) for x in p])
    assert torch.equal(drv.normalise_patches(p.clone(), mean, std), ref)
    pos = np.array([[17, 17, 12], [18, 17, 12], [18, 18, 12], [46, 50, 12], [47, 46, 12], [45, 47, 12], [30, 47, 12]])
    assert drv.keep_candidates(pos, (30, 64, 64), 32).tolist() == [[18, 17, 12], [18, 18, 12], [45, 47, 12], [30, 47, 12]]
    assert drv.keep_candidates(np.zeros((0, 3), int), (30, 64, 64), 32).shape == (0, 3)


def test_mrc_byte_order_mode12_extended_header_and_truncation(tmp_path):
    """read_mrc like mrcfile.open(path, permissive=True): big-endian files (machine stamp 0x11), mode 12 (float16), an
    extended header before the data, and a clear error for a file shorter than its header says."""
    from cet_pick_b200.utils import mrcio
    rng = np.random.default_rng(1)
    data = rng.standard_normal((3, 4, 5)).astype(np.float32)
    h = bytearray(1024)
    struct.pack_into(">4i", h, 0, 5, 4, 3, 2)
    h[212:216] = bytes([0x11, 0x11, 0, 0])
    p = str(tmp_path / "be.mrc")
    open(p, "wb").write(bytes(h) + data.astype(">f4").tobytes())
    out = mrcio.read_mrc(p)
    assert out.dtype == np.float32 and out.dtype.byteorder in "=<" and np.array_equal(out, data)
    half = data.astype(np.float16)
    h = bytearray(1024)
    struct.pack_into("<4i", h, 0, 5, 4, 3, 12)
    struct.pack_into("<i", h, 92, 160)                               # nsymbt: 160 bytes of extended header
    h[212:216] = bytes([0x44, 0x44, 0, 0])
    p = str(tmp_path / "f16.mrc")
    open(p, "wb").write(bytes(h) + b"\x07" * 160 + half.tobytes())
    out = mrcio.read_mrc(p)
    assert out.dtype == np.float16 and np.array_equal(out, half)
    p = str(tmp_path / "short.mrc")
    open(p, "wb").write(bytes(h) + b"\x07" * 160 + half.tobytes()[:-10])
    with pytest.raises(ValueError, match="truncated"):
        mrcio.read_mrc(p)
    h2 = bytearray(1024)                                              # no machine stamp at all: the sane order wins
    struct.pack_into(">4i", h2, 0, 5, 4, 3, 1)
    p = str(tmp_path / "nostamp.mrc")
    ints = rng.integers(-50, 50, size=(3, 4, 5)).astype(np.int16)
    open(p, "wb").write(bytes(h2) + ints.astype(">i2").tobytes())
    assert np.array_equal(mrcio.read_mrc(p), ints)


def _write_many(args):
    """worker of the test below: one 'rank' writing its tomograms into the shared output directory"""
    rank, path = args
    import types
    import torch
    from cet_pick_b200.detectors.tomo_det import TomodetDetector
    det = TomodetDetector.__new__(TomodetDetector)
    det.opt = types.SimpleNamespace(down_ratio=2, out_thresh=0.25, cutoff_z=1, compress=False, fiber=False, spike=False,
                                    with_score=False, nms=3)
    hm = torch.rand(1, 1, 6, 40, 44)
    dets = {2: [[50.5, 40.5, 2.0, 0.9, 0.9]], 3: [[30.5, 44.5, 3.0, 0.7, 0.7]]}
    for i in range(12):
        det.save_detection(hm, dets, path, None, name=f"r{rank}_t{i}")
    return rank


def test_two_ranks_write_into_the_same_new_directory(tmp_path):
    """torchrun ranks share <save_dir>/<out_id>: none may die on the directory another one is creating (round-1 advice)"""
    import multiprocessing as mp
    out = str(tmp_path / "exp" / "out")                               # neither level exists yet
    with mp.get_context("spawn").Pool(2) as pool:
        assert sorted(pool.map(_write_many, [(0, out), (1, out)])) == [0, 1]
    names = sorted(os.listdir(out))
    assert len(names) == 48 and "r0_t0.txt" in names and "r1_t11_hm.mrc" in names
    assert open(os.path.join(out, "r1_t3.txt")).read() == "50\t2\t40\n30\t3\t44\n"
