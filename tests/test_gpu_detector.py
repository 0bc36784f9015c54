"""End-to-end detector mirror (detectors/tomo_det.py) on the GPU: pick files against the reference's
golden text, and run() through opts + checkpoint loading against the oracle pipeline."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _detector(opt_kw):
    import types
    from cet_pick_b200.detectors.tomo_det import TomodetDetector
    det = TomodetDetector.__new__(TomodetDetector)
    det.opt = types.SimpleNamespace(down_ratio=2, out_thresh=0.25, cutoff_z=3, compress=False, fiber=False,
                                    spike=False, with_score=False, nms=3, K=120)
    for k, v in opt_kw.items():
        setattr(det.opt, k, v)
    return det


@pytest.mark.parametrize("tag,kw", [("plain", {}), ("score", dict(with_score=True)), ("compress", dict(compress=True))])
def test_pick_files_match_reference_golden(golden, tmp_path, tag, kw):
    """decode on the GPU + post_process + save_detection reproduce the reference's `<name>.txt` byte for byte
    (fixture written by the unmodified reference, tests/golden/make_golden.py)."""
    import synthdata as synth
    from cet_pick_b200.models.decode import tomo_decode
    g = golden("pickfile")
    D, H, W = (int(v) for v in g["shape"])
    hm = synth.heatmap_tiefree_np(D, H, W, int(g["seed"]))[None, None]
    hm = ((hm - hm.min()) / (hm.max() - hm.min())).astype(np.float32)
    hm_t = torch.from_numpy(hm).cuda()
    dets = tomo_decode(hm_t, kernel=3, K=120)
    assert np.array_equal(dets.cpu().numpy().view(np.uint32), g["dets"].view(np.uint32))
    det = _detector(kw)
    preds, name = det.post_process(dets.clone(), {"name": ["tomoA"]}, z_dim_tot=D)
    det.save_detection(hm_t, preds, str(tmp_path), None, name=name)
    assert open(os.path.join(tmp_path, "tomoA.txt")).read() == str(g["txt_" + tag])
    assert os.path.exists(os.path.join(tmp_path, "tomoA_hm.mrc"))


def test_run_end_to_end_from_checkpoint(tmp_path, monkeypatch):
    """opts -> detector_factory -> load_model(checkpoint) -> run(): picks agree with the oracle pipeline
    (fp32 forward + decode) within one voxel for every pick whose score clears the K-th by the BF16 tolerance."""
    import synthdata as synth
    from cet_pick_b200.detectors.detector_factory import detector_factory
    from cet_pick_b200.opts import opts
    from oracle import decode_oracle as do
    from oracle import unet_oracle as uo
    sd = synth.unet_state_dict_torch(317, 4)
    ckpt = os.path.join(tmp_path, "model.pth")
    torch.save({"epoch": 3, "state_dict": {"module." + k: v for k, v in sd.items()}}, ckpt)
    monkeypatch.chdir(tmp_path)
    K = 60
    opt = opts().init(["semi", "--arch", "unet_4", "--load_model", ckpt, "--K", str(K), "--out_thresh", "0.0",
                       "--cutoff_z", "0", "--with_score", "--out_id", "out", "--exp_id", "e2e", "--gpus", "0"])
    os.makedirs(opt.save_dir, exist_ok=True)
    det = detector_factory[opt.task](opt)
    D, H, W = 12, 96, 128
    x = torch.from_numpy(synth.tomogram_np(D, H, W, 5))[None]
    ret = det.run(x, {"name": ["tomoB"]})
    assert set(ret) == {"tot_time", "load", "pre", "net", "dec"}
    lines = [ln.split("\t") for ln in open(os.path.join(opt.out_path, "tomoB.txt")).read().splitlines()]
    got = np.array([[float(v) for v in ln] for ln in lines])          # x, z, y, score
    with torch.no_grad():
        hm = uo.sigmoid_clamp(uo.forward(x, sd, want_proj=False)["hm"]).numpy()
    ref = do.tomo_decode(hm, 3, None, K)[0]                             # x+.25, y+.25, z, s, s (half-res)
    tol = 1e-2                     # stated BF16 heat-map tolerance (scores)
    margin = 4e-3                  # picks this far above the K-th score cannot drop out of the top-K (measured
    kth = ref[-1, 3]               # heat-map error is <= 2e-3); nearer ones may legitimately reorder
    strong = ref[ref[:, 3] > kth + margin]
    assert len(strong) > 0
    for r in strong:
        x2, y2, z = 2 * int(np.floor(r[0])), 2 * int(np.floor(r[1])), int(r[2])
        if not (20 < x2 < 2 * hm.shape[-1] - 20 and 20 < y2 < 2 * hm.shape[-2] - 20):
            continue                                                  # the writer drops the 20-pixel border
        d = np.abs(got[:, :3] - np.array([x2, z, y2])).max(axis=1)
        j = int(np.argmin(d))
        assert d[j] <= 2, (r, got[j])                                  # 1 voxel at half resolution = 2 input pixels
        assert abs(got[j, 3] - r[3]) <= tol


def test_pipeline_mrc_to_pick_file(tmp_path, monkeypatch):
    """The refinement-step pipeline of test.py with every arithmetic stage on the device: an int16 MRC
    reconstruction -> load_tomos_from_list (order xzy, --compress, --gauss 0.8; csrc/preproc.cu) -> detector.run
    (forward + decode) -> `<name>.txt`.  The pre-processed volume equals the oracle's 256-level volume, and the
    heat-map written next to the picks is within the BF16 tolerance of the fp32 oracle forward on it."""
    import synthdata as synth
    from cet_pick_b200.detectors.detector_factory import detector_factory
    from cet_pick_b200.opts import opts
    from cet_pick_b200.utils import loader, mrcio
    from oracle import preproc_oracle as po
    from oracle import unet_oracle as uo
    sd = synth.unet_state_dict_torch(317, 4)
    ckpt = os.path.join(tmp_path, "model.pth")
    torch.save({"epoch": 1, "state_dict": sd}, ckpt)
    monkeypatch.chdir(tmp_path)
    # stored MRC array (nz', ny, nx) = (x, z, y) for order 'xzy': 64 x-columns, 24 z-slices, 96 y-rows
    raw = (synth.tomogram_np(64, 24, 96, 11) * 900 - 300).astype(np.int16)
    path = os.path.join(tmp_path, "tomoC.mrc")
    hdr_src = raw.astype(np.float32)
    mrcio.write_mrc(path, hdr_src)                     # float32 file with integer-valued voxels
    opt = opts().init(["semi", "--arch", "unet_4", "--load_model", ckpt, "--K", "40", "--out_thresh", "0.0",
                       "--cutoff_z", "0", "--with_score", "--compress", "--gauss", "0.8", "--out_id", "out",
                       "--exp_id", "pipe", "--gpus", "0"])
    os.makedirs(opt.save_dir, exist_ok=True)
    ims = loader.load_tomos_from_list(["tomoC"], [path], order=opt.order, compress=opt.compress, denoise=opt.gauss,
                                      dtype=torch.float32)
    vol = ims["tomoC"]                                  # (12, 64, 96) float32 on the device
    ref_vol = po.preprocess(po.load_rec(hdr_src, order="xzy", compress=True), denoise=0.8).astype(np.float32)
    assert tuple(vol.shape) == ref_vol.shape == (12, 64, 96)
    assert np.array_equal(vol.cpu().numpy(), ref_vol)
    det = detector_factory[opt.task](opt)
    ret = det.run(vol[None], {"name": ["tomoC"]})
    assert ret["tot_time"] > 0
    lines = open(os.path.join(opt.out_path, "tomoC.txt")).read().splitlines()
    assert len(lines) > 0 and all(len(ln.split("\t")) == 4 for ln in lines)
    hm_file = mrcio.read_mrc(os.path.join(opt.out_path, "tomoC_hm.mrc"))           # axes (H', D, W')
    with torch.no_grad():
        hm_ref = uo.sigmoid_clamp(uo.forward(torch.from_numpy(ref_vol)[None], sd, want_proj=False)["hm"]).numpy()[0, 0]
    assert np.abs(np.swapaxes(hm_file, 1, 0) - hm_ref).max() <= 1e-2


def test_command_line_entry(tmp_path, monkeypatch):
    """`python -m cet_pick_b200.test semi ...` (the reference's test.py command line): image list -> MRC -> GPU
    pre-processing -> detector -> pick files and opt.txt, for two tomograms."""
    import synthdata as synth
    from cet_pick_b200 import test as cli
    from cet_pick_b200.opts import opts
    from cet_pick_b200.utils import mrcio
    sd = synth.unet_state_dict_torch(317, 4)
    ckpt = os.path.join(tmp_path, "model.pth")
    torch.save({"epoch": 1, "state_dict": sd}, ckpt)
    monkeypatch.chdir(tmp_path)
    paths = []
    for i in range(2):
        raw = (synth.tomogram_np(64, 24, 96, 20 + i) * 900 - 300).astype(np.float32)      # stored as (x, z, y): order xzy
        paths.append(os.path.join(tmp_path, f"t{i}.mrc"))
        mrcio.write_mrc(paths[-1], raw)
    lst = os.path.join(tmp_path, "test_images.txt")
    with open(lst, "w") as f:
        f.write("image_name\trec_path\n" + "".join(f"tomo{i}\t{p}\n" for i, p in enumerate(paths)))
    K = 400
    opt = opts().init(["semi", "--arch", "unet_4", "--load_model", ckpt, "--K", str(K), "--out_thresh", "0.0", "--cutoff_z", "0",
                       "--with_score", "--compress", "--gauss", "0.8", "--test_img_txt", lst, "--out_id", "out",
                       "--exp_id", "cli", "--gpus", "0"])
    stats = cli.test(opt)
    assert len(stats["tot_time"]) == 2 and os.path.exists(os.path.join(opt.save_dir, "opt.txt"))
    n_lines = 0
    for i in range(2):
        lines = open(os.path.join(opt.out_path, f"tomo{i}.txt")).read().splitlines()
        assert len(lines) <= K and all(len(ln.split("\t")) == 4 for ln in lines)      # the writer drops the 20-pixel border
        n_lines += len(lines)
        assert mrcio.read_mrc(os.path.join(opt.out_path, f"tomo{i}_hm.mrc")).shape == (32, 12, 48)     # (H', D, W')
    assert n_lines > 0


def test_async_write_and_uint8_levels_give_identical_files(tmp_path, monkeypatch):
    """run() with the writer threads (set_async_write) and with the tomogram shipped as uint8 levels must leave the
    same `<name>.txt` and the same heat-map bytes on disk as the blocking float32 path."""
    import synthdata as synth
    from cet_pick_b200.detectors.detector_factory import detector_factory
    from cet_pick_b200.opts import opts
    sd = synth.unet_state_dict_torch(317, 4)
    ckpt = os.path.join(tmp_path, "model.pth")
    torch.save({"epoch": 1, "state_dict": sd}, ckpt)
    monkeypatch.chdir(tmp_path)
    opt = opts().init(["semi", "--arch", "unet_4", "--load_model", ckpt, "--K", "60", "--out_thresh", "0.0",
                       "--cutoff_z", "0", "--with_score", "--out_id", "out", "--exp_id", "aw", "--gpus", "0"])
    det = detector_factory[opt.task](opt)
    D, H, W = 10, 96, 128
    vols = [synth.tomogram_np(D, H, W, 20 + i) for i in range(3)]
    opt.out_path = os.path.join(tmp_path, "sync")
    for i, v in enumerate(vols):
        det.run(torch.from_numpy(v)[None], {"name": [f"t{i}"]})
    opt.out_path = os.path.join(tmp_path, "async")
    det.set_async_write(True, threads=2)
    for i, v in enumerate(vols):
        q = torch.from_numpy(np.rint(v * 255.0).astype(np.uint8))
        det.run(q[None], {"name": [f"t{i}"], "level_values": None})
    det.flush()
    det.set_async_write(False)
    for i in range(3):
        for suffix in (".txt", "_hm.mrc"):
            a = open(os.path.join(tmp_path, "sync", f"t{i}{suffix}"), "rb").read()
            b = open(os.path.join(tmp_path, "async", f"t{i}{suffix}"), "rb").read()
            assert a == b, f"t{i}{suffix} differs between the blocking and the asynchronous path"
