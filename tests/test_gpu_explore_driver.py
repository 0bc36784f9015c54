"""Exploration-step driver (cet_pick_b200/simsiam_test_hm_3d.py, mirror of cet_pick/simsiam_test_hm_3d.py:136-195): candidate
generator -> slab-sum patches -> PrefetchDatasetProj normalisation -> SimSiam encoder (D = 1 inputs) -> all_output_info.npz."""
import os
import types

import numpy as np
import pytest
import torch

import synthdata as synth

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))


def _model():
    from cet_pick_b200.models.model import create_model
    m = create_model("simsiam3d_18", {"proj": 128, "pred": 128}, 128)
    m.load_state_dict(synth.simsiam3d_state_dict_torch(5))
    return m.cuda().eval()


def test_candidates_to_embeddings_in_memory():
    from cet_pick_b200 import simsiam_test_hm_3d as drv
    from oracle import simsiam_oracle as so
    opt = types.SimpleNamespace(bbox=32, dog=[2.5, 5.0])
    rec = synth.tomogram_np(40, 128, 128, 3)
    patches, coords = drv.candidate_patches(rec, opt)
    n = patches.shape[0]
    assert n >= 4 and patches.shape == (n, 1, 32, 32) and coords.shape == (n, 3)
    mx = 32 // 1.8
    assert ((coords[:, 0] > mx) & (coords[:, 0] < 128 - mx) & (coords[:, 1] >= mx) & (coords[:, 1] <= 128 - mx)).all()
    # the patch of a candidate = z-slab sum of its window, min-max normalised (dataset :117-128)
    x, y, z = (int(v) for v in coords[0])
    win = rec[z - 1:z + 2, y - 16:y + 16, x - 16:x + 16].astype(np.float64).sum(0)
    win = (win - win.min()) / (win.max() - win.min())
    assert np.abs(patches[0, 0].cpu().numpy() - win.astype(np.float32)).max() <= 1e-6
    out = drv.embed(_model(), patches, coords, ["tomoX"] * n)
    assert out["proj"].shape == (n, 256) and out["pred"].shape == (n, 256) and out["subvol"].shape == (n, 1, 32, 32)
    assert out["name"].shape == (n,) and np.array_equal(out["coords"], coords)
    sd = synth.simsiam3d_state_dict_torch(5)
    with torch.no_grad():
        ref = so.forward_test(torch.from_numpy(out["subvol"]), sd)        # (n, 1, 32, 32): D = 1 per sub-volume
    for k in ("proj", "pred"):
        e = rel_err(out[k], ref[k].numpy())
        print(f"driver {k}: relative L2 err {e:.3e} over {n} candidates")
        assert e <= 1e-2


def test_driver_writes_the_reference_npz(tmp_path):
    from cet_pick_b200 import simsiam_test_hm_3d as drv
    from cet_pick_b200.opts import opts
    from cet_pick_b200.utils.mrcio import write_mrc
    vol = (synth.tomogram_np(128, 128, 128, 8) * 200).astype(np.float32)
    write_mrc(str(tmp_path / "t0.mrc"), vol)
    with open(tmp_path / "list.txt", "w") as f:
        f.write("image_name\trec_path\nt0\t%s\n" % (tmp_path / "t0.mrc"))
    torch.save({"epoch": 0, "state_dict": synth.simsiam3d_state_dict_torch(5)}, str(tmp_path / "m.pth"))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        opt = opts().init(["simsiam3d", "--arch", "simsiam3d_18", "--load_model", str(tmp_path / "m.pth"), "--bbox", "32",
                           "--gauss", "0.8", "--test_img_txt", str(tmp_path / "list.txt"), "--exp_id", "run"])
        out_file = drv.test(opt)
    finally:
        os.chdir(cwd)
    z = np.load(out_file)
    n = z["proj"].shape[0]
    assert n >= 1 and set(z.files) == {"proj", "pred", "name", "coords", "subvol"}
    assert z["pred"].shape == (n, 256) and z["coords"].shape == (n, 3) and z["subvol"].shape == (n, 1, 32, 32)
    assert (z["name"] == "t0").all() and np.isfinite(z["proj"]).all()
