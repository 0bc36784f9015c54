"""The flag system against the unmodified reference's opts.py (CPU; skipped where /root/reference is absent):
same defaults and same derived fields for the command lines of docs/refine.md."""
import os

import pytest

from oracle import refbridge

pytestmark = pytest.mark.skipif(not refbridge.available(), reason="reference checkout not present")

CASES = [
    ["semi", "--arch", "unet_4", "--load_model", "m.pth", "--K", "900", "--compress", "--gauss", "0.8",
     "--test_img_txt", "t.txt", "--out_id", "out", "--with_score"],
    ["semi"],
    ["semiclass", "--nms", "5", "--out_thresh", "0.4", "--cutoff_z", "7", "--gpus", "0"],
    ["semi", "--arch", "unet_5", "--fiber", "--exp_id", "abc", "--down_ratio", "2", "--order", "zxy"],
    # the exploration step (simsiam_test_hm_3d.py): head width 128 by default, DoG sigmas, candidate box
    ["simsiam3d", "--arch", "simsiam3d_18", "--load_model", "e.pth", "--bbox", "32", "--dog", "2.5,5", "--gauss", "0.8",
     "--compress", "--test_img_txt", "t.txt", "--exp_id", "explore"],
    ["simsiam", "--arch", "simsiam2d_18", "--bbox", "24", "--spike", "--distance_cutoff", "12"],
]


@pytest.mark.parametrize("argv", CASES)
def test_parse_matches_reference(argv, tmp_path, monkeypatch):
    refbridge.install()
    from cet_pick.opts import opts as ref_opts
    from cet_pick_b200.opts import opts as our_opts
    monkeypatch.chdir(tmp_path)
    ref = vars(ref_opts().parse(list(argv)))
    ours = vars(our_opts().parse(list(argv)))
    skip = {"root_dir", "data_dir", "exp_dir", "save_dir", "debug_dir", "out_path", "device"}   # absolute paths of each checkout
    common = (set(ref) & set(ours)) - skip
    assert len(common) >= 60
    diff = {k: (ref[k], ours[k]) for k in common if ref[k] != ours[k]}
    assert not diff, diff
    for k in ("save_dir", "debug_dir", "out_path"):          # same layout below the respective root
        if k in ref and k in ours:
            assert os.path.relpath(ref[k], ref["root_dir"]) == os.path.relpath(ours[k], ours["root_dir"])


def test_gather_helpers_match_reference():
    """models/utils.py:171-193 (`_gather_feat`, `_transpose_and_gather_feat`) on CPU tensors."""
    import torch
    u = refbridge.utils_module()
    from cet_pick_b200.models import utils as ours
    g = torch.Generator().manual_seed(0)
    feat = torch.randn(2, 3, 4, 5, 6, generator=g)
    ind = torch.randint(0, 4 * 5 * 6, (2, 7), generator=g)
    assert torch.equal(ours._transpose_and_gather_feat(feat, ind), u._transpose_and_gather_feat(feat, ind))
    flat = torch.randn(2, 50, 3, generator=g)
    idx = torch.randint(0, 50, (2, 9), generator=g)
    mask = torch.rand(2, 9, generator=g) > 0.4
    assert torch.equal(ours._gather_feat(flat, idx), u._gather_feat(flat, idx))
    assert torch.equal(ours._gather_feat(flat, idx, mask), u._gather_feat(flat, idx, mask))
