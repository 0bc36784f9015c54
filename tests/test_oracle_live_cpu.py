"""Live cross-check of the oracle against the UNMODIFIED reference imported from /root/reference (CPU; skipped on
boxes without the checkout, where the committed fixtures in tests/golden/ are the pin).  Random tie-free inputs beyond
the fixed fixtures: decode (all window sizes, fiber, reg), greedy distance suppression, the detector forward."""
import numpy as np
import pytest
import torch

import synthdata as synth
from oracle import decode_oracle as do
from oracle import refbridge
from oracle import unet_oracle as uo

pytestmark = pytest.mark.skipif(not refbridge.available(), reason="reference checkout not present")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("shape,kernel,fiber,K,seed", [
    ((5, 17, 23), 3, False, 60, 1), ((9, 30, 31), 5, False, 100, 2), ((4, 12, 40), 7, False, 33, 3),
    ((6, 24, 24), 3, True, 80, 4), ((3, 9, 9), 1, False, 243, 5), ((12, 20, 18), 3, False, 1, 6),
])
def test_decode_against_live_reference(shape, kernel, fiber, K, seed):
    d = refbridge.decode_module()
    hm = synth.heatmap_tiefree_np(*shape, seed)[None, None]
    ref = d.tomo_decode(torch.from_numpy(hm), kernel=kernel, reg=None, K=K, if_fiber=fiber).numpy()
    out = do.tomo_decode(hm, kernel, None, K, fiber)
    # rows beyond the last real peak have score 0 and, in torch.topk, unspecified positions
    n = int((ref[0, :, 3] > 0).sum())
    assert n == int((out[0, :, 3] > 0).sum()) and n > 0
    assert np.array_equal(bits(out[:, :n]), bits(ref[:, :n]))
    assert not out[0, n:, 3].any() and not ref[0, n:, 3].any()


def test_decode_reg_against_live_reference():
    d = refbridge.decode_module()
    D, H, W, K = 5, 14, 18, 40
    hm = np.stack([synth.heatmap_tiefree_np(D, H, W, s) for s in (7, 8)])[:, None]
    reg = (synth.uniform_np(3, 2 * 2 * D * H * W).reshape(2, 2, D, H, W) - 0.5).astype(np.float32)
    ref = d.tomo_decode(torch.from_numpy(hm), kernel=3, reg=torch.from_numpy(reg), K=K).numpy()
    assert np.array_equal(bits(do.tomo_decode(hm, 3, reg, K)), bits(ref))


@pytest.mark.parametrize("shape,dd,q", [((5, 16, 18), 3, 0.5), ((6, 20, 20), 6, 0.8), ((4, 30, 12), 9, 0.9)])
def test_greedy_nms_against_live_reference(shape, dd, q):
    d = refbridge.decode_module()
    x = synth.heatmap_tiefree_np(*shape, 11 + dd)
    thr = float(np.quantile(x, q))
    rs, rc = d.non_maximum_suppression_3d(x, dd, threshold=thr)
    s, c = do.greedy_distance_nms(x, dd, threshold=thr)
    assert np.array_equal(bits(s), bits(rs)) and np.array_equal(c, rc)


@pytest.mark.parametrize("n_blocks,shape", [(4, (3, 40, 44)), (5, (2, 32, 48))])
def test_forward_against_live_reference(n_blocks, shape):
    m = refbridge.create_model(f"unet_{n_blocks}", {"hm": 1, "proj": 32}, 32, last_k=3)
    sd = synth.unet_state_dict_torch(23, n_blocks)
    m.load_state_dict(sd)
    m.eval()
    x = torch.from_numpy(synth.tomogram_np(*shape, 4))[None]
    with torch.no_grad():
        ref = m(x)[-1]
        out = uo.forward(x, sd)
    assert (out["hm"] - ref["hm"]).abs().max().item() <= 1e-5
    assert (out["proj"] - ref["proj"]).abs().max().item() <= 1e-5


@pytest.mark.parametrize("shape,sigma,tilt", [((9, 20, 22), 0.8, False), ((7, 16, 30), 0.0, False), ((12, 14, 14), 2.2, False),
                                              ((5, 18, 20), 1.0, True), ((4, 24, 16), 0.0, True)])
def test_preprocess_against_live_reference(shape, sigma, tilt):
    """utils/loader.py:90-121 on arrays (no file involved): reconstruction and tilt branches."""
    refbridge.install()
    import cet_pick.utils.loader as rl
    from oracle import preproc_oracle as po
    rec = (synth.tomogram_np(*shape, 3).astype(np.float64) - 0.4) * 3.0
    ref = rl.preprocess(rec.copy(), denoise=sigma, is_tilt=tilt)
    if tilt:
        out = po.preprocess_tilt(rec.copy(), sigma)
        assert out.dtype == ref.dtype == np.float32 and np.abs(out - ref).max() <= 1.2e-7
    else:
        assert np.array_equal(po.preprocess(rec.copy(), sigma), ref)


@pytest.mark.parametrize("shape,params", [((33, 70, 50), (16, 24, 8, 12)), ((10, 30, 31), (32, 96, 16, 24))])
def test_patch_dataset_against_live_reference(shape, params):
    refbridge.install()
    from cet_pick.detectors.tomo_det_classify import PatchDataset
    from oracle import classify_oracle as co
    vol = synth.heatmap_tiefree_np(*shape, 2)
    ds = PatchDataset(vol, *params)
    for n in range(len(ds)):
        idx, x = ds[n]
        oi, ox, grid = co.patch(vol, n, *params)
        assert np.array_equal(idx, oi) and np.array_equal(x, ox) and tuple(grid) == tuple(ds.shape)


def test_training_step_against_live_reference():
    """Train-mode forward (batch-statistics BatchNorm, b > 1 branch) + PULoss + autograd of the oracle against the
    imported reference model and loss: loss, every parameter gradient and the updated running statistics."""
    from oracle import train_oracle as to
    refbridge.install()
    from cet_pick.models.loss import PULoss
    from cet_pick.models.utils import _sigmoid
    torch.manual_seed(0)
    m = refbridge.create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
    sd = {k: v.clone() for k, v in synth.unet_state_dict_torch(29, 4).items()}
    m.load_state_dict(sd)
    m.train()
    b, d, h, w = 2, 4, 36, 44
    x = torch.from_numpy(np.stack([synth.tomogram_np(d, h, w, 4 + i) for i in range(b)]))
    gt = torch.full((b, 1, d, h // 2, w // 2), -1.0)
    gt[0, 0, 1, 5, 7] = 1.0
    gt[1, 0, 2, 9, 3] = 1.0
    gt[0, 0, 1, 5, 8] = 0.6
    gt[1, 0, 2, 10, 3] = 0.3
    out = m(x)[-1]
    loss = PULoss(0.02)(_sigmoid(out["hm"]), gt)
    loss.backward()
    oloss, grads, ohm = to.training_step(x, gt, sd, 0.02)
    assert abs(float(loss) - float(oloss)) <= 1e-6 * max(1.0, abs(float(loss)))
    # models/utils.py:167-169 `_sigmoid` works in place: out["hm"] holds the clamped sigmoid by now
    assert (uo.sigmoid_clamp(ohm) - out["hm"].detach()).abs().max().item() <= 1e-6
    for k, p in m.named_parameters():
        if k.startswith("proj"):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        ref = p.grad
        err = (grads[k] - ref).abs().max().item()
        assert err <= 1e-5 * max(1e-3, ref.abs().max().item()) + 1e-9, (k, err, ref.abs().max().item())
    for k, v in m.named_buffers():
        if "running" in k:
            assert (sd[k] - v).abs().max().item() <= 1e-6, k
