"""Device pieces of the refinement training step (csrc/train.cu, cet_pick_b200/trains/step.py) against the
reference-generated fixture `train_losses` (values AND gradients of cet_pick/models/loss.py) and torch.optim.Adam."""
import numpy as np
import pytest
import torch

import synthdata as synth

pytestmark = pytest.mark.gpu


def fixture_inputs(g):
    D, H, W = [int(v) for v in g["shape"]]
    pred0 = (0.02 + 0.96 * synth.uniform_np(int(g["seed"]), D * H * W).reshape(1, D, H, W)).astype(np.float32)
    return pred0, g["gt"]


@pytest.mark.parametrize("tag,tau", [("pu_tau01", 0.1), ("pu_tau06", 0.6)])
def test_pu_loss_value_and_gradient_vs_reference_golden(golden, tag, tau):
    from cet_pick_b200.trains.step import pu_loss
    g = golden("train_losses")
    pred0, gt = fixture_inputs(g)
    loss, grad, stats = pu_loss(torch.from_numpy(pred0).cuda(), torch.from_numpy(gt).cuda(), tau, apply_sigmoid=False)
    ref_l, ref_g = float(g[tag + "_loss"]), g[tag + "_grad"]
    assert abs(float(loss) - ref_l) <= 1e-5 * max(1.0, abs(ref_l))
    assert np.abs(grad.cpu().numpy() - ref_g).max() <= 1e-6 + 1e-4 * np.abs(ref_g).max()
    assert int(stats[3]) == int((gt == 1).sum())


def test_pu_loss_through_sigmoid_vs_oracle_autograd():
    """the trainer's actual chain: logits -> _sigmoid (clamped) -> PULoss; gradient w.r.t. the logits, zero where clamped"""
    from cet_pick_b200.trains.step import pu_loss
    from oracle import train_oracle as to
    from oracle import unet_oracle as uo
    rng = np.random.default_rng(4)
    D, H, W = 5, 33, 47
    logits = rng.normal(0, 3.0, size=(1, 1, D, H, W)).astype(np.float32)
    logits[0, 0, 0, 0, :4] = [25.0, -25.0, 12.0, -12.0]                     # clamped entries
    gt = -np.ones((1, 1, D, H, W), np.float32)
    gt[0, 0, 2, 10:14, 20:24] = 0.4
    gt[0, 0, 2, 12, 22] = 1.0
    gt[0, 0, 3, 5, 5] = 1.0
    x = torch.from_numpy(logits).requires_grad_(True)
    ref = to.pu_focal_loss(uo.sigmoid_clamp(x), torch.from_numpy(gt), 0.3)
    ref.backward()
    loss, grad, _ = pu_loss(torch.from_numpy(logits).cuda(), torch.from_numpy(gt).cuda(), 0.3)
    assert abs(float(loss) - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    assert np.abs(grad.cpu().numpy() - x.grad.numpy()).max() <= 1e-7 + 1e-4 * np.abs(x.grad.numpy()).max()
    assert float(grad[0, 0, 0, 0, 0]) == 0.0 and float(grad[0, 0, 0, 0, 1]) == 0.0


def test_pu_loss_without_positives_raises():
    from cet_pick_b200.trains.step import pu_loss
    with pytest.raises(ValueError):
        pu_loss(torch.zeros(1, 1, 2, 4, 4, device="cuda"), -torch.ones(1, 1, 2, 4, 4, device="cuda"), 0.1)


def test_consistency_loss_vs_reference_golden(golden):
    from cet_pick_b200.trains.step import consistency_loss
    g = golden("train_losses")
    pred0, _ = fixture_inputs(g)
    a = torch.from_numpy(pred0).cuda()
    b = torch.from_numpy(pred0[:, :, :, ::-1].copy()).cuda()
    loss, grad = consistency_loss(a, b)
    assert abs(float(loss) - float(g["cons_loss"])) <= 1e-7
    assert np.abs(grad.cpu().numpy() - g["cons_grad"]).max() <= 1e-9


def test_flat_bucket_adam_matches_torch_adam():
    """five steps of the fused Adam over the flat bucket of the real detector's parameters vs torch.optim.Adam"""
    from cet_pick_b200.models.model import create_model
    from cet_pick_b200.trains.step import FlatBucket
    torch.manual_seed(1)
    m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3).cuda()
    ref_params = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    opt = torch.optim.Adam(ref_params, lr=1e-3)
    bucket = FlatBucket(m)
    assert bucket.numel == sum(p.numel() for p in ref_params)
    g = torch.Generator(device="cuda").manual_seed(2)
    for step in range(5):
        grads = [torch.randn(p.shape, device="cuda", generator=g) * 0.01 for p in ref_params]
        for p, gr, rp in zip(m.parameters(), grads, ref_params):
            p.grad.copy_(gr)                      # the bucket's gradient views
            rp.grad = gr.clone()
        opt.step()
        bucket.adam_step(1e-3)
    for p, rp in zip(m.parameters(), ref_params):
        assert torch.allclose(p.data, rp.data, rtol=1e-5, atol=1e-7)
    # the parameters the model (and the next plan build) sees ARE the bucket
    assert next(m.parameters()).data.data_ptr() == bucket.params.data_ptr()
