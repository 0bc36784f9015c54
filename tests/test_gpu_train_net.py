"""Training-mode layers (csrc/train_net.cu) and the whole training step (trains/engine.py) against torch autograd.

Oracle: `oracle/train_oracle.training_step` (pinned to the unmodified reference model + PULoss + autograd by
tests/test_oracle_live_cpu.py::test_training_step_against_live_reference), run here in FLOAT64 on the same GPU.
Tolerance of the whole step: loss within 1e-5; every parameter gradient, as max-abs error relative to that tensor's
largest gradient, within max(5e-3, 3 x the error torch's OWN fp32 path makes against the same float64 oracle) and never
above 5e-2.  The backward of this network through ~20 batch-statistics BatchNorms amplifies fp32 rounding (a few ReLU
masks of near-zero activations flip): at 128^3 the error grows from 1e-7 at the head to ~2e-2 in the bottom block for
torch-fp32/cuDNN and for these kernels alike (measured: ours <= torch's on 55 of 62 tensors), so a fixed 1e-3 against
float64 is not reachable in fp32 by either.  In the default TF32 mode (tensor-core 3x3 convolutions and weight gradients,
the arithmetic of the reference's own run) the same amplification gives 0.3-0.4 in the bottom block for torch-TF32/cuDNN and
for these kernels alike, so there the only yardstick is torch-TF32's own error against float64 (no absolute cap).
The single layers are compared with torch's fp32 ops at 1e-4 .. 2e-5.
SURVEY.md 8f-4; BASELINE.json configs[4] names 128^3 crops: `test_training_step_128_cube`.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synthdata as synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _ops(tf32=None):
    """tf32: None = leave the library's mode (default on), else switch the tensor-core weight gradient on / off"""
    from cet_pick_b200 import _lib
    from cet_pick_b200.trains import engine as E
    if tf32 is not None:
        _lib.check(_lib.lib().cetpick_train_set_tf32(int(tf32)), "cetpick_train_set_tf32")
    return E, E.Ops(torch.device("cuda", torch.cuda.current_device()))


@pytest.fixture(autouse=True)
def _tf32_wgrad_default():
    yield
    from cet_pick_b200 import _lib
    _lib.lib().cetpick_train_set_tf32(1)


def _view(E, t):
    n, c, h, w = t.shape
    return E.View(t, n, c, h, w)


def rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


@pytest.mark.parametrize("name,cin,cout,shape,zdepth", [
    ("CONV3", 16, 32, (6, 33, 41), 1), ("CONV3", 64, 24, (3, 17, 20), 1), ("CONV3", 48, 80, (3, 9, 11), 1),
    ("CONV3", 128, 256, (20, 8, 8), 1), ("CONV1", 32, 32, (5, 19, 23), 1),
    ("STEM", 1, 16, (4, 37, 50), 1), ("HEAD3D", 32, 32, (8, 21, 27), 4), ("HM", 32, 1, (6, 14, 18), 3)])
@pytest.mark.parametrize("tf32", [False, True])
def test_conv_forward_dgrad_wgrad(name, cin, cout, shape, zdepth, tf32):
    """conv_f32_kernel / flip_weights + conv / weight-gradient kernels against F.conv2d / F.conv3d and autograd.
    tf32 = False: fp32 FMA everywhere (2e-5 / 1e-4); True: the wide 3x3 layers and the wide weight gradients contract on
    the tensor cores with TF32 operands (2e-3 of the largest value: 10-bit mantissas, fp32 accumulation)."""
    E, o = _ops(tf32)
    spec = getattr(E, name)
    g = torch.Generator(device="cuda").manual_seed(3)
    N, H, W = shape
    x = torch.randn((N, cin, H, W), device="cuda", generator=g, requires_grad=True)
    w = (torch.randn((cout, cin) + spec.k, device="cuda", generator=g) * 0.2).requires_grad_(True)
    if spec.k[0] == 1:
        ref = F.conv2d(x, w[:, :, 0], stride=spec.stride, padding=spec.pad[1:], dilation=spec.dil[1:])
    else:   # 3-D over crops of zdepth slices: (N/zd, C, zd, H, W)
        x5 = x.view(N // zdepth, zdepth, cin, H, W).permute(0, 2, 1, 3, 4)
        r5 = F.conv3d(x5, w, padding=spec.pad, dilation=spec.dil)
        ref = r5.permute(0, 2, 1, 3, 4).reshape(N, cout, r5.shape[-2], r5.shape[-1])
    gy = torch.randn(ref.shape, device="cuda", generator=g)
    gy[gy.abs() < 0.3] = 0.0                                   # exact zeros like a ReLU-masked gradient
    ref.backward(gy)
    y = E.new_view(N, cout, ref.shape[2], ref.shape[3], x.device)
    o.conv(_view(E, x.detach()), w.detach(), None, y, spec, zdepth)
    assert rel(y.t, ref.detach()) <= (2e-3 if tf32 else 2e-5)
    dw = torch.zeros_like(w)
    o.wgrad(_view(E, x.detach()), _view(E, gy), dw, spec, zdepth)
    assert rel(dw, w.grad) <= (2e-3 if tf32 else 1e-4)
    if spec.stride == 1:
        dx = E.new_view(N, cin, H, W, x.device)
        o.dgrad(_view(E, gy), w.detach(), dx, spec, zdepth)
        assert rel(dx.t, x.grad) <= (2e-3 if tf32 else 2e-5)


@pytest.mark.parametrize("cin,cout,shape,crop", [(64, 32, (5, 9, 11), (18, 22)), (32, 16, (3, 10, 7), (19, 13))])
@pytest.mark.parametrize("tf32", [False, True])
def test_upconv_forward_backward(cin, cout, shape, crop, tf32):
    """ConvTranspose2d(2, 2) + bias with the autocrop of unet.py:285-292; data gradient = stride-2 2x2 conv, weight
    gradient = the generic kernel with the roles of input and output-gradient swapped."""
    E, o = _ops(tf32)
    g = torch.Generator(device="cuda").manual_seed(4)
    N, H, W = shape
    x = torch.randn((N, cin, H, W), device="cuda", generator=g, requires_grad=True)
    w = (torch.randn((cin, cout, 2, 2), device="cuda", generator=g) * 0.2).requires_grad_(True)
    bias = torch.randn(cout, device="cuda", generator=g).requires_grad_(True)
    ref = F.conv_transpose2d(x, w, bias, stride=2)[:, :, :crop[0], :crop[1]]
    gy = torch.randn(ref.shape, device="cuda", generator=g)
    ref.backward(gy)
    y = E.new_view(N, cout, crop[0], crop[1], x.device)
    o.upconv(_view(E, x.detach()), w.detach(), bias.detach(), y)
    assert rel(y.t, ref.detach()) <= 2e-5
    gyv = _view(E, gy.contiguous())
    dx = E.new_view(N, cin, H, W, x.device)
    o.conv(gyv, w.detach(), None, dx, E.UP_DGRAD, 1)
    assert rel(dx.t, x.grad) <= 2e-5
    dw = torch.zeros_like(w)
    o.wgrad(gyv, _view(E, x.detach()), dw, E.UP_DGRAD, 1)
    assert rel(dw, w.grad) <= (2e-3 if tf32 else 1e-4)
    db = torch.zeros_like(bias)
    o.channel_sum(gyv, db)
    assert rel(db, bias.grad) <= 1e-5


@pytest.mark.parametrize("relu", [True, False])
def test_batchnorm_train_forward_backward_strided(relu):
    """Batch statistics, running-stat update, ReLU mask; output and gradient live in a channel slice of wider buffers."""
    E, o = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    N, Cc, H, W = 6, 24, 13, 17
    x = (torch.randn((N, Cc, H, W), device="cuda", generator=g) * 2 + 0.5).requires_grad_(True)
    gamma = torch.rand(Cc, device="cuda", generator=g).add_(0.5).requires_grad_(True)
    beta = torch.randn(Cc, device="cuda", generator=g).requires_grad_(True)
    rm, rv = torch.randn(Cc, device="cuda", generator=g), torch.rand(Cc, device="cuda", generator=g) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = F.batch_norm(x, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)
    if relu:
        ref = F.relu(ref)
    gy = torch.randn(ref.shape, device="cuda", generator=g)
    ref.backward(gy)
    wide = E.new_view(N, 2 * Cc, H, W, x.device, zero=True)
    y = wide.chan(Cc, Cc)
    save = o.bn(_view(E, x.detach()), y, gamma.detach(), beta.detach(), rm, rv, relu=relu)
    assert rel(y.tensor(), ref.detach()) <= 1e-5
    assert float(wide.t[:, :Cc].abs().max()) == 0.0
    assert rel(rm, rm_ref) <= 1e-6 and rel(rv, rv_ref) <= 1e-5
    gwide = E.new_view(N, 2 * Cc, H, W, x.device, zero=True)
    gwide.t[:, Cc:] = gy
    dx = E.new_view(N, Cc, H, W, x.device)
    dg, db = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    o.bn_bwd(_view(E, x.detach()), y, gwide.chan(Cc, Cc), dx, gamma.detach(), save, dg, db, relu=relu)
    assert rel(dx.t, x.grad) <= 1e-4
    assert rel(dg, gamma.grad) <= 1e-5 and rel(db, beta.grad) <= 1e-5


@pytest.mark.parametrize("shape", [(3, 8, 10, 12), (2, 5, 11, 9)])
def test_maxpool_ceil_forward_backward(shape):
    E, o = _ops()
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.relu(torch.randn(shape, device="cuda", generator=g)).requires_grad_(True)     # ties at 0 like after a ReLU
    ref = F.max_pool2d(x, 2, ceil_mode=True)
    gy = torch.randn(ref.shape, device="cuda", generator=g)
    ref.backward(gy)
    N, Cc, H, W = shape
    y = E.new_view(N, Cc, ref.shape[2], ref.shape[3], x.device)
    o.pool(_view(E, x.detach()), y)
    assert torch.equal(y.t, ref.detach())
    dx = E.new_view(N, Cc, H, W, x.device, zero=True)
    dx.t.fill_(1.0)
    o.pool_bwd(_view(E, x.detach()), _view(E, gy), dx, accumulate=True)
    # where several zeros tie, ATen and this kernel both give the gradient to the first one
    assert torch.equal(dx.t - 1.0, x.grad) or rel(dx.t - 1.0, x.grad) <= 1e-6


def _labels(b, d, h, w, seed):
    rng = np.random.default_rng(seed)
    gt = np.full((b, 1, d, h, w), -1.0, np.float32)
    for bi in range(b):
        for _ in range(max(2, d * h * w // 4000)):
            z, y, x = rng.integers(0, d), rng.integers(1, h - 1), rng.integers(1, w - 1)
            gt[bi, 0, z, y, x] = 1.0
            for dy, dx in ((0, 1), (1, 0), (0, -1), (-1, 0)):
                if gt[bi, 0, z, y + dy, x + dx] < 1.0:
                    gt[bi, 0, z, y + dy, x + dx] = 0.6
    return torch.from_numpy(gt)


def _step_vs_oracle(b, d, h, w, tau, tol, with_aug=False, tf32=False):
    from oracle import train_oracle as to
    from cet_pick_b200.models.model import create_model
    from cet_pick_b200.trains.engine import DetectorTrainer
    from cet_pick_b200 import _lib
    _lib.check(_lib.lib().cetpick_train_set_tf32(int(tf32)), "cetpick_train_set_tf32")
    sd = {k: v.cuda() for k, v in synth.unet_state_dict_torch(41, 4).items()}
    m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
    m.load_state_dict(sd)
    m = m.cuda()
    tr = DetectorTrainer(m, tau=tau)
    x = torch.stack([synth.tomogram_torch(d, h, w, seed=10 + i, device="cuda") for i in range(b)])
    gt = _labels(b, d, (h - 1) // 2 + 1, (w - 1) // 2 + 1, 3).cuda()
    x_aug = x.flip(-1).contiguous() if with_aug else None
    tr.zero_grad()
    loss, logits = tr.forward_backward(x, gt, x_aug=x_aug, want_logits=True)
    sdo = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    oloss, grads, ohm = to.training_step(x.double(), gt.double(), sdo, tau)
    if with_aug:
        from oracle import unet_oracle as uo
        with torch.no_grad():
            uo.forward(x_aug.double(), sdo, want_proj=False, train=True)        # moves the running statistics only
    # TF32 mode: forward and backward round their tensor-core operands to 10-bit mantissas, like the reference's own run
    assert abs(float(loss) - float(oloss)) <= (2e-3 if tf32 else 1e-5) * max(1.0, abs(float(oloss)))
    assert float((logits - ohm.view_as(logits)).abs().max()) <= (1e-2 if tf32 else 1e-4) * max(1.0, float(ohm.abs().max()))
    # a bias in front of a batch-statistics BatchNorm (the transposed convs') has NO gradient mathematically: both sides
    # hold rounding noise there, so the error of a tensor is taken relative to max(its own largest gradient, 1e-4 of the
    # largest gradient of the whole model)
    gmax = max(float(g.abs().max()) for g in grads.values())

    def errors(get):
        out = []
        for k, p in m.named_parameters():
            if not k.startswith("proj"):
                out.append((float((get(k, p) - grads[k]).abs().max()) / max(float(grads[k].abs().max()), 1e-4 * gmax), k))
        return sorted(out, reverse=True)

    for k, p in m.named_parameters():
        if k.startswith("proj"):
            assert float(p.grad.abs().max()) == 0.0
    errs = errors(lambda k, p: p.grad.double())
    msg = (f"training step {b}x{d}x{h}x{w}: loss {float(loss):.6f}, worst gradient errors vs float64 "
           f"{[(k, f'{e:.1e}') for e, k in errs[:3]]}, largest gradient {gmax:.3e}, {tr.stats['launches']} launches")
    # torch's own path in the same arithmetic class (fp32 / TF32 cuDNN) against the same float64 oracle: the co-reference
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    _, g32, _ = to.training_step(x, gt, {k: v.clone() for k, v in sd.items()}, tau)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    e32 = errors(lambda k, p: g32[k].double())
    msg += f"; torch {'TF32' if tf32 else 'fp32 (TF32 off)'} (cuDNN) by the same measure: {[(k, f'{e:.1e}') for e, k in e32[:3]]}"
    print(msg)
    t32 = {k: e for e, k in e32}
    for e, k in errs:
        assert e <= max(tol, 3.0 * t32[k]) and (tf32 or e <= 5e-2), (k, e, t32[k])
    for k, v in m.named_buffers():
        if "running" in k:
            assert rel(v.double(), sdo[k]) <= (5e-3 if tf32 else 1e-4), k     # running statistics of TF32 conv outputs
    return tr, m, x, gt


@pytest.mark.parametrize("tf32", [False, True])
@pytest.mark.parametrize("b,d,h,w,aug", [(1, 6, 32, 48, False), (2, 5, 38, 42, True)])
def test_training_step_small_vs_autograd(b, d, h, w, aug, tf32):
    """Even and odd (ceil-mode pool + autocrop) sizes, one crop (b == 1 branch) and two (b > 1 branch, 3-D taps must not
    cross crops), optionally the second train-mode forward of the augmented view."""
    _step_vs_oracle(b, d, h, w, 0.02, 2e-2 if tf32 else 5e-3, aug, tf32=tf32)


@pytest.mark.parametrize("tf32", [False, True])
def test_training_step_128_cube(tf32):
    """BASELINE.json configs[4]'s crop: one 128^3 crop, forward + backward, gradients against torch autograd, in the
    fp32 mode and in the default TF32 tensor-core mode (co-reference: torch with cudnn.allow_tf32 as in the reference)."""
    free, _ = torch.cuda.mem_get_info()
    if free < 30 << 30:
        pytest.skip("needs ~30 GB of free device memory")
    _step_vs_oracle(1, 128, 128, 128, 0.01, 2e-2 if tf32 else 5e-3, tf32=tf32)


def test_adam_steps_follow_torch_optimizer():
    """Three full steps (forward, backward, Adam over the flat bucket) against torch.optim.Adam driven by the oracle's
    gradients.  Adam divides by sqrt(v): an element whose gradient is noise-level (|g| << 1e-3 max|g|) can move by +-lr in
    either direction, so the bound is 2 * lr per step at worst and 5 % of that on average; the inference plan is rebuilt
    from the new weights."""
    from oracle import train_oracle as to
    tr, m, x, gt = _step_vs_oracle(1, 4, 32, 32, 0.02, 5e-3, tf32=False)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    names = [k for k, _ in m.named_parameters()]
    ref_params = [sd[k].clone().requires_grad_(True) for k in names]
    opt = torch.optim.Adam(ref_params, lr=1e-3)
    tr.zero_grad()
    tr.bucket.exp_avg.zero_(); tr.bucket.exp_avg_sq.zero_(); tr.bucket.step_count = 0
    for it in range(3):
        tr.zero_grad()
        tr.forward_backward(x, gt)
        tr.step(1e-3)
        sdo = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
        sdo.update({k: p.detach().double() for k, p in zip(names, ref_params)})
        _, grads, _ = to.training_step(x.double(), gt.double(), sdo, 0.02, param_names=[k for k in names if not k.startswith("proj")])
        for k in list(sd):
            if "running" in k:
                sd[k] = sdo[k].float()
        opt.zero_grad()
        for k, p in zip(names, ref_params):
            p.grad = grads[k].float() if k in grads else torch.zeros_like(p)
        opt.step()
    tot, cnt = 0.0, 0
    for k, p in zip(names, ref_params):
        diff = (dict(m.named_parameters())[k].data - p.detach()).abs()
        assert float(diff.max()) <= 3 * 2 * 1e-3 + 1e-6, k
        tot, cnt = tot + float(diff.sum()), cnt + diff.numel()
    assert tot / cnt <= 0.05 * 3 * 1e-3, tot / cnt
    m.eval()
    hm = m(x[0][None])[-1]["hm"]
    assert torch.isfinite(hm).all()
