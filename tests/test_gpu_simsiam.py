"""Exploration-step embedding network (csrc/simsiam.cu + csrc/conv_small.cu): TomoResClassifier.forward_test of
cet_pick/models/networks/simsiam_model.py:325-366 against the reference-generated fixture `simsiam3d_small` and the
fp32 oracle; the small-map convolution kernel against torch fp32 on the same bf16-rounded operands."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synthdata as synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from cet_pick_b200 import _lib
    return _lib


def run_small(L, x, w, stride, Ho, Wo, taps, bias=None, residual=None, relu=False, out_f32=False):
    """x: (B, Z, Hin, Win, C) bf16; w: (Cout, C, ntaps) fp32/bf16 values"""
    B, Z, Hin, Win, Cc = x.shape
    Cout = w.shape[0]
    out = torch.full((B, Z, Ho, Wo, Cout), float("nan"), device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    wh = w.float().cpu().contiguous()
    tp = (C.c_int * (3 * len(taps)))(*[v for t in taps for v in t])
    rc = L.test_lib().cetpick_conv_small_bf16(x.data_ptr(), Cc, B, Z, Hin, Win, stride, Ho, Wo, wh.data_ptr(), Cout, len(taps), tp,
                                         bias.data_ptr() if bias is not None else None,
                                         residual.data_ptr() if residual is not None else None, int(relu), int(out_f32),
                                         out.data_ptr(), L.stream_ptr())
    L.check(rc, "cetpick_conv_small_bf16")
    torch.cuda.synchronize()
    return out


TAPS9 = [(0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]


@pytest.mark.parametrize("hw,cin,cout,n", [(8, 64, 64, 37), (4, 128, 128, 50), (2, 256, 256, 70), (8, 64, 128, 5)])
def test_small_conv3x3_residual_relu(L, hw, cin, cout, n):
    g = torch.Generator(device="cuda").manual_seed(hw + cin + n)
    x = torch.randn(1, n, hw, hw, cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
    b = torch.randn(cout, device="cuda", generator=g) * 0.2
    res = torch.randn(1, n, hw, hw, cout, device="cuda", generator=g).bfloat16()
    out = run_small(L, x, w.reshape(cout, cin, 9), 1, hw, hw, TAPS9, b, res, True)
    ref = F.conv2d(x[0].float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1) + res[0].float()
    ref = F.relu(ref)
    assert (out[0].float() - ref).abs().max().item() <= 3e-2


@pytest.mark.parametrize("hin,cin,cout,n", [(8, 64, 128, 21), (4, 128, 256, 64)])
def test_small_conv_stride2(L, hin, cin, cout, n):
    """the convolution stride is the traversal stride of the TMA tensor map (elementStrides)"""
    g = torch.Generator(device="cuda").manual_seed(hin + cin)
    x = torch.randn(1, n, hin, hin, cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
    ho = hin // 2
    out = run_small(L, x, w.reshape(cout, cin, 9), 2, ho, ho, TAPS9, None, None, True)
    ref = F.relu(F.conv2d(x[0].float().permute(0, 3, 1, 2), w.float(), None, stride=2, padding=1)).permute(0, 2, 3, 1)
    assert (out[0].float() - ref).abs().max().item() <= 3e-2
    # 1x1 stride-2 shortcut (no BN, no ReLU)
    w1 = (torch.randn(cout, cin, 1, 1, device="cuda", generator=g) / cin ** 0.5).bfloat16()
    out1 = run_small(L, x, w1.reshape(cout, cin, 1), 2, ho, ho, [(0, 0, 0)])
    ref1 = F.conv2d(x[0].float().permute(0, 3, 1, 2), w1.float(), None, stride=2).permute(0, 2, 3, 1)
    assert (out1[0].float() - ref1).abs().max().item() <= 3e-2


def test_small_conv3d_and_linear(L):
    g = torch.Generator(device="cuda").manual_seed(9)
    B, D, c = 5, 32, 256
    x = torch.randn(B, D, 2, 2, c, device="cuda", generator=g).bfloat16()
    w = (torch.randn(c, c, 3, 3, 3, device="cuda", generator=g) / (27 * c) ** 0.5).bfloat16()
    taps = [(kz - 1, ky - 1, kx - 1) for kz in range(3) for ky in range(3) for kx in range(3)]
    out = run_small(L, x, w.reshape(c, c, 27), 1, 2, 2, taps, None, None, True)
    ref = F.relu(F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.float(), padding=1)).permute(0, 2, 3, 4, 1)
    assert (out.float() - ref).abs().max().item() <= 3e-2
    # Linear over a batch of 300 rows (ragged last tile), fp32 out
    v = torch.randn(1, 300, 1, 1, c, device="cuda", generator=g).bfloat16()
    wl = (torch.randn(c, c, device="cuda", generator=g) / c ** 0.5).bfloat16()
    bl = torch.randn(c, device="cuda", generator=g)
    o = run_small(L, v, wl.reshape(c, c, 1), 1, 1, 1, [(0, 0, 0)], bl, None, False, True)
    refl = v.view(300, c).float() @ wl.float().t() + bl
    assert (o.view(300, c) - refl).abs().max().item() <= 2e-3


def build_model(seed_w=5):
    from cet_pick_b200.models.model import create_model
    m = create_model("simsiam3d_18", {"proj": 256, "pred": 256}, 0)
    m.load_state_dict(synth.simsiam3d_state_dict_torch(seed_w))
    return m.cuda().eval()


def rel_err(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-12))


@pytest.mark.parametrize("name", ["simsiam3d_small", "simsiam3d_d1", "simsiam3d_d5"])
def test_forward_test_vs_reference_golden(golden, name):
    g = golden(name)
    m = build_model(int(g["seed_w"]))
    D, H, W = [int(v) for v in g["shape"]]
    x = torch.from_numpy(np.stack([synth.tomogram_np(D, H, W, int(s)) for s in g["seeds"]])).cuda()
    out = m.forward_test(x)
    torch.cuda.synchronize()
    for k in ("proj", "pred"):
        got, ref = out[k].cpu().numpy(), g[k]
        e_abs, e_rel = float(np.abs(got - ref).max()), rel_err(got, ref)
        print(f"simsiam3d {k}: max-abs err {e_abs:.3e} (|ref| max {np.abs(ref).max():.3f}), relative L2 err {e_rel:.3e}")
        assert got.shape == ref.shape
        assert e_rel <= 1e-2 and e_abs <= 1e-2      # BF16 operands through 20 layers (measured 3e-3 / 2.6e-3)


def test_forward_test_vs_oracle_batch():
    """a batch that spans several tiles in every layer (and a 5-D input like the dataset's (B,1,D,H,W) tensors)"""
    from oracle import simsiam_oracle as so
    B = 19
    sd = synth.simsiam3d_state_dict_torch(5)
    m = build_model(5)
    x = torch.from_numpy(np.stack([synth.tomogram_np(32, 32, 32, 100 + s) for s in range(B)]))
    with torch.no_grad():
        ref = so.forward_test(x, sd)
    out = m.forward_test(x[:, None].cuda())
    for k in ("proj", "pred"):
        got, r = out[k].cpu().numpy(), ref[k].numpy()
        print(f"simsiam3d batch {k}: max-abs {np.abs(got - r).max():.3e}, relative L2 {rel_err(got, r):.3e}")
        assert rel_err(got, r) <= 1e-2
    # embeddings are what the exploration step clusters: nearest neighbours by cosine similarity must agree
    a, b = out["proj"].cpu().numpy(), ref["proj"].numpy()
    cos = lambda v: (v / np.linalg.norm(v, axis=1, keepdims=True)) @ (v / np.linalg.norm(v, axis=1, keepdims=True)).T
    assert np.abs(cos(a) - cos(b)).max() <= 2e-2


# ---------------------------------------------------------------------------------------------------------------
# 2-D exploration variant (simsiam_model_2d.py:617-774, arch simsiam2d_18): maps of 32x32 / 16x16 / 8x8 pixels; the
# small-map kernel takes the larger maps in row bands of 128 pixels
@pytest.mark.parametrize("hw,cin,cout,n", [(32, 64, 64, 5), (16, 128, 128, 9), (64, 64, 64, 2), (16, 64, 128, 3)])
def test_small_conv3x3_row_bands(L, hw, cin, cout, n):
    g = torch.Generator(device="cuda").manual_seed(hw + cin + n)
    x = torch.randn(1, n, hw, hw, cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
    b = torch.randn(cout, device="cuda", generator=g) * 0.2
    res = torch.randn(1, n, hw, hw, cout, device="cuda", generator=g).bfloat16()
    out = run_small(L, x, w.reshape(cout, cin, 9), 1, hw, hw, TAPS9, b, res, True)
    ref = F.conv2d(x[0].float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1) + res[0].float()
    ref = F.relu(ref)
    assert (out[0].float() - ref).abs().max().item() <= 3e-2


@pytest.mark.parametrize("hin,cin,cout,n", [(32, 64, 128, 7), (16, 128, 256, 11), (64, 64, 128, 2)])
def test_small_conv_stride2_row_bands(L, hin, cin, cout, n):
    g = torch.Generator(device="cuda").manual_seed(hin + cin)
    x = torch.randn(1, n, hin, hin, cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
    ho = hin // 2
    out = run_small(L, x, w.reshape(cout, cin, 9), 2, ho, ho, TAPS9, None, None, True)
    ref = F.relu(F.conv2d(x[0].float().permute(0, 3, 1, 2), w.float(), None, stride=2, padding=1)).permute(0, 2, 3, 1)
    assert (out[0].float() - ref).abs().max().item() <= 3e-2
    w1 = (torch.randn(cout, cin, 1, 1, device="cuda", generator=g) / cin ** 0.5).bfloat16()
    out1 = run_small(L, x, w1.reshape(cout, cin, 1), 2, ho, ho, [(0, 0, 0)])
    ref1 = F.conv2d(x[0].float().permute(0, 3, 1, 2), w1.float(), None, stride=2).permute(0, 2, 3, 1)
    assert (out1[0].float() - ref1).abs().max().item() <= 3e-2


def build_model_2d(seed_w=6, out_dim=128):
    from cet_pick_b200.models.model import create_model
    m = create_model("simsiam2d_18", {"proj": out_dim, "pred": out_dim}, out_dim)
    m.load_state_dict(synth.simsiam2d_state_dict_torch(seed_w, out_dim=out_dim))
    return m.cuda().eval()


@pytest.mark.parametrize("name", ["simsiam2d_small", "simsiam2d_hw16", "simsiam2d_hc32"])
def test_forward_test_2d_vs_reference_golden(golden, name):
    g = golden(name)
    hw, od = int(g["hw"]), int(g["out_dim"])
    m = build_model_2d(int(g["seed_w"]), od)
    x = torch.from_numpy(np.stack([synth.tomogram_np(1, hw, hw, int(s)) for s in g["seeds"]])).cuda()
    out = m.forward_test(x)
    torch.cuda.synchronize()
    for k in ("proj", "pred"):
        got, ref = out[k].cpu().numpy(), g[k]
        e_abs, e_rel = float(np.abs(got - ref).max()), rel_err(got, ref)
        print(f"{name} {k}: max-abs err {e_abs:.3e} (|ref| max {np.abs(ref).max():.3f}), relative L2 err {e_rel:.3e}")
        assert got.shape == ref.shape
        assert e_rel <= 1e-2 and e_abs <= 2e-2      # BF16 operands through 17 layers


def test_forward_test_2d_vs_oracle_batch():
    """a batch that spans many tiles in every layer, given as the 5-D tensor a dataset would hand over"""
    from oracle import simsiam_oracle as so
    B = 37
    sd = synth.simsiam2d_state_dict_torch(6, out_dim=128)
    m = build_model_2d(6, 128)
    x = torch.from_numpy(np.stack([synth.tomogram_np(1, 32, 32, 200 + s) for s in range(B)]))
    with torch.no_grad():
        ref = so.forward_test_2d(x, sd)
    out = m.forward_test(x[:, None].cuda())
    for k in ("proj", "pred"):
        got, r = out[k].cpu().numpy(), ref[k].numpy()
        print(f"simsiam2d batch {k}: max-abs {np.abs(got - r).max():.3e}, relative L2 {rel_err(got, r):.3e}")
        assert rel_err(got, r) <= 1e-2
    with pytest.raises(RuntimeError):
        m.forward_test(torch.zeros(2, 3, 32, 32, device="cuda"))        # conv1 has ONE input channel
