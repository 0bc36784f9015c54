"""The tcgen05 implicit-GEMM convolution (csrc/conv_tc.cu) against torch fp32 convolutions on the
same bf16-rounded operands (floating point => torch fp32 reference; tolerance stated per test)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from cet_pick_b200 import _lib
    return _lib


def pack(w, nsrc, csrc, kc):
    """(Cout, nsrc*csrc, ntaps) -> [k-block][Cout][KC], k-block = (source, tap, chunk)."""
    cout, cin, ntaps = w.shape
    chunks = csrc // kc
    out = torch.empty((nsrc * ntaps * chunks, cout, kc), dtype=w.dtype, device=w.device)
    for s in range(nsrc):
        for t in range(ntaps):
            for ch in range(chunks):
                lo = s * csrc + ch * kc
                out[(s * ntaps + t) * chunks + ch] = w[:, lo:lo + kc, t]
    return out.contiguous()


def run_conv(L, srcs, wpk, kc, taps, ntot, bias, relu, epi, out, cstride=0, Ho=0, Wo=0, Cout=0):
    n, h, w, _ = srcs[0].shape
    tp = (C.c_int * (3 * len(taps)))(*[v for t in taps for v in t])
    rc = L.test_lib().cetpick_conv_bf16(len(srcs), srcs[0].data_ptr(), srcs[0].shape[3],
                                   srcs[1].data_ptr() if len(srcs) > 1 else None,
                                   srcs[1].shape[3] if len(srcs) > 1 else 0, n, h, w, wpk.data_ptr(), kc,
                                   len(taps), tp, ntot, bias.data_ptr() if bias is not None else None,
                                   int(relu), epi, out.data_ptr(), cstride, Ho, Wo, Cout, L.stream_ptr())
    L.check(rc, "cetpick_conv_bf16")
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", [(128, 16, 16), (256, 32, 32), (384, 64, 64), (256, 128, 128),
                                   (128, 256, 192), (640, 48, 96), (256, 32, 16)])
def test_gemm_selftest(L, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    kc = 64 if K % 64 == 0 else 32 if K % 32 == 0 else 16
    Bp = pack(B.view(N, K, 1), 1, K, kc)
    Cm = torch.full((M, N), float("nan"), device="cuda")
    L.check(L.test_lib().cetpick_selftest_gemm_bf16(A.data_ptr(), Bp.data_ptr(), Cm.data_ptr(), M, N, K,
                                               L.stream_ptr()), "selftest")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = (Cm - ref).abs().max().item()
    assert err <= 1e-3 * K ** 0.5 + 1e-4, err      # fp32 accumulation-order noise only


TAPS3 = [(0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 16, 16, 32), (3, 13, 21, 32, 32), (2, 19, 35, 64, 64),
                                            (1, 9, 17, 128, 128), (1, 5, 7, 256, 256), (2, 24, 40, 32, 64)])
def test_conv3x3_bias_relu(L, n, h, w, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
    b = torch.randn(cout, device="cuda", generator=g)
    kc = min(64, cin)
    wpk = pack(wt.reshape(cout, cin, 9), 1, cin, kc)
    out = torch.full((n, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    run_conv(L, [x], wpk, kc, TAPS3, cout, b, True, L.EPI_BF16_NHWC, out, cstride=cout)
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1)).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2, err      # one bf16 rounding of O(1) outputs (2^-8 relative) + accumulation order


def test_conv3x3_two_sources_is_concat(L):
    n, h, w, c = 2, 11, 18, 32
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16()
    s = torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(c, 2 * c, 3, 3, device="cuda", generator=g) / (3 * (2 * c) ** 0.5)).bfloat16()
    wpk = pack(wt.reshape(c, 2 * c, 9), 2, c, 32)
    out = torch.full((n, h, w, c), float("nan"), device="cuda", dtype=torch.bfloat16)
    run_conv(L, [a, s], wpk, 32, TAPS3, c, None, False, L.EPI_BF16_NHWC, out, cstride=c)
    xin = torch.cat((a, s), 3).float().permute(0, 3, 1, 2)
    ref = F.conv2d(xin, wt.float(), padding=1).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 2e-2


@pytest.mark.parametrize("cin,cout,h,w,Ho,Wo", [(64, 32, 6, 9, 12, 18), (256, 128, 5, 7, 9, 13), (128, 64, 4, 4, 7, 8)])
def test_upconv_2x2_scatter_with_autocrop(L, cin, cout, h, w, Ho, Wo):
    n = 2
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cin, cout, 2, 2, device="cuda", generator=g) / cin ** 0.5).bfloat16()
    b = torch.randn(cout, device="cuda", generator=g)
    kc = min(64, cin)
    # rows j = (dy*2+dx)*cout + co ; B[j][ci] = w[ci][co][dy][dx]
    Bm = wt.permute(2, 3, 1, 0).reshape(4 * cout, cin, 1).contiguous()
    wpk = pack(Bm, 1, cin, kc)
    out = torch.full((n, Ho, Wo, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    run_conv(L, [x], wpk, kc, [(0, 0, 0)], 4 * cout, b.repeat(4).contiguous(), True, L.EPI_UPCONV_2X2, out,
             Ho=Ho, Wo=Wo, Cout=cout)
    ref = F.relu(F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=2))[:, :, :Ho, :Wo]
    assert (out.float() - ref.permute(0, 2, 3, 1)).abs().max().item() <= 2e-2


def test_conv3d_dilated_27_taps(L):
    d, h, w, c = 5, 20, 27, 32
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(d, h, w, c, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(c, c, 3, 3, 3, device="cuda", generator=g) / (27 * c) ** 0.5).bfloat16()
    taps = [(kz - 1, 4 * (ky - 1), 4 * (kx - 1)) for kz in range(3) for ky in range(3) for kx in range(3)]
    wpk = pack(wt.reshape(c, c, 27), 1, c, 32)
    out = torch.full((d, h, w, c), float("nan"), device="cuda", dtype=torch.bfloat16)
    run_conv(L, [x], wpk, 32, taps, c, None, True, L.EPI_BF16_NHWC, out, cstride=c)
    xin = x.float().permute(3, 0, 1, 2)[None]
    ref = F.relu(F.conv3d(xin, wt.float(), padding=(1, 4, 4), dilation=(1, 4, 4)))[0].permute(1, 2, 3, 0)
    assert (out.float() - ref).abs().max().item() <= 2e-2


def test_proj_head_l2norm_ncdhw(L):
    d, h, w, c = 4, 9, 14, 32
    g = torch.Generator(device="cuda").manual_seed(13)
    x = torch.randn(d, h, w, c, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(32, c, 3, 1, 1, device="cuda", generator=g) / (3 * c) ** 0.5).bfloat16()
    taps = [(kz - 1, 0, 0) for kz in range(3)]
    wpk = pack(wt.reshape(32, c, 3), 1, c, 32)
    out = torch.full((32, d, h, w), float("nan"), device="cuda")
    run_conv(L, [x], wpk, 32, taps, 32, None, False, L.EPI_F32_L2NORM_NCDHW, out)
    ref = F.normalize(F.conv3d(x.float().permute(3, 0, 1, 2)[None], wt.float(), padding=(1, 0, 0)), dim=1)[0]
    assert (out - ref).abs().max().item() <= 1e-4


# ------------------------------------------------------------------------------------------------
# marching kernel (csrc/conv_march.cu): N-stacked taps along the march axis, resident weights
# ------------------------------------------------------------------------------------------------
def run_march(L, mode, dil, srcs, wt, cout, bias, relu):
    n, h, w, c = srcs[0].shape
    out = torch.full((n, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    wh = wt.float().cpu().contiguous()
    rc = L.test_lib().cetpick_conv_march_bf16(mode, dil, len(srcs), srcs[0].data_ptr(),
                                         srcs[1].data_ptr() if len(srcs) > 1 else None, c, n, h, w,
                                         wh.data_ptr(), cout, bias.data_ptr() if bias is not None else None,
                                         int(relu), out.data_ptr(), L.stream_ptr())
    L.check(rc, "cetpick_conv_march_bf16")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("n,h,w,cin,cout", [
    (1, 1, 5, 32, 32),          # a single row: every output row is first and last
    (2, 2, 16, 16, 32),
    (3, 13, 21, 32, 32),
    (2, 40, 130, 32, 32),       # two x blocks, ring of 16 slots wraps
    (2, 19, 35, 64, 64),        # ring of 8 slots
    (1, 37, 300, 32, 64),
    (2, 24, 40, 16, 32),
    (1, 70, 257, 64, 32),
])
def test_march2d_conv3x3_bias_relu(L, n, h, w, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(cin * 7 + cout + h)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
    b = torch.randn(cout, device="cuda", generator=g)
    out = run_march(L, L.MARCH_2D_ROWS, 1, [x], wt, cout, b, True)
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1)).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-2, err      # one bf16 rounding of O(1) outputs + accumulation order


@pytest.mark.parametrize("c", [32, 64])
def test_march2d_two_sources_is_concat(L, c):
    n, h, w = 2, 27, 140
    g = torch.Generator(device="cuda").manual_seed(5 + c)
    a = torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16()
    s = torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(c, 2 * c, 3, 3, device="cuda", generator=g) / (3 * (2 * c) ** 0.5)).bfloat16()
    out = run_march(L, L.MARCH_2D_ROWS, 1, [a, s], wt, c, None, False)
    xin = torch.cat((a, s), 3).float().permute(0, 3, 1, 2)
    ref = F.conv2d(xin, wt.float(), padding=1).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 2e-2


def test_march2d_many_strips_persistent(L):
    """more strips than SMs and several strips along y: ring state carries across strips"""
    n, h, w, c = 40, 96, 520, 32
    g = torch.Generator(device="cuda").manual_seed(99)
    x = torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(c, c, 3, 3, device="cuda", generator=g) / (3 * c ** 0.5)).bfloat16()
    out = run_march(L, L.MARCH_2D_ROWS, 1, [x], wt, c, None, True)
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1)).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 2e-2


@pytest.mark.parametrize("d,h,w", [(1, 9, 7), (5, 20, 27), (9, 33, 50), (40, 64, 64)])
def test_march3d_dilated_27_taps(L, d, h, w):
    c = 32
    g = torch.Generator(device="cuda").manual_seed(11 + d)
    x = torch.randn(d, h, w, c, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(c, c, 3, 3, 3, device="cuda", generator=g) / (27 * c) ** 0.5).bfloat16()
    out = run_march(L, L.MARCH_3D_PLANES, 4, [x], wt, c, None, True)
    xin = x.float().permute(3, 0, 1, 2)[None]
    ref = F.relu(F.conv3d(xin, wt.float(), padding=(1, 4, 4), dilation=(1, 4, 4)))[0].permute(1, 2, 3, 0)
    assert (out.float() - ref).abs().max().item() <= 2e-2


@pytest.mark.parametrize("cin,cout,n,h,w,Ho,Wo", [(64, 32, 2, 6, 9, 12, 18), (256, 128, 2, 5, 7, 9, 13),
                                                  (128, 64, 3, 4, 4, 7, 8), (64, 32, 5, 33, 47, 66, 93),
                                                  (128, 64, 2, 40, 40, 80, 80), (256, 128, 1, 16, 24, 32, 48)])
def test_conv_up_kernel(L, cin, cout, n, h, w, Ho, Wo):
    """csrc/conv_up.cu: resident weights, flat pixel tiles, pixel-shuffle epilogue with autocrop"""
    g = torch.Generator(device="cuda").manual_seed(cin + cout + h)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cin, cout, 2, 2, device="cuda", generator=g) / cin ** 0.5).bfloat16()
    b = torch.randn(cout, device="cuda", generator=g)
    out = torch.full((n, Ho, Wo, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    wh, bh = wt.float().cpu().contiguous(), b.cpu().contiguous()
    L.check(L.test_lib().cetpick_upconv_bf16(x.data_ptr(), cin, n, h, w, wh.data_ptr(), bh.data_ptr(), cout,
                                        out.data_ptr(), Ho, Wo, L.stream_ptr()), "cetpick_upconv_bf16")
    torch.cuda.synchronize()
    ref = F.relu(F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt.float(), b, stride=2))[:, :, :Ho, :Wo]
    assert (out.float() - ref.permute(0, 2, 3, 1)).abs().max().item() <= 2e-2


@pytest.mark.parametrize("nsrc,c,cout,n,h,w", [(1, 64, 128, 2, 16, 16), (1, 128, 128, 2, 19, 35), (1, 128, 256, 1, 9, 40),
                                               (1, 256, 256, 3, 33, 17), (2, 128, 128, 2, 27, 30), (1, 64, 128, 40, 64, 64)])
def test_conv_halo_kernel(L, nsrc, c, cout, n, h, w):
    """csrc/conv_halo.cu: one halo tile per channel chunk, nine shifted descriptors, streamed weights"""
    g = torch.Generator(device="cuda").manual_seed(c + cout + h)
    srcs = [torch.randn(n, h, w, c, device="cuda", generator=g).bfloat16() for _ in range(nsrc)]
    cin = nsrc * c
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
    b = torch.randn(cout, device="cuda", generator=g)
    out = torch.full((n, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    wh, bh = wt.float().cpu().contiguous(), b.cpu().contiguous()
    L.check(L.test_lib().cetpick_conv_halo_bf16(nsrc, srcs[0].data_ptr(), srcs[1].data_ptr() if nsrc > 1 else None, c,
                                           n, h, w, wh.data_ptr(), bh.data_ptr(), cout, 1, out.data_ptr(),
                                           L.stream_ptr()), "cetpick_conv_halo_bf16")
    torch.cuda.synchronize()
    xin = torch.cat(srcs, 3).float().permute(0, 3, 1, 2)
    ref = F.relu(F.conv2d(xin, wt.float(), b, padding=1)).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 2e-2


# ---- tensor-core stem (csrc/conv_stem.cu): Conv2d(1,16,7,s2,p3)+BN+ReLU, unet_small.py:35-37,72-74 ----
@pytest.mark.parametrize("d,h,w", [(2, 32, 64), (3, 37, 68), (1, 5, 8), (2, 300, 520), (1, 129, 1028), (2, 2, 4)])
def test_stem_tensor_core_march(L, d, h, w):
    g = torch.Generator(device="cuda").manual_seed(d * 100 + h + w)
    x = torch.rand(d, h, w, device="cuda", generator=g)
    wt = (torch.rand(16, 1, 7, 7, device="cuda", generator=g) * 2 - 1) / 7
    scale = torch.rand(16, device="cuda", generator=g) + 0.5
    shift = torch.randn(16, device="cuda", generator=g) * 0.1
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = torch.full((d, ho, wo, 16), float("nan"), device="cuda", dtype=torch.bfloat16)
    wh, sh, bh = wt.cpu().contiguous(), scale.cpu().contiguous(), shift.cpu().contiguous()
    rc = L.test_lib().cetpick_conv_stem_bf16(x.data_ptr(), d, h, w, wh.data_ptr(), sh.data_ptr(), bh.data_ptr(),
                                        out.data_ptr(), L.stream_ptr())
    L.check(rc, "cetpick_conv_stem_bf16")
    torch.cuda.synchronize()
    # reference on the same bf16-rounded operands (input and BN-folded weight), fp32 accumulate
    wf = (wt * scale.view(16, 1, 1, 1)).bfloat16().float()
    ref = F.relu(F.conv2d(x.bfloat16().float()[:, None], wf, shift, stride=2, padding=3)).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    err = (out.float() - ref).abs().max().item()
    assert err <= 1.5e-2, err      # one bf16 rounding of O(1) outputs + accumulation order
    # and against the exact fp32 conv: operand rounding adds ~2^-9 relative per product
    ref32 = F.relu(F.conv2d(x[:, None], wt * scale.view(16, 1, 1, 1), shift, stride=2, padding=3)).permute(0, 2, 3, 1)
    assert (out.float() - ref32).abs().max().item() <= 3e-2


# ---- MaxPool2d(2, ceil_mode=True) fused into the marching conv's epilogue (unet.py:225,237-238) ----
@pytest.mark.parametrize("n,h,w,cin,cout", [
    (2, 12, 40, 16, 32),        # MT = 1 (one M-tile), even sizes
    (1, 13, 21, 32, 32),        # odd H and W: ceil mode, last row / column pooled alone
    (2, 38, 300, 32, 32),       # MT = 2, several strips, x block with a ragged tail
    (2, 19, 35, 64, 64),        # Cout 64: row pairs alternate between the two epilogue groups
    (1, 64, 257, 32, 64),
    (1, 1, 7, 32, 32),          # a single row
])
def test_march_fused_maxpool(L, n, h, w, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(n + h + w + cin)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / (3 * cin ** 0.5)).bfloat16()
    b = torch.randn(cout, device="cuda", generator=g) * 0.2
    out = torch.full((n, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    hp, wp = (h + 1) // 2, (w + 1) // 2
    pool = torch.full((n, hp, wp, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
    wh = wt.float().cpu().contiguous()
    rc = L.test_lib().cetpick_conv_march_pool_bf16(0, 1, 1, x.data_ptr(), None, cin, n, h, w, wh.data_ptr(), cout,
                                              b.data_ptr(), 1, out.data_ptr(), pool.data_ptr(), L.stream_ptr())
    L.check(rc, "cetpick_conv_march_pool_bf16")
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), b, padding=1)).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 2e-2
    # the pooled tensor is EXACTLY the ceil-mode max-pool of the bf16 output the kernel wrote
    pref = F.max_pool2d(out.float().permute(0, 3, 1, 2), 2, ceil_mode=True).permute(0, 2, 3, 1)
    assert torch.equal(pool.float(), pref)


# ---- fused full-resolution block (csrc/conv_block.cu): conv+ReLU -> conv+ReLU (-> pool), intermediate map in smem ----
@pytest.mark.parametrize("n,h,w,c1,nsrc,pool", [
    (2, 20, 130, 16, 1, True),       # cluster of 2 CTAs, second one holds 2 valid pixels
    (3, 33, 256, 16, 1, True),       # exactly two M-tiles, odd height (ceil-mode pool)
    (2, 41, 300, 32, 2, False),      # concat of two sources (UpConv), cluster of 3
    (1, 64, 512, 32, 2, False),      # cluster of 4: the configs[1] row width at level 0
    (1, 17, 700, 32, 1, False),      # cluster of 6
    (40, 96, 384, 16, 1, True),      # more strips than resident clusters: ring / slot state carries across strips
])
def test_fused_block_equals_two_march_convs(L, n, h, w, c1, nsrc, pool):
    """The fused kernel must reproduce, bit for bit, conv1 -> bf16 -> conv2 run as two marching kernels (same MMA
    shapes, same accumulation order), and agree with torch fp32 on the same bf16-rounded operands."""
    g = torch.Generator(device="cuda").manual_seed(n * 3 + h + w + c1)
    srcs = [torch.randn(n, h, w, c1, device="cuda", generator=g).bfloat16() for _ in range(nsrc)]
    w1 = (torch.randn(32, nsrc * c1, 3, 3, device="cuda", generator=g) / (3 * (nsrc * c1) ** 0.5)).bfloat16()
    w2 = (torch.randn(32, 32, 3, 3, device="cuda", generator=g) / (3 * 32 ** 0.5)).bfloat16()
    b1 = torch.randn(32, device="cuda", generator=g) * 0.2
    b2 = torch.randn(32, device="cuda", generator=g) * 0.2
    out = torch.full((n, h, w, 32), float("nan"), device="cuda", dtype=torch.bfloat16)
    hp, wp = (h + 1) // 2, (w + 1) // 2
    pl = torch.full((n, hp, wp, 32), float("nan"), device="cuda", dtype=torch.bfloat16) if pool else None
    w1h, w2h = w1.float().cpu().contiguous(), w2.float().cpu().contiguous()
    b1h, b2h = b1.cpu().contiguous(), b2.cpu().contiguous()
    rc = L.test_lib().cetpick_conv_block_bf16(nsrc, srcs[0].data_ptr(), srcs[1].data_ptr() if nsrc > 1 else None, c1, n, h, w,
                                         w1h.data_ptr(), b1h.data_ptr(), w2h.data_ptr(), b2h.data_ptr(),
                                         out.data_ptr(), pl.data_ptr() if pool else None, L.stream_ptr())
    L.check(rc, "cetpick_conv_block_bf16")
    torch.cuda.synchronize()
    mid = run_march(L, L.MARCH_2D_ROWS, 1, srcs, w1, 32, b1, True)
    two = run_march(L, L.MARCH_2D_ROWS, 1, [mid], w2, 32, b2, True)
    assert not torch.isnan(out.float()).any()
    ndiff = int((out.view(torch.int16) != two.view(torch.int16)).sum())
    assert ndiff == 0, f"{ndiff} of {out.numel()} values differ from the two-kernel result (max {float((out.float() - two.float()).abs().max()):.3e})"
    xin = torch.cat(srcs, 3).float().permute(0, 3, 1, 2)
    r1 = F.relu(F.conv2d(xin, w1.float(), b1, padding=1)).bfloat16().float()
    ref = F.relu(F.conv2d(r1, w2.float(), b2, padding=1)).permute(0, 2, 3, 1)
    assert (out.float() - ref).abs().max().item() <= 4e-2
    if pool:
        pref = F.max_pool2d(out.float().permute(0, 3, 1, 2), 2, ceil_mode=True).permute(0, 2, 3, 1)
        assert torch.equal(pl.float(), pref)
