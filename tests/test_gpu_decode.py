"""Parity of the CUDA decode (csrc/decode.cu, through the C-ABI) against the oracle and the
reference-generated golden fixtures.  Bit-exact: same rows, same order, same float bits."""
import numpy as np
import pytest
import torch

import synthdata as synth

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def dec():
    from cet_pick_b200.models import decode
    return decode


@pytest.fixture(scope="module")
def do():
    from oracle import decode_oracle
    return decode_oracle


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def hm_of(g):
    D, H, W = [int(v) for v in g["shape"]]
    return synth.heatmap_tiefree_np(D, H, W, int(g["seed"]))[None, None]


@pytest.mark.parametrize("name", ["decode_tiefree_k3", "decode_tiefree_k5", "decode_tiefree_k1", "decode_fiber_k3",
                                  "decode_fiber_k5", "decode_fiber_k7"])
def test_golden_small(golden, dec, name):
    g = golden(name)
    out = dec.tomo_decode(cu(hm_of(g)), kernel=int(g["kernel"]), K=int(g["K"]), if_fiber=bool(g["fiber"]))
    assert np.array_equal(bits(out.cpu().numpy()), bits(g["dets"]))


@pytest.mark.parametrize("name", ["decode_quirk_20x1024x1024", "decode_quirk_70x500x500"])
def test_golden_fp32_index_quirk(golden, dec, name):
    """> 2^24 voxels: the sampled-threshold path and the reference's fp32 index arithmetic."""
    g = golden(name)
    hm = hm_of(g).copy()
    for (z, y, x), v in zip(g["plant"], g["plant_vals"]):
        hm[0, 0, z, y, x] = v
    out = dec.tomo_decode(cu(hm), kernel=3, K=int(g["K"]))
    assert np.array_equal(bits(out.cpu().numpy()), bits(g["dets"]))
    flags, ncand = dec.decode_status()
    assert flags == 0 and ncand >= int(g["K"])


def test_golden_reg_batch2(golden, dec):
    g = golden("decode_reg_b2")
    D, H, W = [int(v) for v in g["shape"]]
    hm = np.stack([synth.heatmap_tiefree_np(D, H, W, int(s)) for s in g["seeds"]])[:, None]
    reg = (synth.uniform_np(int(g["reg_seed"]), 2 * 2 * D * H * W).reshape(2, 2, D, H, W) - 0.5).astype(np.float32)
    out = dec.tomo_decode(cu(hm), kernel=3, reg=cu(reg), K=int(g["K"]))
    assert np.array_equal(bits(out.cpu().numpy()), bits(g["dets"]))


def test_golden_parts(golden, dec):
    g = golden("decode_parts")
    hm = cu(hm_of(g))
    assert np.array_equal(bits(dec._nms(hm, 3).cpu().numpy()), bits(g["nms"]))
    assert np.array_equal(bits(dec._nms_xy(hm, 3).cpu().numpy()), bits(g["nms_xy"]))
    assert np.array_equal(bits(dec._nms_z(hm, 3).cpu().numpy()), bits(g["nms_z"]))
    ts, zs, ys, xs, ti = dec._topk(dec._nms(hm, 3), K=25)
    assert np.array_equal(bits(ts.cpu().numpy()), bits(g["topk_scores"]))
    assert np.array_equal(ti.cpu().numpy(), g["topk_inds"])
    assert np.array_equal(zs.cpu().numpy(), g["topk_zs"]) and np.array_equal(xs.cpu().numpy(), g["topk_xs"])
    assert np.array_equal(bits(ys.cpu().numpy()), bits(g["topk_ys"]))


def test_golden_sigmoid(golden):
    from cet_pick_b200.models.utils import _sigmoid
    g = golden("sigmoid")
    x = cu(g["x"].copy())
    y = _sigmoid(x)
    assert y.data_ptr() == x.data_ptr()                        # in place, like the reference
    assert np.abs(y.cpu().numpy() - g["y"]).max() <= 2e-7      # expf vs Sleef: <= 2 ulp at 1.0
    assert y.min().item() == np.float32(1e-4) and y.max().item() == np.float32(1 - 1e-4)


def test_plateau_matches_oracle_bit_exact(golden, dec, do):
    """K > number of real peaks: filler rows come from the clamp-floor plateau; the oracle and the
    kernel agree on (score desc, index asc), and on every row above the floor with the reference."""
    g = golden("decode_plateau")
    D, H, W = [int(v) for v in g["shape"]]
    hm = synth.heatmap_peaks_np(D, H, W, int(g["n_peaks"]), int(g["seed"]))[None, None]
    out = dec.tomo_decode(cu(hm), kernel=3, K=int(g["K"])).cpu().numpy()
    assert np.array_equal(bits(out), bits(do.tomo_decode(hm, 3, None, int(g["K"]))))
    n = int((g["dets"][0, :, 3] > np.float32(1e-4)).sum())
    key = lambda a: a[np.lexsort((a[:, 0], a[:, 1], a[:, 2], -a[:, 3]))]
    assert np.array_equal(bits(key(out[0, :n])), bits(key(g["dets"][0, :n])))


CASES = [
    # (D, H, W, K, kernel, fiber)   ragged / unaligned / tiny / K == N
    (1, 1, 1, 1, 3, False),
    (3, 5, 7, 105, 3, False),
    (7, 33, 129, 64, 3, False),
    (5, 40, 130, 50, 5, False),
    (4, 31, 37, 40, 7, False),
    (9, 64, 256, 300, 3, True),
    (2, 9, 515, 33, 1, False),
    (70, 37, 45, 500, 3, False),
]


@pytest.mark.parametrize("D,H,W,K,kernel,fiber", CASES)
def test_oracle_shapes(dec, do, D, H, W, K, kernel, fiber):
    hm = synth.heatmap_tiefree_np(D, H, W, D * 1000 + W)[None, None]
    out = dec.tomo_decode(cu(hm), kernel=kernel, K=K, if_fiber=fiber).cpu().numpy()
    assert np.array_equal(bits(out), bits(do.tomo_decode(hm, kernel, None, K, fiber)))


def test_oracle_signed_values_and_zeros(dec, do):
    """negative heat, -0.0 products and exact zeros: heat*keep keeps the sign of heat."""
    D, H, W = 6, 20, 36
    u = synth.uniform_np(5, D * H * W).reshape(1, 1, D, H, W)
    hm = (u - np.float32(0.6)).astype(np.float32)
    hm[0, 0, 2, 3:9, 4:30] = 0.0
    K = D * H * W
    out = dec.tomo_decode(cu(hm), kernel=3, K=K).cpu().numpy()
    assert np.array_equal(bits(out), bits(do.tomo_decode(hm, 3, None, K)))


def test_oracle_medium_sampled_path(dec, do):
    """4.2 Mvoxel (> candidate capacity): exercises sample select + COLLECT + final select."""
    D, H, W = 64, 256, 256
    hm = synth.heatmap_tiefree_np(D, H, W, 21)[None, None]
    out = dec.tomo_decode(cu(hm), kernel=3, K=1000).cpu().numpy()
    assert np.array_equal(bits(out), bits(do.tomo_decode(hm, 3, None, 1000)))
    flags, ncand = dec.decode_status()
    assert flags == 0 and 1000 <= ncand < 2_000_000


def test_oracle_medium_plateau_eq_path(dec, do):
    """big floor plateau + K larger than the number of peaks: EQ pass fills from plane 0 upwards."""
    D, H, W = 48, 256, 256
    hm = synth.heatmap_peaks_np(D, H, W, 200, seed=3)[None, None]
    out = dec.tomo_decode(cu(hm), kernel=3, K=2000).cpu().numpy()
    assert np.array_equal(bits(out), bits(do.tomo_decode(hm, 3, None, 2000)))


# ---- sieve_kernel (threshold-first COLLECT; aligned rows, volumes above the candidate capacity) ----
@pytest.mark.parametrize("kernel,fiber,K", [(5, False, 800), (7, False, 300), (3, True, 1000), (1, False, 500),
                                            (5, True, 400), (7, True, 200)])
def test_sieve_window_modes(dec, do, kernel, fiber, K):
    D, H, W = 40, 256, 256
    hm = synth.heatmap_tiefree_np(D, H, W, 100 + kernel)[None, None]
    out = dec.tomo_decode(cu(hm), kernel=kernel, K=K, if_fiber=fiber).cpu().numpy()
    assert np.array_equal(bits(out), bits(do.tomo_decode(hm, kernel, None, K, fiber)))
    flags, _ = dec.decode_status()
    assert flags & 1 == 0 and (flags == 0 or (fiber and kernel != 3))   # wide fiber windows: top-K of a mostly-zero map


def test_sieve_plain_topk(dec, do):
    """no NMS (P = 0, no z neighbourhood): every voxel is its own window maximum."""
    D, H, W = 40, 256, 256
    hm = synth.heatmap_tiefree_np(D, H, W, 9)[None, None]
    ts, zs, ys, xs, ti = dec._topk(cu(hm), K=1500)
    rs, rz, ry, rx, ri = do.topk(hm, 1500)
    assert np.array_equal(ti.cpu().numpy(), ri) and np.array_equal(bits(ts.cpu().numpy()[:, 0]), bits(rs))


def test_sieve_degenerate_threshold_negative_map(dec, do):
    """all-negative map: the K largest NMS outputs are the -0.0 products of suppressed voxels, the
    threshold is 0 and every voxel is a hit (the exact slow path of the sieve)."""
    D, H, W = 36, 256, 256
    u = synth.uniform_np(11, D * H * W).reshape(1, 1, D, H, W)
    hm = (u - np.float32(2.0)).astype(np.float32)
    out = dec.tomo_decode(cu(hm), kernel=3, K=777).cpu().numpy()
    assert np.array_equal(bits(out), bits(do.tomo_decode(hm, 3, None, 777)))


def test_sieve_dense_plateau_at_threshold(dec, do):
    """the K-th score sits on a huge plateau (map clipped from above): hits are dense, picks come
    from the plateau in index order through the EQ pass."""
    D, H, W = 36, 256, 256
    hm = np.minimum(synth.heatmap_tiefree_np(D, H, W, 5), np.float32(0.75))[None, None]
    out = dec.tomo_decode(cu(hm), kernel=3, K=900).cpu().numpy()
    assert np.array_equal(bits(out), bits(do.tomo_decode(hm, 3, None, 900)))


def test_sieve_nan_flag(dec):
    D, H, W = 36, 256, 256
    hm = synth.heatmap_tiefree_np(D, H, W, 6).copy()
    hm[17, 100, 31] = np.nan
    dec.tomo_decode(cu(hm[None, None]), kernel=3, K=100)
    flags, _ = dec.decode_status()
    assert flags & 1


def test_topk_plain_full_ranking(dec, do):
    """_topk has no NMS: the full descending ranking (K == N) of a random map must match."""
    D, H, W = 9, 24, 40
    hm = synth.heatmap_tiefree_np(D, H, W, 44)[None, None]
    K = D * H * W
    ts, zs, ys, xs, ti = dec._topk(cu(hm), K=K)
    rs, rz, ry, rx, ri = do.topk(hm, K)
    assert np.array_equal(ti.cpu().numpy(), ri)
    assert np.array_equal(bits(ts.cpu().numpy()[:, 0]), bits(rs))
    assert np.array_equal(zs.cpu().numpy(), rz) and np.array_equal(xs.cpu().numpy(), rx)


def test_fallback_exact_select(dec, do):
    """adversarial map for the sampled bound: plain top-K (every voxel is a candidate) with a
    low-valued sample region -> candidate overflow -> exact full-volume select (flag bit1)."""
    D, H, W = 40, 256, 256
    hm = synth.heatmap_tiefree_np(D, H, W, 33).copy()
    hm[D // 2 - 2:D // 2 + 3] *= np.float32(0.5)
    hm = hm[None, None]
    ts, zs, ys, xs, ti = dec._topk(cu(hm), K=700)
    rs, rz, ry, rx, ri = do.topk(hm, 700)
    assert np.array_equal(ti.cpu().numpy(), ri) and np.array_equal(bits(ts.cpu().numpy()[:, 0]), bits(rs))
    flags, _ = dec.decode_status()
    assert flags & 2, dec.decode_debug_state()


def test_bad_arguments(dec):
    hm = torch.rand(1, 1, 4, 8, 8, device="cuda")
    with pytest.raises(ValueError):
        dec.tomo_decode(hm, kernel=4, K=10)           # even kernel: the reference fails too
    with pytest.raises(ValueError):
        dec.tomo_decode(hm, kernel=3, K=4 * 8 * 8 + 1)  # K > N: torch.topk raises
    with pytest.raises(ValueError):
        dec.tomo_decode(hm, kernel=4, K=10, if_fiber=True)
    with pytest.raises(RuntimeError):
        dec.tomo_decode(hm.cpu(), kernel=3, K=10)     # no CPU fallback


def test_large_vs_torch_cuda_reference(dec):
    """67 Mvoxel tie-free map generated on the device: the reference's own op sequence
    (max_pool3d, ==, *, topk; decode.py:27-33,84) run by PyTorch on the same GPU must agree
    exactly; plus size-independent properties."""
    D, H, W, K = 256, 512, 512, 10000
    hm = synth.heatmap_tiefree_torch(D, H, W, 7, device="cuda")[None, None]
    out = dec.tomo_decode(hm, kernel=3, K=K)
    hmax = torch.nn.functional.max_pool3d(hm, (3, 3, 3), stride=1, padding=(1, 1, 1))
    ref = hm * (hmax == hm).float()
    ts, ti = torch.topk(ref.view(1, -1), K)
    assert torch.equal(out[0, :, 3], ts[0])
    s = out[0, :, 3]
    assert bool((s[:-1] > s[1:]).all())                      # strictly descending (tie-free)
    z = torch.floor(ti.float() / (H * W)).int()
    t = ti.int() - z * H * W
    assert torch.equal(out[0, :, 2], z[0].float())
    assert torch.equal(out[0, :, 1], torch.floor(t.float() / W)[0] + 0.25)
    assert torch.equal(out[0, :, 0], (t % W)[0].float() + 0.25)


def test_sieve_dense_hits_hand_over_to_scan(dec, do):
    """a smooth, nearly flat map (what a random-init detector emits): a few per cent of all voxels reach the
    sampled bound, the sieve's density watchdog abandons the hit-by-hit pass and scan_kernel redoes COLLECT."""
    D, H, W = 48, 256, 256
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((1, 1, D, H, W), device="cuda", generator=g)
    hm = torch.nn.functional.avg_pool3d(x, 7, 1, 3)
    hm = torch.sigmoid(0.05 * hm).contiguous()
    out = dec.tomo_decode(hm, kernel=3, K=900).cpu().numpy()
    st = dec.decode_debug_state()
    assert np.array_equal(bits(out), bits(do.tomo_decode(hm.cpu().numpy(), 3, None, 900)))
    flags, _ = dec.decode_status()
    assert flags == 0


@pytest.mark.parametrize("shape", [(12, 64, 64), (40, 256, 256)])
def test_large_K_takes_the_sort_path(dec, do, shape):
    """K above the rank-sort limit (16 384): the one-CTA bitonic sort writes the picks; small volume = collect-all
    path, large volume = sieve path without the fast final select."""
    D, H, W = shape
    K = 20000
    hm = synth.heatmap_tiefree_np(D, H, W, 77)[None, None]
    out = dec.tomo_decode(cu(hm), kernel=3, K=K).cpu().numpy()
    ref = do.tomo_decode(hm, 3, None, K)
    n = int((ref[0, :, 3] > 0).sum())
    assert np.array_equal(bits(out[:, :n]), bits(ref[:, :n]))
    assert np.array_equal(bits(out), bits(ref))            # zero-score filler rows are canonical in both (index order)


def test_sieve_batch_of_two_reuses_the_workspace(dec, do):
    """B = 2 at a size that takes the sieve path: the second element starts from a freshly initialised state
    (running threshold, histogram, density watchdog) in the same workspace."""
    D, H, W, K = 40, 256, 256, 500
    hm = np.stack([synth.heatmap_tiefree_np(D, H, W, s) for s in (41, 42)])[:, None]
    hm[1] = np.minimum(hm[1], np.float32(0.9))            # a plateau at the top in the second element only
    out = dec.tomo_decode(cu(hm), kernel=3, K=K).cpu().numpy()
    assert np.array_equal(bits(out), bits(do.tomo_decode(hm, 3, None, K)))


def test_repeated_identical_calls_replay_a_graph_and_stay_bit_exact(dec):
    """The third call with the same pointers / shape / K is one cudaGraphLaunch (csrc/decode.cu); results identical to
    the first (plain) call, also after the map's CONTENT changed (the graph holds launches, not data)."""
    import ctypes as C
    from cet_pick_b200 import _lib
    T = _lib.test_lib()
    D, H, W, K = 40, 512, 512, 500
    hm = synth.heatmap_tiefree_torch(D, H, W, 11)[None, None]
    hm2 = synth.heatmap_tiefree_torch(D, H, W, 12)[None, None]
    L = _lib.lib()
    nb = C.c_size_t(0)
    _lib.check(L.cetpick_decode_workspace_bytes(D, H, W, K, C.byref(nb)), "ws")
    ws = torch.empty(nb.value + 256, dtype=torch.uint8, device="cuda")
    wp = (ws.data_ptr() + 255) // 256 * 256
    dets = torch.empty((1, K, 5), device="cuda")

    def call(lib):
        _lib.check(lib.cetpick_decode_f32(hm.data_ptr(), 1, D, H, W, 3, K, 1, None, dets.data_ptr(), None, wp, nb.value,
                                          _lib.stream_ptr()), "decode")
        return dets.clone()

    ref1 = dec.tomo_decode(hm, kernel=3, K=K)
    h0 = T.cetpick_decode_graph_hits()
    outs = [call(T) for _ in range(4)]
    assert T.cetpick_decode_graph_hits() - h0 == 3          # call 1 plain, call 2 captures and launches, 3 and 4 replay
    for o in outs:
        assert torch.equal(o.view(torch.int32), ref1.view(torch.int32))
    hm.copy_(hm2)
    ref2 = dec.tomo_decode(hm2, kernel=3, K=K)
    assert torch.equal(call(T).view(torch.int32), ref2.view(torch.int32))
    assert not torch.equal(ref1, ref2)
