"""Parity at the sizes the performance numbers are quoted on (BASELINE.json configs[1] and configs[2]).

configs[2]: decode of a 512x1024x1024 fp32 heat-map, K = 10 000 -- bit-exact against the reference's own op
sequence (cet_pick/models/decode.py:27-41,84,123-155: max_pool3d, ==, .float(), *, topk, fp32
`_convert_1d_to_3d`, +0.25, cat) executed by PyTorch on the same GPU.  2^29 voxels: almost every linear index is
>= 2^24, so the reference's fp32 index arithmetic is in play for nearly every pick.  (h*w = 2^20 is a power of
two, so torch-CUDA's multiply-by-reciprocal form of `inds.float() / (h*w)` rounds like the CPU's true division.)

configs[1]: the detector forward of ONE 1024x1024x256 tomogram on the seeded non-degenerate weights, against the
fp32 oracle on z-slabs.  The 2-D trunk is per slice and the 3-D head reaches +-3 slices (feature_head.0,
feature_head.2, hm: one slice each), so the oracle on slices [z0-3, z0+8+3) pins the 8 core slices exactly
(SURVEY.md section 7.7 probe: halo 3 is exact, a slab at a true volume end needs no halo on that side).
"""
import numpy as np
import pytest
import torch

import synthdata as synth

pytestmark = pytest.mark.gpu


def reference_decode_torch(hm, kernel, K):
    """decode.py:123-155 (`tomo_decode`, reg=None, if_fiber=False) op by op in PyTorch on hm's device, freeing
    the full-size temporaries as it goes (five 2 GiB tensors at configs[2])."""
    B, C, D, H, W = hm.shape
    pad = (kernel - 1) // 2
    hmax = torch.nn.functional.max_pool3d(hm, (3, kernel, kernel), stride=1, padding=(1, pad, pad))   # :30-31
    keep = (hmax == hm)
    del hmax
    keep = keep.float()                                                                                   # :32
    heat = hm * keep                                                                                      # :33
    del keep
    scores, inds = torch.topk(heat.view(B, C, -1), K)                                                     # :84
    del heat
    z = torch.floor(inds.float() / (H * W)).int()                                                         # :36
    t = inds.int() - (z * H * W)                                                                          # :37
    y = torch.floor(t.float() / W)                                                                        # :38
    x = t % W                                                                                             # :39
    xs = x.view(B, K, 1) + 0.25                                                                           # :141
    ys = y.view(B, K, 1) + 0.25
    zs = z.view(B, K, 1)
    sc = scores.view(B, K, 1).float()
    centers = torch.cat([xs.float(), ys.float(), zs.float()], dim=2)                                      # :153
    return torch.cat([centers, sc, sc], dim=2), inds.view(B, K)                                           # :154


@pytest.fixture(scope="module")
def c2_map():
    D, H, W = 512, 1024, 1024
    free, _ = torch.cuda.mem_get_info()
    if free < 24 << 30:
        pytest.skip("configs[2] parity needs ~24 GB of free device memory")
    return synth.heatmap_tiefree_torch(D, H, W, 2, device="cuda")[None, None]


def test_config2_decode_tiefree_bit_exact_vs_torch_cuda(c2_map):
    """Hm(i): all 2^29 values distinct => order fully determined; all five columns and the row order identical."""
    from cet_pick_b200.models import decode as dec
    K = 10000
    out = dec.tomo_decode(c2_map, kernel=3, K=K)
    flags, ncand = dec.decode_status()
    ref, inds = reference_decode_torch(c2_map, 3, K)
    assert flags == 0 and ncand >= K
    assert torch.equal(out.view(torch.int32), ref.view(torch.int32))
    s = out[0, :, 3]
    assert bool((s[:-1] > s[1:]).all())                                   # strictly descending
    assert int((inds >= (1 << 24)).sum()) > K * 9 // 10                   # the fp32 index arithmetic was exercised
    # the quirk itself: some rows carry the reference's wrong-by-design coordinates (y = -1 + 0.25)
    exact_y = ((inds[0] % (1024 * 1024)) // 1024).float() + 0.25
    print("configs[2] tie-free: rows whose fp32 y differs from the exact y:", int((out[0, :, 1] != exact_y).sum()))


def test_config2_decode_index_quirk_bit_exact_vs_torch_cuda(c2_map):
    """Hm(iii): maxima planted at y = H-1, x >= W-16 of high planes, where float32(ind) rounds across the plane
    boundary (ind = 315621375 -> (z,y,x) = (301,-1,1023))."""
    from cet_pick_b200.models import decode as dec
    K, H, W = 10000, 1024, 1024
    hm = c2_map.clone()
    planes = [17, 100, 300, 302, 400, 511]          # no two adjacent in z: every planted maximum survives the NMS
    vals = []
    for i, z in enumerate(planes):
        for j, x in enumerate(range(W - 16, W, 3)):
            v = 1.5 + 0.01 * i + 0.001 * j
            hm[0, 0, z, H - 1, x] = v
            vals.append(v)
    out = dec.tomo_decode(hm, kernel=3, K=K)
    ref, inds = reference_decode_torch(hm, 3, K)
    del hm
    assert torch.equal(out.view(torch.int32), ref.view(torch.int32))
    top = out[0, :len(vals)]
    assert bool((top[:, 3] > 1.0).all())                                  # the planted maxima lead the list
    assert bool((top[:, 1] == -0.75).any())                               # y = -1 + 0.25: the quirk is reproduced


def test_config2_decode_properties_plateau_map():
    """A realistic map at configs[2] size (clamped sigmoid floor = one huge plateau + smooth peaks): size-independent
    properties -- descending scores, every pick an NMS survivor of its 3x3x3 window, and the same multiset of scores
    as torch.topk of the reference NMS output."""
    from cet_pick_b200.models import decode as dec
    D, H, W, K = 512, 1024, 1024, 10000
    free, _ = torch.cuda.mem_get_info()
    if free < 24 << 30:
        pytest.skip("needs ~24 GB of free device memory")
    g = torch.Generator(device="cuda").manual_seed(5)
    hm = torch.full((D, H, W), 1e-4, device="cuda")
    n_peaks = 20000
    pz = torch.randint(2, D - 2, (n_peaks,), device="cuda", generator=g)
    py = torch.randint(2, H - 2, (n_peaks,), device="cuda", generator=g)
    px = torch.randint(2, W - 2, (n_peaks,), device="cuda", generator=g)
    amp = 0.2 + 0.79 * torch.rand(n_peaks, device="cuda", generator=g)
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                w = float(np.exp(-0.5 * (dz * dz + dy * dy + dx * dx)))
                hm.index_put_((pz + dz, py + dy, px + dx), amp * w, accumulate=True)
    hm.clamp_(1e-4, 1 - 1e-4)
    hm = hm[None, None]
    out = dec.tomo_decode(hm, kernel=3, K=K)
    ref, _ = reference_decode_torch(hm, 3, K)
    s = out[0, :, 3]
    assert bool((s[:-1] >= s[1:]).all())
    assert torch.equal(s, ref[0, :, 3])                                   # same scores in the same (descending) order
    # rows with a unique score must agree in every column
    uniq = torch.ones(K, dtype=torch.bool, device="cuda")
    uniq[1:] &= s[1:] != s[:-1]
    uniq[:-1] &= s[:-1] != s[1:]
    assert int(uniq.sum()) > K // 2
    assert torch.equal(out[0][uniq].view(torch.int32), ref[0][uniq].view(torch.int32))


# ----------------------------------------------------------------------------------------------------------- configs[1]
def test_config1_forward_slabs_vs_oracle_and_picks():
    """One 1024x1024x256 tomogram: heat-map of three 8-slice slabs (both volume ends and the middle) against the
    fp32 oracle, max-abs <= 1e-2 (BF16 operands, fp32 accumulation); picks of each slab within one voxel."""
    from oracle import unet_oracle as uo, decode_oracle as do
    from cet_pick_b200.models.decode import tomo_decode
    from cet_pick_b200.models.model import create_model
    D, H, W, CORE, HALO, K = 256, 1024, 1024, 8, 3, 300
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("configs[1] parity needs ~40 GB of free device memory")
    sd = synth.unet_state_dict_torch(317, 4)
    m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.compute_proj, m.fuse_sigmoid = False, True
    x = synth.tomogram_torch(D, H, W, seed=0, device="cuda")
    hm = m(x[None])[-1]["hm"]
    assert hm.shape == (1, 1, D, H // 2, W // 2)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    worst = 0.0
    for z0 in (0, 124, D - CORE):
        lo, hi = max(0, z0 - HALO), min(D, z0 + CORE + HALO)
        xs = x[lo:hi].cpu()[None]
        with torch.no_grad():
            ref = uo.sigmoid_clamp(uo.forward(xs, sd, want_proj=False)["hm"])[0, 0, z0 - lo:z0 - lo + CORE].numpy()
        got = hm[0, 0, z0:z0 + CORE]
        err = float(np.abs(got.cpu().numpy() - ref).max())
        worst = max(worst, err)
        print(f"configs[1] slab z=[{z0},{z0 + CORE}): hm max-abs err {err:.3e}; ref range {ref.min():.4f}..{ref.max():.4f}")
        assert err <= 1e-2
        # picks of the slab (decode of the same 8 slices from each side) agree within one voxel
        dets = tomo_decode(got[None, None].contiguous(), kernel=3, K=K).cpu().numpy()[0]
        rdets = do.tomo_decode(ref[None, None], 3, None, K)[0]
        g = got.cpu().numpy()
        ours = {(int(r[0] - 0.25), int(r[1] - 0.25), int(r[2])) for r in dets}
        kth = dets[-1, 3]
        checked = listed = 0
        for r in rdets:
            px, py, pz = int(r[0] - 0.25), int(r[1] - 0.25), int(r[2])
            if not (2 <= pz < CORE - 2 and 2 <= py < H // 2 - 2 and 2 <= px < W // 2 - 2):
                continue
            cube = ref[pz - 2:pz + 3, py - 2:py + 3, px - 2:px + 3].copy()
            cube[1:4, 1:4, 1:4] = -1.0
            if r[3] - cube.max() <= 2 * err:
                continue                                # broad peak: its argmax is not stable under the tolerance
            oc = g[pz - 2:pz + 3, py - 2:py + 3, px - 2:px + 3]
            dz, dy, dx = np.unravel_index(np.argmax(oc), oc.shape)
            assert max(abs(dz - 2), abs(dy - 2), abs(dx - 2)) <= 1
            checked += 1
            if oc.max() > kth:
                assert (px + dx - 2, py + dy - 2, pz + dz - 2) in ours
                listed += 1
        print(f"  {checked} well-conditioned reference picks checked, {listed} of them in our top-{K}")
        assert checked > 0
    print(f"configs[1]: worst slab error {worst:.3e} (tolerance 1e-2)")
