"""Generate tests/golden/*.npz by RUNNING THE REAL REFERENCE (nextpyp/cet_pick) in the build
container.  The reference has no tests or golden vectors of its own, so
these files are the parity pin for oracle/ (and through it for the CUDA path).

    python tests/golden/make_golden.py            # needs /root/reference (or $CET_PICK_REF)

Inputs are never stored when they can be regenerated from a seed (synthdata);
only reference OUTPUTS are.  Reference environment: torch CPU fp32 (version recorded below).
"""
import io
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import synthdata as synth          # noqa: E402
from oracle import refbridge             # noqa: E402

torch.set_num_threads(max(1, os.cpu_count() or 1))
d = refbridge.decode_module()
u = refbridge.utils_module()
META = dict(torch=torch.__version__, ref="nextpyp/cet_pick @ /root/reference")


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print("wrote", name, {k: getattr(v, "shape", v) for k, v in arrs.items()})


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


# ---------------------------------------------------------------- decode (decode.py:123-155)
def decode_case(name, heat, kernel, K, fiber=False, reg=None, extra=None):
    out = d.tomo_decode(T(heat), kernel=kernel, reg=None if reg is None else T(reg), K=K,
                        if_fiber=fiber).numpy()
    save(name, dets=out, kernel=kernel, K=K, fiber=int(fiber), **(extra or {}))


hm = synth.heatmap_tiefree_np(6, 20, 28, 3)[None, None]
decode_case("decode_tiefree_k3", hm, 3, 40, extra=dict(shape=(6, 20, 28), seed=3))
nm = d._nms(T(hm), 3).numpy()
ts, tz, ty, tx, ti = d._topk(T(nm), K=25)
save("decode_parts", nms=nm, topk_scores=ts.numpy(), topk_zs=tz.numpy(), topk_ys=ty.numpy(),
     topk_xs=tx.numpy(), topk_inds=ti.numpy(), nms_xy=d._nms_xy(T(hm), 3).numpy(),
     nms_z=d._nms_z(T(hm), 3).numpy(), shape=(6, 20, 28), seed=3)

hm = synth.heatmap_tiefree_np(5, 17, 23, 4)[None, None]
decode_case("decode_tiefree_k5", hm, 5, 17, extra=dict(shape=(5, 17, 23), seed=4))
decode_case("decode_tiefree_k1", hm, 1, 9, extra=dict(shape=(5, 17, 23), seed=4))
decode_case("decode_fiber_k3", hm, 3, 30, fiber=True, extra=dict(shape=(5, 17, 23), seed=4))

hm2 = np.stack([synth.heatmap_tiefree_np(4, 12, 20, 5), synth.heatmap_tiefree_np(4, 12, 20, 6)])[:, None]
reg = (synth.uniform_np(77, 2 * 2 * 4 * 12 * 20).reshape(2, 2, 4, 12, 20) - 0.5).astype(np.float32)
decode_case("decode_reg_b2", hm2, 3, 21, reg=reg, extra=dict(shape=(4, 12, 20), seeds=(5, 6), reg_seed=77))

# fp32 index quirk (decode.py:35-41): linear indices >= 2**24 whose float32 rounding crosses a
# plane boundary decode to y = -1.  Plant isolated maxima at the last voxel of high planes.
def quirk_case(name, shape, seed, planes):
    hmq = synth.heatmap_tiefree_np(*shape, seed)[None, None].copy()
    plant = np.array([(z, shape[1] - 1, shape[2] - 1) for z in planes], np.int64)
    vals = (1.5 + 0.01 * np.arange(len(planes))).astype(np.float32)
    for (z, y, x), v in zip(plant, vals):
        hmq[0, 0, z, y, x] = v
    decode_case(name, hmq, 3, 300, extra=dict(shape=shape, seed=seed, plant=plant, plant_vals=vals))


quirk_case("decode_quirk_20x1024x1024", (20, 1024, 1024), 9, (13, 15, 17, 19))
# non-power-of-two plane (h*w = 250000): 70*500*500 = 17.5 M
quirk_case("decode_quirk_70x500x500", (70, 500, 500), 10, (65, 67, 69))

# plateau map: more K than peaks -> filler rows from the clamp floor (tie order unspecified)
hp = synth.heatmap_peaks_np(8, 32, 32, 12, seed=2)[None, None]
decode_case("decode_plateau", hp, 3, 300, extra=dict(shape=(8, 32, 32), n_peaks=12, seed=2))

# _sigmoid (models/utils.py:167-169)
xs = np.linspace(-20, 20, 4001).astype(np.float32)
save("sigmoid", x=xs, y=u._sigmoid(T(xs.copy())).numpy())

# greedy distance NMS (decode.py:42-79) on a tie-free map, threshold = its median
g = synth.heatmap_tiefree_np(6, 12, 14, 8)
thr = float(np.median(g))
sc, co = d.non_maximum_suppression_3d(g, 3, threshold=thr)
save("greedy_nms_d3", scores=sc, coords=co, threshold=thr, d=3, shape=(6, 12, 14), seed=8)
sc, co = d.non_maximum_suppression_3d(g, 5, threshold=thr)
save("greedy_nms_d5", scores=sc, coords=co, threshold=thr, d=5, shape=(6, 12, 14), seed=8)


# ---------------------------------------------------------------- detector forward (unet_small.py:63-97)
def unet_case(name, arch, n_blocks, shape, seed_w, seed_x, K):
    m = refbridge.create_model(arch, {"hm": 1, "proj": 32}, 32, last_k=3)
    sd = synth.unet_state_dict_torch(seed_w, n_blocks)
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd)
    m.eval()
    x = T(synth.tomogram_np(*shape, seed_x))[None]
    with torch.no_grad():
        out = m(x)[-1]
        hm_raw = out["hm"].clone()
        hm_s = u._sigmoid(out["hm"])
        dets = d.tomo_decode(hm_s, kernel=3, reg=None, K=K)
    save(name, hm_raw=hm_raw.numpy(), hm=hm_s.numpy(), proj=out["proj"].numpy().astype(np.float32),
         dets=dets.numpy(), shape=shape, seed_w=seed_w, seed_x=seed_x, n_blocks=n_blocks, K=K)
    return m, sd


unet_case("unet4_even", "unet_4", 4, (5, 32, 48), 317, 1, 30)
unet_case("unet4_odd", "unet_4", 4, (4, 36, 52), 317, 2, 30)      # exercises ceil-pool + autocrop
unet_case("unet5_small", "unet_5", 5, (3, 32, 32), 11, 3, 10)

# ---------------------------------------------------------------- post-process + pick files
# TomodetDetector.post_process / save_detection (tomo_det.py:42-95) on a fixed det tensor.
refbridge.install()
from cet_pick.detectors.tomo_det import TomodetDetector  # noqa: E402

det = TomodetDetector.__new__(TomodetDetector)
hm_pf = synth.heatmap_tiefree_np(30, 40, 44, 12)[None, None]
hm_pf = (hm_pf - hm_pf.min()) / (hm_pf.max() - hm_pf.min())      # spread over [0,1]
hm_pf = hm_pf.astype(np.float32)
dets = d.tomo_decode(T(hm_pf), kernel=3, K=120)
files = {}
for tag, kw in {"plain": {}, "score": dict(with_score=True), "compress": dict(compress=True)}.items():
    det.opt = types.SimpleNamespace(down_ratio=2, out_thresh=0.25, cutoff_z=3, compress=False,
                                    fiber=False, spike=False, with_score=False)
    for k, v in kw.items():
        setattr(det.opt, k, v)
    preds, name = det.post_process(dets.clone(), {"name": ["tomoA"]}, z_dim_tot=30)
    with tempfile.TemporaryDirectory() as td:
        det.save_detection(T(hm_pf), preds, td, None, name=name)
        files[tag] = open(os.path.join(td, "tomoA.txt")).read()
        hm_saved = np.load(os.path.join(td, "tomoA_hm.mrc.npy"))
save("pickfile", dets=dets.numpy(), txt_plain=files["plain"], txt_score=files["score"],
     txt_compress=files["compress"], hm_saved_shape=hm_saved.shape, shape=(30, 40, 44), seed=12,
     keys=np.array(sorted(preds.keys())))
print("done", META)
