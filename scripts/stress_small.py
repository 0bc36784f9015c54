"""Repeated forwards of the kernels whose producer / issuer warps were restructured late in round 2 (conv_small_kernel,
conv_tc_kernel): every repeat must finish (a lost barrier phase would hang -> run under `timeout`) and reproduce the
first result bit for bit (no atomics on these paths)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthdata as synth                                           # noqa: E402
from cet_pick_b200.models.model import create_model                 # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
t0 = time.time()
m3 = create_model("simsiam3d_18", {"proj": 256, "pred": 256}, 0)
m3.load_state_dict(synth.simsiam3d_state_dict_torch(5))
m3 = m3.cuda().eval()
m2 = create_model("simsiam2d_18", {"proj": 128, "pred": 128}, 128)
m2.load_state_dict(synth.simsiam2d_state_dict_torch(6, out_dim=128))
m2 = m2.cuda().eval()
bad = 0
for B in (1, 3, 37, 148, 149, 1000, 2048):
    x3 = torch.rand(B, 32, 32, 32, device="cuda")
    x2 = torch.rand(B, 1, 32, 32, device="cuda")
    r3 = {k: v.clone() for k, v in m3.forward_test(x3).items()}
    r2 = {k: v.clone() for k, v in m2.forward_test(x2).items()}
    for _ in range(reps):
        o3, o2 = m3.forward_test(x3), m2.forward_test(x2)
        bad += int(any(not torch.equal(o3[k], r3[k]) for k in r3)) + int(any(not torch.equal(o2[k], r2[k]) for k in r2))
    torch.cuda.synchronize()
    print(f"B={B}: {reps} repeats of both encoders ok, mismatches so far {bad}, {time.time() - t0:.1f} s", flush=True)
u = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
u.load_state_dict(synth.unet_state_dict_torch(317, 4))
u = u.cuda().eval()
for mode, shape in (("tf32", (24, 160, 224)), ("bf16", (24, 160, 224)), ("tf32", (5, 36, 52))):
    u.precision = mode
    x = synth.tomogram_torch(*shape, seed=3, device="cuda")[None]
    ref = {k: v.clone() for k, v in u(x)[-1].items()}
    for _ in range(max(10, reps // 4)):
        o = u(x)[-1]
        bad += int(any(not torch.equal(o[k], ref[k]) for k in ref))
    torch.cuda.synchronize()
    print(f"unet {mode} {shape}: repeats ok (hm and proj), mismatches so far {bad}, {time.time() - t0:.1f} s", flush=True)
print("STRESS", "FAILED" if bad else "OK")
sys.exit(1 if bad else 0)
