"""Time of one refinement training step on the device (BASELINE.json configs[4]: 128^3 crops): forward + backward + Adam
of `DetectorTrainer`, and the same step in plain PyTorch (oracle ops + autograd + torch.optim.Adam: fp32 with TF32 off,
and TF32) on the same GPU.

    python scripts/bench_train.py [--crops 1] [--size 128] [--iters 3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timed(fn, iters, warmup=1):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--crops", type=int, default=1)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--no-torch", action="store_true")
    a = ap.parse_args()
    import torch
    import synthdata as synth
    from cet_pick_b200.models.model import create_model
    from cet_pick_b200.trains.engine import DetectorTrainer
    from test_gpu_train_net import _labels
    b, d = a.crops, a.size
    sd = {k: v.cuda() for k, v in synth.unet_state_dict_torch(41, 4).items()}
    m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
    m.load_state_dict(sd)
    m = m.cuda()
    tr = DetectorTrainer(m, tau=0.01)
    x = torch.stack([synth.tomogram_torch(d, d, d, seed=10 + i, device="cuda") for i in range(b)])
    gt = _labels(b, d, d // 2, d // 2, 3).cuda()

    def ours():
        tr.zero_grad()
        tr.forward_backward(x, gt)
        tr.step(1e-4)

    out = {"workload": f"training step, {b} x {d}^3 crop(s), unet_4, PU loss, Adam", "ours_ms": timed(ours, a.iters),
           "launches": tr.stats["launches"]}
    # forward FLOPs per crop (214.6 GFLOP at 128^3), backward = 2x
    flops = 3 * 214.6e9 * b * (d / 128.0) ** 3
    out["ours_tflops"] = flops / (out["ours_ms"] * 1e-3) / 1e12
    if not a.no_torch:
        from oracle import train_oracle as to
        names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k and not k.startswith("proj")]
        for tag, tf32 in (("torch_fp32_ms", False), ("torch_tf32_ms", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            sdt = {k: v.clone() for k, v in sd.items()}
            params = [sdt[k].requires_grad_(True) for k in names]
            opt = torch.optim.Adam(params, lr=1e-4)

            def ref():
                opt.zero_grad()
                _, grads, _ = to.training_step(x, gt, sdt, 0.01, param_names=names)
                for p, k in zip(params, names):
                    p.grad = grads[k]
                opt.step()

            out[tag] = timed(ref, a.iters)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
