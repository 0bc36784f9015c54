"""Decode-only measurement (BASELINE.json configs[2]): 3x3x3 NMS + top-K on one fp32 heat-map.

    python scripts/bench_decode.py [--shape 512,1024,1024] [--K 10000] [--iters 10] [--kind tiefree|peaks]

Prints one JSON line: achieved GB/s = algorithmic bytes (4*D*H*W read once + 20*K written) / CUDA-event
time per decode over a stream of `iters` decodes (ms_median; ms_single_call = one isolated call, events around it,
including the CPU launch latency the GPU idles through), against MEASURED_PEAKS.json's HBM copy bandwidth.  The map (2 GiB at the default shape) is far
larger than L2, so every iteration streams it from HBM.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="512,1024,1024")
    ap.add_argument("--K", type=int, default=10000)
    ap.add_argument("--nms", type=int, default=3)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--kind", default="tiefree", choices=["tiefree", "peaks"])
    a = ap.parse_args()
    import torch
    import synthdata as synth
    from cet_pick_b200.models import decode as dec

    D, H, W = (int(v) for v in a.shape.split(","))
    if a.kind == "tiefree":
        hm = synth.heatmap_tiefree_torch(D, H, W, 7)
    else:
        # smooth random field through sigmoid+clamp: a huge floor plateau plus isolated maxima
        g = torch.Generator(device="cuda").manual_seed(7)
        hm = torch.randn((D, H, W), device="cuda", generator=g)
        hm = torch.nn.functional.avg_pool3d(hm[None, None], 5, 1, 2)[0, 0]
        hm = torch.clamp(torch.sigmoid(8.0 * hm - 6.0), 1e-4, 1 - 1e-4).contiguous()
    hm = hm.reshape(1, 1, D, H, W)
    for _ in range(a.warmup):
        dec.tomo_decode(hm, kernel=a.nms, K=a.K)
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dec.tomo_decode(hm, kernel=a.nms, K=a.K)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    flags, ncand = dec.decode_status()
    ts.sort()
    single = ts[len(ts) // 2]
    # stream of heat-maps: the same decode back to back, events around the whole run (the first kernels of decode i+1 are
    # enqueued while decode i still streams, as in the detector loop where the decode queues behind the forward)
    reps = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            dec.tomo_decode(hm, kernel=a.nms, K=a.K)
        e1.record()
        torch.cuda.synchronize()
        reps.append(e0.elapsed_time(e1) / a.iters)
    reps.sort()
    med = reps[len(reps) // 2]
    nbytes = 4 * D * H * W + 20 * a.K
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    gbs = nbytes / (med * 1e-3) / 1e9
    print(json.dumps({"workload": f"decode {a.kind} {D}x{H}x{W} K={a.K} nms={a.nms}", "ms_median": med, "ms_single_call": single,
                      "ms_min": ts[0], "ms_max": ts[-1], "algorithmic_bytes": nbytes, "achieved_gbs": gbs,
                      "hbm_peak_gbs": hbm, "frac": gbs / hbm, "gvoxels_per_sec": D * H * W / (med * 1e-3) / 1e9,
                      "flags": flags, "n_candidates": ncand}))


if __name__ == "__main__":
    main()
