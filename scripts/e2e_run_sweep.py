"""e2e_run leg of bench.py (TomodetDetector.run incl. heat-map D2H + MRC/pick files on tmpfs) for several run lengths and
writer-thread counts: how much of the per-tomogram time is the final flush amortised over a short run."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                                       # noqa: E402
import bench                                                       # noqa: E402
import synthdata as synth                                          # noqa: E402

dev = torch.device("cuda", 0)
a = types.SimpleNamespace(K=900, nms=3)
D, H, W = 256, 1024, 1024
host_q = []
for i in range(3):
    t = synth.tomogram_torch(D, H, W, seed=i, device=dev)
    hb = torch.empty((D, H, W), dtype=torch.uint8).pin_memory()
    hb.copy_((t * 255.0).round_().to(torch.uint8))
    host_q.append(hb)
    del t
for n, th in ((16, 8), (48, 8), (48, 12), (48, 4)):
    r = bench.leg_e2e_run(dev, a, (D, H, W), n, host_q, threads=th)
    print(f"tomograms {n:3d} writer threads {th:2d}: {r['value']:.2f} tomograms/s, {r['ms_per_tomogram']:.1f} ms each, stages {r['stage_ms_median']}", flush=True)
