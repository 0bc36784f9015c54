"""Kernel-time table of one SimSiam encoder forward (torch.profiler / CUPTI): 3-D (configs[3], 8192 x 32^3) and 2-D
(simsiam2d_18, 8192 x 32^2 patches)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthdata as synth                                           # noqa: E402
from cet_pick_b200.models.model import create_model                 # noqa: E402
from torch.profiler import ProfilerActivity, profile                # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
NOPROF = len(sys.argv) > 2 and sys.argv[2] == "noprof"          # under ncu: timing only, no CUPTI profiler
for tag in ("3d", "2d"):
    if tag == "3d":
        m = create_model("simsiam3d_18", {"proj": 256, "pred": 256}, 0)
        m.load_state_dict(synth.simsiam3d_state_dict_torch(5))
        x = torch.rand(B, 32, 32, 32, device="cuda")
    else:
        m = create_model("simsiam2d_18", {"proj": 128, "pred": 128}, 128)
        m.load_state_dict(synth.simsiam2d_state_dict_torch(6, out_dim=128))
        x = torch.rand(B, 1, 32, 32, device="cuda")
    m = m.cuda().eval()
    for _ in range(3):
        m.forward_test(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.forward_test(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"simsiam{tag}: {e0.elapsed_time(e1) / 5:.3f} ms per batch of {B}, launches {m.last_launches}")
    if NOPROF:
        continue
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        m.forward_test(x)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    for i, e in enumerate(evs):
        print(f"  {i:2d} {e.name[:60]:60s} {e.device_time_total / 1000:8.3f} ms")
