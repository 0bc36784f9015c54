"""Kernel-time table of one training step on a 128^3 crop (torch.profiler / CUPTI)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synthdata as synth                                           # noqa: E402
from cet_pick_b200.models.model import create_model                 # noqa: E402
from cet_pick_b200.trains.engine import DetectorTrainer             # noqa: E402
from test_gpu_train_net import _labels                               # noqa: E402

d = 128
sd = {k: v.cuda() for k, v in synth.unet_state_dict_torch(41, 4).items()}
m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
m.load_state_dict(sd)
m = m.cuda()
tr = DetectorTrainer(m, tau=0.01)
x = torch.stack([synth.tomogram_torch(d, d, d, seed=10, device="cuda")])
gt = _labels(1, d, d // 2, d // 2, 3).cuda()
for _ in range(2):
    tr.zero_grad(); tr.forward_backward(x, gt); tr.step(1e-4)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile                # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.zero_grad(); tr.forward_backward(x, gt); tr.step(1e-4)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=80))
