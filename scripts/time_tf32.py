"""TF32-mode and BF16-mode forward time of the detector on one 1024x1024x256 tomogram (configs[1] shape)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import synthdata as synth                                           # noqa: E402
from cet_pick_b200.models.model import create_model                 # noqa: E402

D, H, W = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 1024, 1024)
m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
m.load_state_dict(synth.unet_state_dict_torch(317, 4))
m = m.cuda().eval()
x = synth.tomogram_torch(D, H, W, seed=0, device="cuda")[None]
for mode in ("bf16", "tf32"):
    m.precision = mode
    for _ in range(2):
        out = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = m(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"{mode}: {e0.elapsed_time(e1) / 3:.2f} ms per forward of {D}x{H}x{W}")
