import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cet_pick_b200 import synth
from cet_pick_b200.models import decode as dec

D, H, W = 40, 256, 256
hm = synth.heatmap_tiefree_np(D, H, W, 33).copy()
hm[D // 2 - 2:D // 2 + 3] *= np.float32(0.5)
t = torch.from_numpy(hm[None, None]).cuda()
out = dec._topk(t, K=700)
print("fallback state", dec.decode_debug_state(), dec.decode_status())

# timing of the unet forward, layer breakdown comes from ncu later
from cet_pick_b200.models.model import create_model
m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
m.load_state_dict(synth.unet_state_dict_torch(317, 4))
m = m.cuda().eval(); m.compute_proj = False; m.fuse_sigmoid = True
for shape in [(16, 256, 256), (128, 512, 512)]:
    x = synth.tomogram_torch(*shape, seed=1)[None]
    torch.cuda.synchronize()
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = m(x); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        vox = shape[0] * shape[1] * shape[2]
        print(shape, f"forward {ms:.2f} ms  {vox * 100792 / ms / 1e9:.1f} TFLOP/s (no proj)  launches {m.last_launches}")
    hmv = o[-1]["hm"]
    print("hm stats", hmv.min().item(), hmv.max().item(), hmv.std().item())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        e0.record(); d = dec.tomo_decode(hmv, kernel=3, K=900); e1.record(); torch.cuda.synchronize()
        print("decode ms", e0.elapsed_time(e1), dec.decode_status())
