import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import synthdata as synth
from cet_pick_b200.models.model import create_model
m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
m.load_state_dict(synth.unet_state_dict_torch(317, 4)); m = m.cuda().eval(); m.compute_proj=False; m.fuse_sigmoid=True
for shape in [(2, 64, 96), (5, 64, 96), (3, 128, 256), (4, 1024, 1024)]:
    D,H,W = shape
    xf = synth.tomogram_np(D,H,W,11); q = np.rint(xf*255).astype(np.uint8)
    a = m(torch.from_numpy(xf)[None].cuda())[-1]["hm"]; torch.cuda.synchronize()
    t0=time.time()
    try:
        b = m(torch.from_numpy(q)[None].cuda())[-1]["hm"]; torch.cuda.synchronize()
        print(shape, "u8 ok", time.time()-t0, "equal", torch.equal(a,b), float((a-b).abs().max()))
    except Exception as e:
        print(shape, "u8 FAILED after", time.time()-t0, str(e)[:300]); break
