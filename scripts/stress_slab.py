"""slab-mode forward stress (config0 size), one process: prints ok / the CUDA error"""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthdata as synth
from cet_pick_b200.models.model import create_model
m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
m.load_state_dict(synth.unet_state_dict_torch(317, 4)); m = m.cuda().eval()
m.compute_proj, m.fuse_sigmoid = False, True
x = synth.tomogram_torch(128, 512, 512, seed=0, device="cuda")[None]
hm = m(x)[-1]["hm"].clone(); torch.cuda.synchronize()
m.slab_z = 48
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
t0 = time.time()
try:
    for i in range(n):
        hm2 = m(x)[-1]["hm"]; torch.cuda.synchronize()
        assert float((hm2 - hm).abs().max()) <= 2e-7
    print("ok", n, "slab forwards", round(time.time() - t0, 2), "s", os.environ.get("CETPICK_LIB", "default"))
except Exception as e:
    print("FAIL at", i, type(e).__name__, str(e)[:80], os.environ.get("CETPICK_LIB", "default"))
