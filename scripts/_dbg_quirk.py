import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import synthdata as synth
from cet_pick_b200.models import decode as dec
g = np.load("/root/repo/tests/golden/decode_quirk_20x1024x1024.npz")
D, H, W = [int(v) for v in g["shape"]]
hm = synth.heatmap_tiefree_np(D, H, W, int(g["seed"]))[None, None].copy()
for (z, y, x), v in zip(g["plant"], g["plant_vals"]):
    hm[0, 0, z, y, x] = v
out = dec.tomo_decode(torch.from_numpy(hm).cuda(), kernel=3, K=int(g["K"])).cpu().numpy()[0]
ref = g["dets"][0]
bad = np.where((out.view(np.uint32) != ref.view(np.uint32)).any(1))[0]
print("K", g["K"], "bad rows", len(bad), bad[:20])
for r in bad[:8]:
    print(r, out[r], ref[r])
print(dec.decode_status(), dec.decode_debug_state() if hasattr(dec, "decode_debug_state") else "")
# where do the reference rows appear in ours?
for r in bad[:8]:
    m = np.where((out.view(np.uint32) == ref[r].view(np.uint32)).all(1))[0]
    print("ref row", r, "found in ours at", m)
