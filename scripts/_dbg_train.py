import sys, torch, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
import synthdata as synth
from oracle import train_oracle as to
from cet_pick_b200.models.model import create_model
from cet_pick_b200.trains.engine import DetectorTrainer
from test_gpu_train_net import _labels
b, d, h, w = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (1, 6, 32, 48)))
sd = {k: v.cuda() for k, v in synth.unet_state_dict_torch(41, 4).items()}
m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3); m.load_state_dict(sd); m = m.cuda()
tr = DetectorTrainer(m, tau=0.02)
x = torch.stack([synth.tomogram_torch(d, h, w, seed=10 + i, device="cuda") for i in range(b)])
gt = _labels(b, d, (h - 1) // 2 + 1, (w - 1) // 2 + 1, 3).cuda()
tr.zero_grad()
loss, logits = tr.forward_backward(x, gt, want_logits=True)
sdo = {k: v.clone() for k, v in sd.items()}
oloss, grads, ohm = to.training_step(x, gt, sdo, 0.02)
# a float64 oracle on the CPU tells which of the two fp32 results is off
sd64 = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
l64, g64, _ = to.training_step(x.double(), gt.double(), sd64, 0.02)
print("loss", float(loss), float(oloss), float(l64))
gmax = max(float(g.abs().max()) for g in g64.values())
for k, p in m.named_parameters():
    if k.startswith("proj"): continue
    r64 = g64[k].float().cuda()
    den = max(float(r64.abs().max()), 1e-4 * gmax)
    print(f"{k:40s} |g|max {float(r64.abs().max()):.3e}  ours-vs-f64 {float((p.grad - r64).abs().max()) / den:.2e}  torch32-vs-f64 {float((grads[k] - r64).abs().max()) / den:.2e}")
