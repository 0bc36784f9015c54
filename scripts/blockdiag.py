"""bring-up helper for csrc/conv_block.cu: one case per process (a device fault kills the context)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from cet_pick_b200 import _lib as L
n, h, w, c1, nsrc, pool = [int(v) for v in sys.argv[1:7]]
ident = len(sys.argv) > 7 and sys.argv[7] == "ident"      # conv2 = identity: the output is conv1's bf16 row ring
g = torch.Generator(device="cuda").manual_seed(1)
srcs = [torch.randn(n, h, w, c1, device="cuda", generator=g).bfloat16() for _ in range(nsrc)]
w1 = (torch.randn(32, nsrc * c1, 3, 3, device="cuda", generator=g) / (3 * (nsrc * c1) ** 0.5)).bfloat16()
w2 = (torch.randn(32, 32, 3, 3, device="cuda", generator=g) / (3 * 32 ** 0.5)).bfloat16()
if ident:
    w2 = torch.zeros_like(w2)
    for c in range(32):
        w2[c, c, 1, 1] = 1.0
b1 = torch.randn(32, device="cuda", generator=g) * 0.2
b2 = torch.randn(32, device="cuda", generator=g) * (0.0 if ident else 0.2)
out = torch.full((n, h, w, 32), float("nan"), device="cuda", dtype=torch.bfloat16)
pl = torch.full((n, (h + 1) // 2, (w + 1) // 2, 32), float("nan"), device="cuda", dtype=torch.bfloat16) if pool else None
w1h, w2h, b1h, b2h = w1.float().cpu().contiguous(), w2.float().cpu().contiguous(), b1.cpu().contiguous(), b2.cpu().contiguous()
import ctypes
dbg = ctypes.POINTER(ctypes.c_uint32)()
L.test_lib().cetpick_block_debug_buffer(ctypes.byref(dbg))
t0 = time.time()
rc = L.test_lib().cetpick_conv_block_bf16(nsrc, srcs[0].data_ptr(), srcs[1].data_ptr() if nsrc > 1 else None, c1, n, h, w,
                                     w1h.data_ptr(), b1h.data_ptr(), w2h.data_ptr(), b2h.data_ptr(),
                                     out.data_ptr(), pl.data_ptr() if pool else None, L.stream_ptr())
print(sys.argv[1:7], "rc", rc, L.lib().cetpick_last_cuda_error().decode(), "t=%.2fs" % (time.time() - t0))
if rc != 0:
    n_to = dbg[0]
    print("  timeouts:", n_to)
    names = {1: "in_empty", 2: "bar_w", 3: "c1_empty", 4: "in_full", 5: "c2_empty", 6: "ring_full", 7: "relay ring_empty",
             8: "c1_full", 9: "epi ring_empty", 10: "nb_empty[left]", 11: "nb_empty[right]", 12: "c2_full"}
    for k in range(min(n_to, 60)):
        e = [dbg[4 + 4 * k + i] for i in range(4)]
        print("   cta %3d warp %2d waits %-18s a=%d b=%d" % (e[0] >> 8, e[0] & 255, names.get(e[1], e[1]), e[2], e[3]))
if rc == 0 and os.environ.get("TIME"):
    import ctypes as C
    # time the kernel alone: pre-packed weights through the product path are not exposed, so time the hook minus its
    # fixed host work by differencing two sizes is overkill: use CUDA events around 5 calls (the hook synchronises)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        ev0.record()
        L.test_lib().cetpick_conv_block_bf16(nsrc, srcs[0].data_ptr(), srcs[1].data_ptr() if nsrc > 1 else None, c1, n, h, w,
                                        w1h.data_ptr(), b1h.data_ptr(), w2h.data_ptr(), b2h.data_ptr(),
                                        out.data_ptr(), pl.data_ptr() if pool else None, L.stream_ptr())
        ev1.record(); torch.cuda.synchronize()
        ts.append(ev0.elapsed_time(ev1))
    print("  ms per call (incl. weight upload): min %.3f median %.3f" % (min(ts), sorted(ts)[2]),
          {k: v for k, v in os.environ.items() if k.startswith("CETPICK_BLOCK")})
if rc == 0 and not os.environ.get("TIME"):
    xin = torch.cat(srcs, 3).float().permute(0, 3, 1, 2)
    r1 = F.relu(F.conv2d(xin, w1.float(), b1, padding=1)).bfloat16().float()
    ref = F.relu(F.conv2d(r1, w2.float(), b2, padding=1)).permute(0, 2, 3, 1)
    d = (out.float() - ref).abs()
    bad = (torch.nan_to_num(d, nan=99.) > 0.05)
    if bad.any():
        bm = bad.any(dim=3)[0]
        for r in range(bm.shape[0]):
            print("   row %2d: " % r + "".join("X" if v else "." for v in bm[r].tolist()))
    print("  nan", int(torch.isnan(out.float()).sum()), "max err", float(torch.nan_to_num(d, nan=99.).max()),
          "bad cols", sorted(set((torch.nan_to_num(d, nan=99.) > 0.05).nonzero()[:, 2].tolist()))[:20],
          "bad rows", sorted(set((torch.nan_to_num(d, nan=99.) > 0.05).nonzero()[:, 1].tolist()))[:20])
