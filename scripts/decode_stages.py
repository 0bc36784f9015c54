"""Where the decode's time goes, warm and in-stream (not under a profiler): the test build's
`cetpick_decode_set_stop_stage(n)` cuts the launch sequence after stage n; differences of the CUDA-event medians
are the stages' in-stream costs.

    python scripts/decode_stages.py [--shape 512,1024,1024] [--K 10000] [--kind tiefree|peaks]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["all", "init", "+sample select", "+sieve", "+gated fall-backs, EQ", "+tail (select, compact, order, rows)"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="512,1024,1024")
    ap.add_argument("--K", type=int, default=10000)
    ap.add_argument("--iters", type=int, default=15)
    ap.add_argument("--kind", default="tiefree", choices=["tiefree", "peaks"])
    a = ap.parse_args()
    import torch
    import synthdata as synth
    from cet_pick_b200 import _lib
    T = _lib.test_lib()
    _lib.lib = lambda: T                       # the decode below goes through the test twin of the product library
    from cet_pick_b200.models import decode as dec
    D, H, W = (int(v) for v in a.shape.split(","))
    if a.kind == "tiefree":
        hm = synth.heatmap_tiefree_torch(D, H, W, 7)
    else:
        g = torch.Generator(device="cuda").manual_seed(7)
        hm = torch.randn((D, H, W), device="cuda", generator=g)
        hm = torch.nn.functional.avg_pool3d(hm[None, None], 5, 1, 2)[0, 0]
        hm = torch.clamp(torch.sigmoid(8.0 * hm - 6.0), 1e-4, 1 - 1e-4).contiguous()
    hm = hm.reshape(1, 1, D, H, W)

    def med(stage):
        T.cetpick_decode_set_stop_stage(stage)
        ts = []
        for i in range(a.iters + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dec.tomo_decode(hm, kernel=3, K=a.K)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    full = med(0)
    prev = 0.0
    for n in range(1, 6):
        t = med(n)
        print(f"stage {n} {STAGES[n]:32s} cumulative {t * 1e3:8.1f} us   stage {1e3 * (t - prev):8.1f} us")
        prev = t
    print(f"whole decode {full * 1e3:8.1f} us")
    d = dec.decode_debug_state()
    print(f"tail kernel, CTA 0 (ns): select {d['sel_prefix']} ({d['sel_kleft']} passes, csel_done {d['csel_done']}), "
          f"compaction + barrier {d['n_gt']}, ordering + rows {d['hit_total']}; candidates {d['n_final']}; "
          f"class width 2^{d['csel_wl']}, rank inside {d['csel_kleft']}, t0key {d['t0key']:#x}, kth key {d['kth_comp_hi']:#x}, t_run {d['t_run']:#x}")
    T.cetpick_decode_set_stop_stage(0)


if __name__ == "__main__":
    main()
