#!/usr/bin/env python
"""bench.py -- localisation hot path (detector forward + heat-map decode) on N B200s of one node.

    python bench.py [--gpus N --steps K --warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                     # CPU arm: oracle port on host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[1]: a batch of 64 synthetic 1024x1024x256 tomograms, detector
(unet_4, BF16 tensor cores) + decode (3x3x3 NMS, top-K), sharded by tomogram over the ranks with no
data-path collective ("strong" scaling: the 64-tomogram batch is fixed); NCCL only gathers the pick
lists.  One step = one pass over the rank's share of the batch.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tomograms_per_sec"
UNIT = "tomograms/s"
FLOP_PER_VOXEL_NO_PROJ = 102328 - 1536      # unet_4 algorithmic conv FLOPs, 'proj' head skipped
# dram__bytes_read.sum + dram__bytes_write.sum of the 19 tcgen05 conv launches of ONE 1024x1024x256 forward, from the
# `ncu --set full` capture summarised in profiles/r1y_conv_full.txt (ids 0-20 without pool2x2 and stem)
CONV_DRAM_BYTES_PER_FORWARD_C2 = 92.84e9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="tomograms per step over all ranks")
    ap.add_argument("--shape", default="256,1024,1024", help="D,H,W of one tomogram")
    ap.add_argument("--K", type=int, default=900, help="picks per tomogram (docs/refine.md: --K 900)")
    ap.add_argument("--nms", type=int, default=3)
    ap.add_argument("--cpu-sample-slices", type=int, default=8,
                    help="z-slices of one tomogram the CPU baseline times per sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_sustained=d["bf16_tflops_sustained"], bf16_burst=d["bf16_tflops"], hbm=d["hbm_gbs"],
                    src="measured")
    return dict(bf16_sustained=1400.0, bf16_burst=1590.0, hbm=6650.0, src="fallback")


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(shape, K, nms, slices, threads=None):
    """Time the oracle port (torch fp32 restatement of the reference forward + numpy decode) on a
    bounded sample: `slices` z-slices of one tomogram.  The 2-D trunk is per-slice and the 3-D head /
    decode are O(voxels), so cost scales linearly in slices; value = 1 / (t * D / slices)."""
    import numpy as np
    import torch
    from cet_pick_b200 import synth
    from oracle import decode_oracle as do
    from oracle import unet_oracle as uo
    D, H, W = shape
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = synth.unet_state_dict_torch(317, 4)
    x = torch.from_numpy(synth.tomogram_np(slices, H, W, 0))[None]
    t0 = time.perf_counter()
    with torch.no_grad():
        hm = uo.sigmoid_clamp(uo.forward(x, sd, want_proj=False)["hm"]).numpy()
    t1 = time.perf_counter()
    do.tomo_decode(hm, nms, None, min(K, hm.size))
    t2 = time.perf_counter()
    sec_per_tomo = (t2 - t0) * D / slices
    return dict(value=1.0 / sec_per_tomo, unit=UNIT, cores=threads, kind="port",
                sample=f"{slices} of {D} z-slices of one {H}x{W} tomogram (forward {t1 - t0:.2f} s + decode "
                       f"{t2 - t1:.2f} s), scaled linearly to the full tomogram",
                forward_s=t1 - t0, decode_s=t2 - t1)


def run_reference(a, rank):
    """--impl reference: the reference's algorithm on the host cores (oracle port; the Python
    reference itself cannot travel to the GPU box)."""
    if rank != 0:
        return
    shape = tuple(int(v) for v in a.shape.split(","))
    for _ in range(min(a.warmup, 1)):
        cpu_sample(shape, a.K, a.nms, max(1, a.cpu_sample_slices // 2))
    vals = [cpu_sample(shape, a.K, a.nms, a.cpu_sample_slices) for _ in range(max(1, a.steps))]
    best = max(vals, key=lambda v: v["value"])
    v = sum(x["value"] for x in vals) / len(vals)
    D, H, W = shape
    line = {"metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * a.batch / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "gvoxels_per_sec": v * D * H * W / 1e9,
            "config": {"workload": f"batch of {a.batch} synthetic {H}x{W}x{D} tomograms, detector+decode "
                                   "(configs[1])", "arch": "unet_4", "K": a.K, "nms": a.nms},
            "cpu_baseline": {**{k: best[k] for k in ("unit", "cores", "kind", "sample")}, "value": v},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ GPU arm
def run_b200(a, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from cet_pick_b200 import _lib, synth
    from cet_pick_b200.shard import gather_picks, shard_range
    from cet_pick_b200.models.decode import tomo_decode
    from cet_pick_b200.models.model import create_model

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; cet_pick_b200 has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # the native library is built in-tree (a no-op when it is up to date); one rank builds, the others wait
    if rank == 0:
        from cet_pick_b200 import build as _build
        _build.build()
    if world > 1:
        dist.barrier()
    D, H, W = (int(v) for v in a.shape.split(","))
    first, n_local = shard_range(a.batch, rank, world)

    torch.manual_seed(317)                                    # opts.py:47 default seed
    model = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)   # random init (W0)
    model = model.to(dev).eval()
    model.compute_proj = False        # the detector never reads 'proj' (tomo_det.py:26-27)
    model.fuse_sigmoid = True
    lib = _lib.lib()

    # the rank's share of the batch, resident in HBM (generated on the device from the seed)
    pool = [synth.tomogram_torch(D, H, W, seed=first + i, device=dev) for i in range(n_local)]
    # two pinned host staging buffers for the end-to-end leg
    host = [torch.empty((D, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i, hb in enumerate(host):
        hb.copy_(pool[i % max(1, n_local)].cpu() if n_local else torch.zeros(()))
    dets_host = torch.empty((max(1, n_local), a.K, 5), dtype=torch.float32).pin_memory()
    launches = [0]

    def one(x):
        hm = model(x[None])[-1]["hm"]
        launches[0] += model.last_launches
        d = tomo_decode(hm, kernel=a.nms, reg=None, K=a.K)
        launches[0] += lib.cetpick_last_launch_count()
        return d

    def step_resident():
        outs = [one(x) for x in pool]
        if world > 1:                              # the only collective: gather the pick lists
            mine = torch.cat(outs, 0) if outs else torch.empty((0, a.K, 5), device=dev)
            gather_picks(mine, a.batch)
        return outs

    copy_stream = torch.cuda.Stream(dev)
    dbuf = [torch.empty((D, H, W), dtype=torch.float32, device=dev) for _ in range(2)]

    def step_e2e():
        """Host buffers in, host picks out: H2D of every tomogram (pinned, copy stream, double
        buffered against compute) and D2H of its picks are inside the timed region."""
        main = torch.cuda.current_stream(dev)
        free = [torch.cuda.Event(), torch.cuda.Event()]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        for e in free:
            e.record(main)
        for i in range(n_local):
            b = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[b])
                dbuf[b].copy_(host[b], non_blocking=True)
                ready[b].record(copy_stream)
            main.wait_event(ready[b])
            d = one(dbuf[b])
            free[b].record(main)
            dets_host[i].copy_(d[0], non_blocking=True)
        main.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    with ClockSampler(local_rank) as clk:
        launches[0] = 0
        ms_res = timed(step_resident, a.steps, a.warmup)
        n_launch = launches[0] // max(1, a.steps + a.warmup) * a.steps
    clocks = clk.summary()
    ms_e2e = timed(step_e2e, a.steps, min(a.warmup, 1) if a.warmup else 0)

    # roofline of the dominant kernel (conv_tc_kernel): per-launch CUDA events inside the library
    prof = _lib.profile_forward(lambda: model(pool[0][None])) if n_local else []
    conv = [(n, ms, fl) for n, ms, fl in prof if n.startswith("conv:")]
    conv_ms, conv_fl, tot_ms = sum(m for _, m, _ in conv), sum(f for _, _, f in conv), sum(m for _, m, _ in prof)
    pk = peaks()
    layers = {}
    for n, ms, fl in prof:
        e = layers.setdefault(n, [0.0, 0.0]); e[0] += ms; e[1] += fl

    if rank == 0:
        tomo_s = a.batch * a.steps / (ms_res / 1e3)
        e2e_s = a.batch * a.steps / (ms_e2e / 1e3)
        vox = D * H * W
        achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else None
        line = {
            "metric": METRIC, "value": tomo_s, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_res / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "gvoxels_per_sec": tomo_s * vox / 1e9,
            "model_tflops": tomo_s * vox * FLOP_PER_VOXEL_NO_PROJ / 1e12,
            "config": {"workload": f"batch of {a.batch} synthetic {H}x{W}x{D} tomograms, detector+decode, BF16 "
                                   "tensor cores, sharded by tomogram (configs[1])",
                       "arch": "unet_4", "weights": "random init, torch.manual_seed(317)", "K": a.K, "nms": a.nms,
                       "proj_head": "skipped (unused by the detector)", "tomograms_per_rank": n_local,
                       "l2": "inputs larger than L2 (1 GiB per tomogram, distinct tomograms back to back)"},
            "e2e": {"value": e2e_s, "unit": UNIT, "h2d_bytes_per_step": n_local * vox * 4,
                    "d2h_bytes_per_step": n_local * a.K * 5 * 4},
            "gpu_launches": n_launch,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": (achieved / pk["bf16_sustained"]) if achieved else None,
                         "traffic": CONV_DRAM_BYTES_PER_FORWARD_C2 if (D, H, W) == (256, 1024, 1024) else None,
                         "traffic_source": "profiles/r1y_conv_full.txt (ncu --set full, per forward like `achieved`)",
                         "algorithmic_flops_per_forward": conv_fl,
                         "kernel": "conv_march_kernel + conv_halo_kernel + conv_up_kernel (all tcgen05 conv layers of one forward)",
                         "peak_source": pk["src"] + " sustained cuBLAS bf16", "share_of_forward": conv_ms / tot_ms if tot_ms else None,
                         "layers_ms": {k: round(v[0], 3) for k, v in layers.items()},
                         "layers_tflops": {k: round(v[1] / v[0] / 1e9, 1) for k, v in layers.items() if v[0] > 0 and v[1] > 0}},
        }
        if not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample((D, H, W), a.K, a.nms, a.cpu_sample_slices)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line: dict):
    """the ONE JSON line of the contract goes to the process's real stdout"""
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def main():
    global _JSON_OUT
    # Libraries write banners to fd 1 (NCCL prints "NCCL version ..." there when NCCL_DEBUG is WARN/VERSION): keep
    # the real stdout for the JSON line only and send everything else that targets fd 1 to stderr.
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    run_b200(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
