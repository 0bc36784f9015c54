"""Condense ncu outputs into the small text summaries kept under profiles/.

    python scripts/ncu_summary.py launches <launches.csv>           # per-kernel totals and shares
    python scripts/ncu_summary.py full <report.ncu-rep> [regex]     # key counters per captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if not r[0].isdigit():
            continue
        k = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("unnamed>::", "")
        v = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * scale
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms under ncu (cold-cache, serialised: compare shares)")
    print(f"{'kernel':44s} {'n':>5s} {'ms':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:44]:44s} {n:5d} {t:10.3f} {100 * t / tot:6.1f}%")


def full(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(short, hdr.index(k), units[hdr.index(k)]) for k, short in KEYS if k in hdr]
    kn = hdr.index("Kernel Name")
    print(f"# {path}")
    print("id  " + " ".join(f"{s + '[' + u + ']':>16s}" for s, _, u in idx) + "  kernel")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[kn]).replace("unnamed>::", "")
        if pat and not re.search(pat, name):
            continue
        print(f"{r[0]:3s} " + " ".join(f"{r[i][:16]:>16s}" for _, i, _ in idx) + "  " + name[:40])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
