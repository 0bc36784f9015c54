"""Build libcetpick_sm100a.so in-tree with nvcc (sm_100a only; cross-compiles without a GPU).

    python -m cet_pick_b200.build [--force] [-v]

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libcetpick_sm100a.so")
TEST_LIB = os.path.join(PKG, "libcetpick_test_sm100a.so")   # product objects + the hooks of include/cetpick_test.h
STAMP = LIB + ".stamp"
SOURCES = ["abi.cu", "decode.cu", "greedy_nms.cu", "preproc.cu", "explore.cu", "sort.cu", "conv_tc.cu", "conv_march.cu", "conv_up.cu", "conv_halo.cu", "conv_stem.cu", "conv_block.cu", "conv_small.cu", "simsiam.cu", "train.cu", "train_net.cu", "unet.cu", "probe.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-cudart", "static", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libcetpick_sm100a.so cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(PKG, "..", "include", "cetpick.h"))
    files.append(os.path.join(PKG, "..", "include", "cetpick_test.h"))
    for p in files:
        if os.path.isfile(p):
            h.update(os.path.basename(p).encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Build the product library and its test twin (every source compiled twice, the second time with
    -DCETPICK_TEST_HOOKS; all nvcc processes run in parallel)."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(TEST_LIB) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for variant, extra in (("", []), ("_t", ["-DCETPICK_TEST_HOOKS"])):
        for s in SOURCES:
            if variant == "" and s == "probe.cu":
                continue                                    # hardware probes exist in the test library only
            obj = os.path.join(objdir, s.replace(".cu", variant + ".o"))
            cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose and not variant else []) + \
                  ["-c", os.path.join(CSRC, s), "-o", obj]
            procs.append((variant, s, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = {"": [], "_t": []}
    for variant, s, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose:
            print(out)
        objs[variant].append(obj)
    for variant, target in (("", LIB), ("_t", TEST_LIB)):
        cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", target] + objs[variant]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    open(STAMP, "w").write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
