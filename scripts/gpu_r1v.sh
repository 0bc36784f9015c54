#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preproc.py -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import time, torch, numpy as np
from cet_pick_b200.utils import loader
from cet_pick_b200 import synth
v = synth.tomogram_torch(512, 1024, 1024, seed=1, device="cuda")   # stored (nz=512, ny, nx) -> order zxy, compress -> 256 slices
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    rec = loader.load_rec(v, order="zxy", compress=True)
    im = loader.preprocess(rec, denoise=0.8, dtype=torch.float32)
    torch.cuda.synchronize(); t1 = time.time()
    print("load_rec+preprocess 512x1024x1024 -> 256x1024x1024:", round((t1 - t0) * 1e3, 2), "ms", im.shape, im.dtype, float(im.min()), float(im.max()))
    del rec, im
PY
