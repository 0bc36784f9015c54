"""Refinement-step inference from the command line, with the arguments of cet_pick/test.py:62-96:

    python -m cet_pick_b200.test semi --arch unet_4 --load_model model.pth --K 900 --compress --gauss 0.8 \\
        --test_img_txt test_images.txt --out_id out [--with_score] [--order xzy]

`--test_img_txt` is the tab-separated list the reference's datasets read (columns `image_name`, `rec_path`;
datasets/tomo_moco.py:133-139), looked up under `<root>/data` unless absolute.  Every tomogram goes
MRC -> GPU pre-processing (utils/loader.py) -> detector.run (forward + decode) -> `<save_dir>/<out_id>/<name>.txt`
and `<name>_hm.mrc`; `<save_dir>/opt.txt` records the options like the reference's Logger (logger.py:17-40).
With several GPUs (`torchrun --nproc-per-node N -m cet_pick_b200.test ...`) the list is split by tomogram."""
from __future__ import annotations

import csv
import os
import sys
import time

import torch

from .detectors.detector_factory import detector_factory
from .opts import opts
from .shard import shard_range
from .utils import loader
from .utils.async_io import Prefetcher


def read_image_list(path):
    """[(image_name, rec_path)] of a tab-separated list with a header row"""
    with open(path, newline="") as f:
        rows = list(csv.DictReader(f, delimiter="\t"))
    if not rows or "image_name" not in rows[0] or "rec_path" not in rows[0]:
        raise ValueError(f"{path}: expected tab-separated columns image_name and rec_path")
    return [(r["image_name"], r["rec_path"]) for r in rows]


def write_opt_file(opt):
    os.makedirs(opt.save_dir, exist_ok=True)
    os.makedirs(opt.debug_dir, exist_ok=True)
    with open(os.path.join(opt.save_dir, "opt.txt"), "wt") as f:
        f.write(f"==> torch version: {torch.__version__}\n==> Cmd:\n{sys.argv}\n==> Opt:\n")
        for k, v in sorted(vars(opt).items()):
            if not k.startswith("_"):
                f.write(f"  {k}: {v}\n")


def test(opt):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    # Device selection: under torchrun LOCAL_RANK wins (one process per GPU); otherwise the first id of `--gpus`
    # picks the device, like the reference's CUDA_VISIBLE_DEVICES = opt.gpus_str (cet_pick/test.py:66).
    if world > 1 or "LOCAL_RANK" in os.environ:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    else:
        first_gpu = int(str(getattr(opt, "gpus_str", "0")).split(",")[0] or 0)
        if first_gpu >= 0:
            torch.cuda.set_device(first_gpu)
    print(opt)
    if rank == 0:
        write_opt_file(opt)
    os.makedirs(opt.out_path, exist_ok=True)      # every rank: no rank may reach save_detection before the directory exists
    detector = detector_factory[opt.task](opt)
    items = read_image_list(os.path.join(opt.data_dir, opt.test_img_txt))
    first, count = shard_range(len(items), rank, world)
    stats = {}
    # Overlap around the detector (the reference overlaps only the read, test.py:77): a reader thread prepares the next
    # tomogram (MRC read + GPU pre-processing on its own stream, shipped as uint8 levels) while the current one is in the
    # detector, and writer threads put finished heat-maps on disk while the GPU moves on.
    if hasattr(detector, "set_async_write"):
        detector.set_async_write(True)

    def prepare(item):
        name, path = item
        rec = loader.load_rec(path, order=opt.order, compress=opt.compress)
        return loader.preprocess_levels(rec, denoise=opt.gauss)

    t0 = time.time()
    for (name, path), (levels, level_values) in Prefetcher(items[first:first + count], prepare):
        ret = detector.run(levels[None], {"name": [name], "zdim": levels.shape[0], "level_values": level_values})
        ret["load"] = max(0.0, time.time() - t0 - ret["tot_time"])     # wait for the reader (read + pre-processing not hidden)
        t0 = time.time()
        print(f"{opt.exp_id} {name}: " + " |".join(f"{k} {v:.3f}s" for k, v in ret.items()))
        for k, v in ret.items():
            stats.setdefault(k, []).append(v)
    if hasattr(detector, "flush"):
        detector.flush()                                   # every <name>.txt / <name>_hm.mrc is on disk when test() returns
    return stats


if __name__ == "__main__":
    test(opts().init())
