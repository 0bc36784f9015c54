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
    for name, path in items[first:first + count]:
        t0 = time.time()
        vol = loader.load_tomos_from_list([name], [path], order=opt.order, compress=opt.compress, denoise=opt.gauss,
                                          dtype=torch.float32)[name]
        torch.cuda.synchronize()
        ret = detector.run(vol[None], {"name": [name], "zdim": vol.shape[0]})
        ret["load"] = time.time() - t0 - ret["tot_time"]           # file read + GPU pre-processing
        print(f"{opt.exp_id} {name}: " + " |".join(f"{k} {v:.3f}s" for k, v in ret.items()))
        for k, v in ret.items():
            stats.setdefault(k, []).append(v)
    return stats


if __name__ == "__main__":
    test(opts().init())
