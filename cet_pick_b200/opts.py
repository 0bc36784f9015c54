"""Flag system of the hot path: accepts the command lines written for cet_pick/opts.py (same flag
names, types and defaults, :14-189) and derives the same fields in parse() (:193-269) and
init()/update_dataset_info_and_set_heads() (:271-332), so `test.py semi --arch unet_4 ...` lines
from docs/refine.md keep working.  Table-driven instead of one add_argument call per flag.
"""
from __future__ import annotations

import argparse
import math
import os
import types


def list_of_floats(arg):
    return list(map(float, arg.split(",")))


# (flag, kwargs) in the reference's order; help texts abbreviated.
_FLAGS = [
    ("--dataset", dict(default="semi")), ("--exp_id", dict(default="default")),
    ("--test", dict(action="store_true")), ("--debug", dict(type=int, default=4)),
    ("--load_model", dict(default="")), ("--pretrain_model", dict(default="")),
    ("--resume", dict(action="store_true")), ("--fiber", dict(action="store_true")),
    ("--spike", dict(action="store_true")),
    ("--gpus", dict(default="0")), ("--num_workers", dict(type=int, default=4)),
    ("--not_cuda_benchmark", dict(action="store_true")), ("--seed", dict(type=int, default=317)),
    ("--world-size", dict(default=-1, type=int)), ("--rank", dict(default=-1, type=int)),
    ("--dist-url", dict(default="env://", type=str)), ("--dist-backend", dict(default="nccl", type=str)),
    ("--local_rank", dict(default=-1, type=int)),
    ("--print_iter", dict(type=int, default=0)), ("--hide_data_time", dict(action="store_true")),
    ("--save_all", dict(action="store_true")), ("--metric", dict(default="loss")),
    ("--vis_thresh", dict(type=float, default=0.3)),
    ("--debugger_theme", dict(default="white", choices=["white", "black"])),
    ("--arch", dict(default="unet_4")),
    # not in the reference: operand precision of the detector's tensor-core convolutions (bf16: heat-map within 1e-2 of
    # the fp32 reference; tf32: within 1e-4)
    ("--precision", dict(default="bf16", choices=["bf16", "tf32"])), ("--last_k", dict(type=int, default=3)),
    ("--head_conv", dict(type=int, default=-1)), ("--down_ratio", dict(type=int, default=2)),
    ("--pretrained_model", dict(type=str, default=None)),
    ("--input_res", dict(type=int, default=-1)), ("--input_h", dict(type=int, default=-1)),
    ("--input_w", dict(type=int, default=-1)),
    ("--lr", dict(type=float, default=1e-3)), ("--lr_step", dict(type=str, default="200, 400, 600")),
    ("--num_epochs", dict(type=int, default=140)), ("--lr_decay_rate", dict(type=float, default=0.1)),
    ("--cosine", dict(action="store_true")), ("--warm", dict(action="store_true")),
    ("--contrastive", dict(action="store_true")),
    ("--batch_size", dict(type=int, default=1)), ("--master_batch_size", dict(type=int, default=-1)),
    ("--num_iters", dict(type=int, default=-1)), ("--val_intervals", dict(type=int, default=5)),
    ("--trainval", dict(action="store_true")), ("--bbox", dict(type=int, default=32)),
    ("--translation_ratio", dict(type=float, default=0.5)), ("--cr_weight", dict(type=float, default=0.1)),
    ("--thresh", dict(type=float, default=0.5)), ("--temp", dict(type=float, default=0.07)),
    ("--tau", dict(type=float, default=0.1)), ("--nclusters", dict(type=int, default=3)),
    ("--nheads", dict(type=int, default=1)), ("--names", dict(type=str)),
    ("--nms", dict(type=int, default=3)), ("--cutoff_z", dict(type=int, default=10)),
    ("--K", dict(type=int, default=200)), ("--not_prefetch_test", dict(action="store_true")),
    ("--fix_res", dict(action="store_true")), ("--keep_res", dict(action="store_true")),
    ("--out_thresh", dict(type=float, default=0.25)), ("--with_score", dict(action="store_true")),
    ("--pn", dict(action="store_true")), ("--ge", dict(action="store_true")),
    ("--distance_cutoff", dict(type=float, default=15)), ("--r2_cutoff", dict(type=float, default=30)),
    ("--curvature_cutoff", dict(type=float, default=0.003)), ("--distance_scale", dict(type=float, default=2)),
    ("--train_img_txt", dict(type=str, default="train_images.txt")),
    ("--train_coord_txt", dict(type=str, default="train_coords.txt")),
    ("--val_img_txt", dict(type=str)), ("--val_coord_txt", dict(type=str)),
    ("--test_img_txt", dict(type=str, default="test_images.txt")),
    ("--test_coord_txt", dict(type=str, default="test_coords.txt")),
    ("--compress", dict(action="store_true")), ("--gauss", dict(type=float, default=0)),
    ("--cluster_head", dict(action="store_true")), ("--out_id", dict(type=str, default="output")),
    ("--order", dict(type=str, default="xzy")), ("--dog", dict(type=list_of_floats, default=[2.5, 5])),
    # cet_pick_b200 additions (default = reference behaviour)
    ("--return_proj", dict(action="store_true")),     # also materialise the unused 'proj' head map
]

_DATASETS = {
    "tomo": ([512, 512], 1), "cr": ([64, 64], 1), "semi": ([64, 64], 1), "semiclass": ([64, 64], 1),
    "semi3d": ([64, 64], 1), "fs": ([128, 128], 1), "simsiam": ([24, 24], 256), "scan": ([24, 24], 256),
    "denoise": ([64, 64], 256), "moco": ([32, 32], 256),
    # the exploration datasets' class attributes (datasets/tomo_pre_proj_angle_select_new3d_vol.py:26-27, ..._new2d3d.py:26-27)
    "simsiam3d": ([256, 256], 1), "simsiam2d3d": ([256, 256], 1),
}


class opts(object):
    def __init__(self):
        self.parser = argparse.ArgumentParser()
        self.parser.add_argument("task", default="semi")
        for flag, kw in _FLAGS:
            self.parser.add_argument(flag, **kw)

    def parse(self, args=""):
        opt = self.parser.parse_args() if args == "" else self.parser.parse_args(args)
        opt.gpus_str = opt.gpus
        gl = [int(g) for g in opt.gpus.split(",")]
        opt.gpus = list(range(len(gl))) if gl[0] >= 0 else [-1]
        opt.lr_step = [int(i) for i in opt.lr_step.split(",")]
        opt.fix_res = not opt.keep_res
        print("Fix size testing." if opt.fix_res else "Keep resolution testing.")
        if opt.head_conv == -1:
            if opt.task in ("simsiam", "simsiam2d3d", "simsiam3d"):
                opt.head_conv = 128
            if opt.task in ("semi", "semiclass"):
                opt.head_conv = 32
        opt.pad = 127 if "hourglass" in opt.arch else 31
        opt.num_stacks = 2 if opt.arch == "hourglass" else 1
        if opt.warm:
            opt.warmup_from, opt.warm_epochs = 0.01, 10
            if opt.cosine:   # the reference reads an undefined opt.learning_rate here (Appendix C)
                eta_min = opt.lr * (opt.lr_decay_rate ** 3)
                opt.warmup_to = eta_min + (opt.lr - eta_min) * (
                    1 + math.cos(math.pi * opt.warm_epochs / opt.num_epochs)) / 2
            else:
                opt.warmup_to = opt.lr
        if opt.val_intervals >= 0 and opt.val_img_txt is None and opt.val_coord_txt is None:
            print("No validation files but validation interval is greater than 1...using training files for validation")
            opt.val_img_txt, opt.val_coord_txt = opt.train_img_txt, opt.train_coord_txt
        if opt.trainval:
            opt.val_interval = 100000000
        if opt.debug > 0:
            opt.num_workers = 0
            opt.gpus = [opt.gpus[0]]
            opt.master_batch_size = -1
        if opt.master_batch_size == -1:
            opt.master_batch_size = opt.batch_size // len(opt.gpus)
        rest = opt.batch_size - opt.master_batch_size
        opt.chunk_sizes = [opt.master_batch_size]
        for i in range(len(opt.gpus) - 1):
            c = rest // (len(opt.gpus) - 1)
            if i < rest % (len(opt.gpus) - 1):
                c += 1
            opt.chunk_sizes.append(c)
        print("Training chunk_sizes:", opt.chunk_sizes)
        opt.root_dir = os.getcwd()
        opt.data_dir = os.path.join(opt.root_dir, "data")
        opt.exp_dir = os.path.join(opt.root_dir, "exp", opt.task)
        opt.save_dir = os.path.join(opt.exp_dir, opt.exp_id)
        opt.debug_dir = os.path.join(opt.save_dir, "debug")
        if opt.task == "scan2d3d":
            opt.simsiam_dir = os.path.join(opt.root_dir, "exp", "simsiam2d3d", opt.exp_id)
        elif opt.task == "scan":
            opt.simsiam_dir = os.path.join(opt.root_dir, "exp", "simsiam", opt.exp_id)
        opt.out_path = os.path.join(opt.save_dir, opt.out_id)
        print("The output will be saved to ", opt.save_dir)
        if opt.resume and opt.load_model == "":
            mp = opt.save_dir[:-4] if opt.save_dir.endswith("TEST") else opt.save_dir
            opt.load_model = os.path.join(mp, "model_last.pth")
        return opt

    def update_dataset_info_and_set_heads(self, opt, dataset):
        input_h, input_w = dataset.default_resolution
        opt.num_classes = dataset.num_classes
        input_h = opt.input_res if opt.input_res > 0 else input_h
        input_w = opt.input_res if opt.input_res > 0 else input_w
        opt.input_h = opt.input_h if opt.input_h > 0 else input_h
        opt.input_w = opt.input_w if opt.input_w > 0 else input_w
        opt.output_h = opt.input_h // opt.down_ratio
        opt.output_w = opt.input_w // opt.down_ratio
        opt.input_res = max(opt.input_h, opt.input_w)
        opt.output_res = max(opt.output_h, opt.output_w)
        t = opt.task
        if t == "tomo":
            opt.heads = {"hm": opt.num_classes, "proj": 16}
        elif t in ("cr", "semi", "semi3d", "semiclass"):
            opt.heads = {"hm": opt.num_classes, "proj": opt.head_conv}
        elif t == "fs":
            opt.heads = {"proj": 16}
        elif t == "tcla":
            opt.heads = {"class": 1}
        elif t in ("simsiam", "simsiam2d3d", "simsiam3d", "scan", "scan2d3d"):
            opt.heads = {"proj": opt.head_conv, "pred": opt.head_conv}
        elif t == "moco":
            opt.heads = {"proj": 256, "pred": 256}
        elif t == "denoise":
            opt.heads = {"proj": 128}
        else:
            assert 0, "task not defined!"
        print("heads", opt.heads)
        return opt

    def init(self, args=""):
        opt = self.parse(args)
        res, ncls = _DATASETS[opt.task]
        ds = types.SimpleNamespace(default_resolution=res, num_classes=ncls, dataset=opt.task)
        opt.dataset = ds.dataset
        return self.update_dataset_info_and_set_heads(opt, ds)
