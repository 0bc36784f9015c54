"""Detector driver with the interface of cet_pick/detectors/base_detector.py:15-106 (`BaseDetector`):
construction = build the model + load the checkpoint; `run(images, meta)` = device copy -> `process` ->
`post_process` -> `save_detection`, returning the reference's timing dictionary
`{'tot_time','load','pre','net','dec'}` (seconds; 'pre' is always 0 there too)."""
from __future__ import annotations

import time

import torch

from ..models.model import create_model, load_model


class _Laps:
    """wall-clock laps in seconds (the reference brackets its stages with time.time())"""

    def __init__(self):
        self.t0 = self.last = time.time()

    def lap(self, now=None):
        now = time.time() if now is None else now
        dt, self.last = now - self.last, now
        return dt

    def total(self):
        return time.time() - self.t0


class BaseDetector(object):
    def __init__(self, opt):
        if opt.gpus[0] < 0:
            raise RuntimeError("cet_pick_b200 has no CPU path: --gpus -1 is not supported (sm_100a only)")
        opt.device = torch.device("cuda")
        print("Creating model...")
        net = create_model(opt.arch, opt.heads, opt.head_conv, last_k=opt.last_k)
        net = load_model(net, opt.load_model)            # a checkpoint is always loaded, like the reference (:24)
        self.model = net.to(opt.device).eval()
        if hasattr(self.model, "precision"):
            self.model.precision = getattr(opt, "precision", "bf16")
        self.max_per_image = 900
        self.opt = opt
        self.pause = True

    # ---- hooks of the concrete detectors (tomo_det.py, tomo_det_classify.py) ----
    def process(self, images, return_time=False):
        raise NotImplementedError

    def post_process(self, dets, meta, scale=1):
        raise NotImplementedError

    def merge_outputs(self, detections):
        raise NotImplementedError

    def debug(self, debugger, images, dets, output, scale=1):
        raise NotImplementedError

    def show_results(self, debugger, image, results):
        raise NotImplementedError

    def save_detection(self, dets, path, meta, prefix="", name=""):
        raise NotImplementedError

    def run(self, image_or_path_or_tensor, meta=None):
        """One tomogram through the whole chain; stage times as in base_detector.py:62-106."""
        clock = _Laps()
        stats = {"load": clock.lap(), "pre": 0}
        volume = image_or_path_or_tensor.to(self.opt.device, non_blocking=True)
        # a uint8 volume holds the quantised levels of utils/loader.py:preprocess_levels; meta carries their values
        self.model.level_values = meta.get("level_values") if isinstance(meta, dict) else None
        clock.lap()                                      # the copy is not attributed to a stage in the reference either
        output, dets, hm, t_forward = self.process(volume, return_time=True)
        stats["net"] = clock.lap(t_forward)              # process() synchronises before taking t_forward
        stats["dec"] = clock.lap()
        if self.opt.debug >= 2:
            self.debug(None, volume, dets, output)
        depth = hm.size(2)                               # hm is (batch, cat, depth, height, width)
        dets, name = self.post_process(dets, meta, z_dim_tot=depth)
        torch.cuda.synchronize()
        self.save_detection(hm, dets, self.opt.out_path, meta, name=name)
        stats["tot_time"] = clock.total()
        return {k: stats[k] for k in ("tot_time", "load", "pre", "net", "dec")}
