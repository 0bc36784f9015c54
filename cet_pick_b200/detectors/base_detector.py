"""Mirror of cet_pick/detectors/base_detector.py:15-106 (BaseDetector): model build + checkpoint
load, and the run() driver returning the same timing dict."""
from __future__ import annotations

import time

import torch

from ..models.model import create_model, load_model


class BaseDetector(object):
    def __init__(self, opt):
        if opt.gpus[0] < 0:
            raise RuntimeError("cet_pick_b200 has no CPU path: --gpus -1 is not supported (sm_100a only)")
        opt.device = torch.device("cuda")
        print("Creating model...")
        self.model = create_model(opt.arch, opt.heads, opt.head_conv, last_k=opt.last_k)
        self.model = load_model(self.model, opt.load_model)      # always loads, like the reference (:24)
        self.model = self.model.to(opt.device)
        self.model.eval()
        self.max_per_image = 900
        self.opt = opt
        self.pause = True

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
        """base_detector.py:62-106.  H2D copy -> process -> post_process -> save_detection."""
        load_time = pre_time = net_time = dec_time = post_time = tot_time = 0
        start_time = time.time()
        loaded_time = time.time()
        load_time += loaded_time - start_time
        images = image_or_path_or_tensor.to(self.opt.device, non_blocking=True)
        pre_process_time = time.time()
        output, dets, hm, forward_time = self.process(images, return_time=True)
        batch, cat, depth, height, width = hm.size()
        net_time += forward_time - pre_process_time
        decode_time = time.time()
        dec_time += decode_time - forward_time
        if self.opt.debug >= 2:
            self.debug(None, images, dets, output)
        dets, name = self.post_process(dets, meta, z_dim_tot=depth)
        torch.cuda.synchronize()
        post_time += time.time() - decode_time
        self.save_detection(hm, dets, self.opt.out_path, meta, name=name)
        tot_time += time.time() - start_time
        return {"tot_time": tot_time, "load": load_time, "pre": pre_time, "net": net_time, "dec": dec_time}
