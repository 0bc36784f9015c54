"""Mirror of cet_pick/detectors/tomo_det.py:18-95 (TomodetDetector): whole-tomogram forward,
_sigmoid, tomo_decode, grouping by z and the `<name>.txt` / `<name>_hm.mrc` writers."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from ..models.decode import decode_status, tomo_decode
from ..utils.mrcio import write_mrc
from ..utils.post_process import tomo_post_process
from .base_detector import BaseDetector


class TomodetDetector(BaseDetector):
    def __init__(self, opt):
        super(TomodetDetector, self).__init__(opt)
        # the detector never reads 'proj' (tomo_det.py:26-27): skip its 32-channel fp32 map unless asked
        self.model.compute_proj = bool(getattr(opt, "return_proj", False))
        # `_sigmoid(hm)` mutates output['hm'] in place in the reference (:33); fusing it into the hm
        # epilogue leaves the same tensor in both places
        self.model.fuse_sigmoid = True

    def process(self, images, return_time=False):
        with torch.no_grad():
            output = self.model(images)[-1]
            hm = output["hm"]                         # already sigmoid + clamp (fused epilogue)
            torch.cuda.synchronize()
            forward_time = time.time()
            dets = tomo_decode(hm, kernel=self.opt.nms, reg=None, K=self.opt.K, if_fiber=self.opt.fiber)
            # remember which heat-map the decode saw: save_detection then takes the NaN verdict from the decode's
            # device-side flag (cetpick_decode_status bit 0) instead of scanning 268 MB on the host
            self._decoded = (hm.data_ptr(), hm._version, tuple(hm.shape))
        if return_time:
            return output, dets, hm, forward_time
        return output, dets, hm

    def post_process(self, dets, meta, scale=1, z_dim_tot=128):
        dets = dets.detach().cpu().numpy().reshape(1, -1, dets.shape[2])
        dets[:, :, :2] *= self.opt.down_ratio
        preds = tomo_post_process(dets, z_dim_tot=z_dim_tot)[0]
        return preds, meta["name"][0]

    def save_detection(self, hm, dets, path, meta, prefix="", name=""):
        """tomo_det.py:53-95: heat-map MRC (axes swapped to (H', D, W')), then one line per pick
        `x\\tz\\ty[\\tscore]`, ordered by z then top-K order, filtered by score / z cutoff / 20-px border."""
        os.makedirs(path, exist_ok=True)            # every rank of a torchrun job writes into the same directory
        hm = hm.detach()
        stats = None
        if hm.is_cuda:
            # NaN check (:64-65), axis swap (:58-60) and the MRC header statistics on the device; one pinned copy out
            if getattr(self, "_decoded", None) == (hm.data_ptr(), hm._version, tuple(hm.shape)) and hm.shape[0] == 1 \
                    and not (self.opt.fiber and self.opt.nms != 3):
                flags, _ = decode_status(hm.device)
                has_nan = bool(flags & 1)
            else:
                has_nan = bool(torch.isnan(hm).any())
            if has_nan:
                raise ValueError("Output contains NaN values")
            vol = hm[0, 0]
            sw = vol.permute(1, 0, 2).contiguous()             # np.swapaxes(hm, 1, 0): (H', D, W')
            sd, mean = torch.std_mean(vol.double(), correction=0)
            mn, mx = torch.aminmax(vol)
            host = self._pinned_like(sw)
            host.copy_(sw, non_blocking=True)
            st = torch.stack([mn.double(), mx.double(), mean, sd]).cpu()   # synchronises: `host` is complete
            stats = tuple(float(v) for v in st)
            hm = host.numpy()
            max_y, max_z, max_x = hm.shape
        else:
            hm = hm.numpy()[0][0]
            max_z, max_y, max_x = hm.shape
            hm = np.swapaxes(hm, 1, 0)
            if np.isnan(hm).any():
                raise ValueError("Output contains NaN values")
        max_x, max_y = max_x * 2, max_y * 2
        write_mrc(os.path.join(path, "{}_hm.mrc".format(name)), hm, stats)
        o = self.opt
        if o.fiber or o.spike:
            raise NotImplementedError("fiber/spike graph post-processing is outside the hot path "
                                      "(utils/post_process.py:31-106; DESIGN.md)")
        lines = []
        for k, v in dets.items():
            for c in v:
                x, y, z, score = int(np.floor(c[0])), int(np.floor(c[1])), int(np.floor(c[2])), float(c[3])
                if (score > o.out_thresh and z >= o.cutoff_z and z <= max_z - o.cutoff_z
                        and 20 < x < max_x - 20 and 20 < y < max_y - 20):
                    if o.compress:
                        z = int(z) * 2
                    if not o.with_score:
                        lines.append(str(x) + "\t" + str(z) + "\t" + str(y))
                    else:
                        lines.append(str(x) + "\t" + str(z) + "\t" + str(y) + "\t" + str(score))
        with open(os.path.join(path, "{}.txt".format(name)), "w+") as f:
            for ln in lines:
                print(ln, file=f)

    def _pinned_like(self, t):
        """page-locked staging buffer for the heat-map copy, kept across tomograms"""
        buf = getattr(self, "_hm_host", None)
        if buf is None or buf.shape != t.shape:
            buf = self._hm_host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        return buf

    def debug(self, debugger, images, dets, output, scale=1):
        """tomo_det.py:107-108: a no-op in the reference as well (run() calls it for --debug >= 2, the default)."""
        pass
