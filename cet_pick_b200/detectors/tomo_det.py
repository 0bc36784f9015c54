"""Mirror of cet_pick/detectors/tomo_det.py:18-95 (TomodetDetector): whole-tomogram forward,
_sigmoid, tomo_decode, grouping by z and the `<name>.txt` / `<name>_hm.mrc` writers."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from ..models.decode import decode_status, tomo_decode
from ..utils.mrcio import write_mrc
from ..utils.post_process import graph_pick_lines, tomo_post_process
from .base_detector import BaseDetector


class TomodetDetector(BaseDetector):
    def __init__(self, opt):
        super(TomodetDetector, self).__init__(opt)
        # the detector never reads 'proj' (tomo_det.py:26-27): skip its 32-channel fp32 map unless asked
        self.model.compute_proj = bool(getattr(opt, "return_proj", False))
        # `_sigmoid(hm)` mutates output['hm'] in place in the reference (:33); fusing it into the hm
        # epilogue leaves the same tensor in both places
        self.model.fuse_sigmoid = True
        # opt.async_write (or set_async_write): finished heat-maps go to <name>_hm.mrc from writer threads while the GPU
        # works on the next tomogram; run() then returns before the files exist and flush() waits for them
        self._writer = None
        self._pool = None
        if getattr(opt, "async_write", False):
            self.set_async_write(True)

    def set_async_write(self, on=True, threads=3):
        from ..utils.async_io import AsyncWriter, PinnedPool
        self.flush()
        if self._writer is not None:
            self._writer.close()
        self._writer = AsyncWriter(threads, max_pending=threads + 1) if on else None
        self._pool = PinnedPool(threads + 2) if on else None

    def flush(self):
        """wait until every pick file / heat-map of the tomograms run() has returned for is on disk"""
        if getattr(self, "_writer", None) is not None:
            self._writer.flush()

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
            max_y, max_z, max_x = sw.shape
            if getattr(self, "_writer", None) is not None:
                # asynchronous path: D2H into a pooled page-locked buffer, the file write happens on a writer thread
                host = self._pool.get(sw.shape, sw.dtype)
                st_dev = torch.stack([mn.double(), mx.double(), mean, sd])
                st_host = torch.empty(4, dtype=torch.float64, pin_memory=True)
                if getattr(self, "_copy_stream", None) is None:
                    self._copy_stream = torch.cuda.Stream(hm.device)
                cs = self._copy_stream                          # the copy engine works under the next tomogram's kernels
                cs.wait_stream(torch.cuda.current_stream(hm.device))
                with torch.cuda.stream(cs):
                    host.copy_(sw, non_blocking=True)
                    st_host.copy_(st_dev, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(cs)
                sw.record_stream(cs)
                st_dev.record_stream(cs)
                lines = self._pick_lines(dets, max_z, max_x * 2, max_y * 2)
                mrc_path, txt_path = os.path.join(path, "{}_hm.mrc".format(name)), os.path.join(path, "{}.txt".format(name))
                pool = self._pool

                def job():
                    done.synchronize()
                    try:
                        write_mrc(mrc_path, host.numpy(), tuple(float(v) for v in st_host))
                    finally:
                        pool.put(host)
                    with open(txt_path, "w+") as f:
                        f.write("".join(ln + "\n" for ln in lines))

                self._writer.submit(job)
                return
            host = self._pinned_like(sw)
            host.copy_(sw, non_blocking=True)
            st = torch.stack([mn.double(), mx.double(), mean, sd]).cpu()   # synchronises: `host` is complete
            stats = tuple(float(v) for v in st)
            hm = host.numpy()
        else:
            hm = hm.numpy()[0][0]
            max_z, max_y, max_x = hm.shape
            hm = np.swapaxes(hm, 1, 0)
            if np.isnan(hm).any():
                raise ValueError("Output contains NaN values")
        max_x, max_y = max_x * 2, max_y * 2
        write_mrc(os.path.join(path, "{}_hm.mrc".format(name)), hm, stats)
        lines = self._pick_lines(dets, max_z, max_x, max_y)
        with open(os.path.join(path, "{}.txt".format(name)), "w+") as f:
            for ln in lines:
                print(ln, file=f)

    def _pick_lines(self, dets, max_z, max_x, max_y):
        """tomo_det.py:69-95: one `x\tz\ty[\tscore]` line per pick that passes the score / z-cutoff / 20-px border filter;
        with --fiber / --spike the kept picks go through the graph post-processing of utils/post_process.py instead.
        --spike without --fiber follows TomoClassdetDetector.save_detection (tomo_det_classify.py:196-214): here the
        reference fills its candidate list only under --fiber and then indexes the empty list (IndexError)."""
        o = self.opt
        from itertools import chain
        rows = list(chain.from_iterable(dets.values()))           # dict order = ascending z, rows in top-K order
        a = np.asarray(rows, dtype=np.float64).reshape(-1, 5)     # the float32 values, exactly (float(c[3]) in the reference)
        x, y, z = (np.floor(a[:, j]).astype(np.int64) for j in range(3))
        score = a[:, 3]
        keep = (score > o.out_thresh) & (z >= o.cutoff_z) & (z <= max_z - o.cutoff_z) & (x > 20) & (x < max_x - 20) \
            & (y > 20) & (y < max_y - 20)
        if o.compress:
            z = z * 2
        xs, ys, zs, sc = x[keep].tolist(), y[keep].tolist(), z[keep].tolist(), score[keep].tolist()
        if o.fiber or o.spike:
            return graph_pick_lines(o, xs, ys, zs, sc, scale=o.distance_scale)
        if not o.with_score:
            return ["%d\t%d\t%d" % t for t in zip(xs, zs, ys)]
        return ["%d\t%d\t%d\t%s" % (xx, zz, yy, str(s)) for xx, zz, yy, s in zip(xs, zs, ys, sc)]

    def _pinned_like(self, t):
        """page-locked staging buffer for the heat-map copy, kept across tomograms"""
        buf = getattr(self, "_hm_host", None)
        if buf is None or buf.shape != t.shape:
            buf = self._hm_host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        return buf

    def debug(self, debugger, images, dets, output, scale=1):
        """tomo_det.py:107-108: a no-op in the reference as well (run() calls it for --debug >= 2, the default)."""
        pass
