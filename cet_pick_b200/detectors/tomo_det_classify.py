"""Mirror of cet_pick/detectors/tomo_det_classify.py:18-229: the semiclass tile scheduler
(`PatchDataset`, `TomoClassdetDetector.process`) and its pick writer.  Tiles are cut and pasted on the
device; the greedy distance suppression (`tomo_decode_classify`) runs in csrc/greedy_nms.cu.

The reference cannot construct this task's networks through its own `create_model`
(models/model.py:65-70 passes kwargs the 'class'/'small' factories reject, SURVEY.md Surprise 3), so the
detector here also accepts a ready model: any module mapping (1,D,H,W) -> [{'hm': (1,1,D,H,W)}] at the
INPUT resolution, which is what the reference's paste logic (:134-140) requires."""
from __future__ import annotations

import os
import time

import numpy as np
import torch

from ..models.decode import tomo_decode_classify
from ..models.utils import _sigmoid
from ..utils.mrcio import write_mrc
from ..utils.post_process import graph_pick_lines
from .base_detector import BaseDetector


class PatchDataset:
    """tomo_det_classify.py:18-75: zero-padded (patch+2*padding) tiles on a regular grid.  `tomo` is a
    (nz,ny,nx) tensor (kept on its device); items are `(index int array (3,), tile tensor)`."""

    def __init__(self, tomo, patch_size_z=48, patch_size_xy=96, padding_z=12, padding_xy=24):
        self.tomo = tomo
        self.patch_size_xy = patch_size_xy
        self.patch_size_z = patch_size_z
        self.padding_xy = padding_xy
        self.padding_z = padding_z
        nz, ny, nx = tomo.shape
        pz = int(np.ceil(nz / patch_size_z))
        py = int(np.ceil(ny / patch_size_xy))
        px = int(np.ceil(nx / patch_size_xy))
        self.shape = (pz, py, px)
        self.num_patches = pz * py * px

    def __len__(self):
        return self.num_patches

    def __getitem__(self, patch):
        if patch < 0 or patch >= self.num_patches:
            raise IndexError(patch)
        i, j, k = np.unravel_index(patch, self.shape)
        psz, psxy, pdz, pdxy = self.patch_size_z, self.patch_size_xy, self.padding_z, self.padding_xy
        tomo = self.tomo
        i, j, k = psz * int(i), psxy * int(j), psxy * int(k)
        x = torch.zeros((psz + 2 * pdz, psxy + 2 * pdxy, psxy + 2 * pdxy), dtype=torch.float32, device=tomo.device)
        si, ei = max(0, i - pdz), min(tomo.shape[0], i + psz + pdz)
        sj, ej = max(0, j - pdxy), min(tomo.shape[1], j + psxy + pdxy)
        sk, ek = max(0, k - pdxy), min(tomo.shape[2], k + psxy + pdxy)
        sic, sjc, skc = pdz - i + si, pdxy - j + sj, pdxy - k + sk
        x[sic:sic + (ei - si), sjc:sjc + (ej - sj), skc:skc + (ek - sk)] = tomo[si:ei, sj:ej, sk:ek]
        return np.array((i, j, k), dtype=int), x


class TomoClassdetDetector(BaseDetector):
    def __init__(self, opt, model=None):
        if model is None:
            super(TomoClassdetDetector, self).__init__(opt)
        else:                      # see the module docstring
            if opt.gpus[0] < 0:
                raise RuntimeError("cet_pick_b200 has no CPU path: --gpus -1 is not supported (sm_100a only)")
            opt.device = torch.device("cuda")
            self.model = model.to(opt.device).eval()
            self.max_per_image = 900
            self.pause = True
        self.opt = opt

    @staticmethod
    def _zero_borders(out_hm):
        """tomo_det_classify.py:107-110,142-145: 30 voxels at the y and x borders of (1,D,H,W)."""
        out_hm[:, :, :30, :] = 0
        out_hm[:, :, -30:, :] = 0
        out_hm[:, :, :, :30] = 0
        out_hm[:, :, :, -30:] = 0

    def process(self, images, return_time=False):
        # :85-92 (the size test reads shape[0..2] of the (1,D,H,W) batch, i.e. batch, depth, height)
        if images.shape[0] <= 85 and images.shape[1] <= 128 and images.shape[2] <= 128:
            patch_size_z = patch_size_xy = 0
        else:
            patch_size_z, patch_size_xy, padding_z, padding_xy = 32, 96, 16, 24
        with torch.no_grad():
            if patch_size_z == 0:
                hm = self.model(images)[-1]["hm"]
                torch.cuda.synchronize()
                forward_time = time.time()
                hm = _sigmoid(hm)
                out_hm = hm[0]
            else:
                out_hm = torch.zeros_like(images, device=self.opt.device)
                patch_data = PatchDataset(images[0], patch_size_z, patch_size_xy, padding_z, padding_xy)
                forward_time = time.time()
                for n in range(len(patch_data)):
                    (i, j, k), x = patch_data[n]
                    xb = _sigmoid(self.model(x[None])[-1]["hm"])[0][0]
                    patch = out_hm[0, i:i + patch_size_z, j:j + patch_size_xy, k:k + patch_size_xy]
                    pz, py, px = patch.shape
                    out_hm[0, i:i + patch_size_z, j:j + patch_size_xy, k:k + patch_size_xy] = \
                        xb[padding_z:padding_z + pz, padding_xy:padding_xy + py, padding_xy:padding_xy + px]
                torch.cuda.synchronize()
            self._zero_borders(out_hm)
            detections = tomo_decode_classify(out_hm, self.opt.nms, self.opt.out_thresh)
            out_hm = out_hm.unsqueeze(0)
        output = None
        if return_time:
            return output, detections, out_hm, forward_time
        return output, detections, out_hm

    def post_process(self, dets, meta, scale=1, z_dim_tot=128):
        dets[:, :2] *= self.opt.down_ratio          # :164
        return dets, meta["name"][0]

    def save_detection(self, hm, dets, path, meta, prefix="", name=""):
        """:173-214: heat-map MRC, then plain / --with_score lines, or the --fiber / --spike graph post-processing of
        the kept picks (utils/post_process.py; the fiber re-sampling step is the reference's default of 2 here)."""
        os.makedirs(path, exist_ok=True)            # every rank of a torchrun job writes into the same directory
        hm = hm.detach().cpu().numpy()[0][0]
        max_z, max_y, max_x = hm.shape
        if np.isnan(hm).any():
            raise ValueError("Output contains NaN values")
        write_mrc(os.path.join(path, "{}_hm.mrc".format(name)), np.float32(np.swapaxes(hm, 1, 0)))
        o = self.opt
        a = np.asarray(dets, dtype=np.float64).reshape(-1, np.asarray(dets).shape[-1] if len(dets) else 4)
        x, y, z = (np.floor(a[:, j]).astype(np.int64) for j in range(3))      # rows [x, y, z, score] at the INPUT resolution
        score = a[:, 3]                                                       # the float32 values, exactly (:191)
        keep = (score > o.out_thresh) & (z >= o.cutoff_z) & (z <= max_z - o.cutoff_z) & (x > 20) & (x < max_x - 20) \
            & (y > 20) & (y < max_y - 20)                                     # no x2 here: the map is full resolution (:180)
        if o.compress:
            z = z * 2
        xs, ys, zs, sc = x[keep].tolist(), y[keep].tolist(), z[keep].tolist(), score[keep].tolist()
        if getattr(o, "fiber", False) or getattr(o, "spike", False):
            lines = graph_pick_lines(o, xs, ys, zs, sc, scale=2)
        elif not o.with_score:
            lines = ["%d\t%d\t%d" % t for t in zip(xs, zs, ys)]
        else:
            lines = ["%d\t%d\t%d\t%s" % (xx, zz, yy, str(s)) for xx, zz, yy, s in zip(xs, zs, ys, sc)]
        with open(os.path.join(path, "{}.txt".format(name)), "w+") as out_detect:
            out_detect.write("".join(ln + "\n" for ln in lines))

    def debug(self, debugger, images, dets, output, scale=1):
        pass
