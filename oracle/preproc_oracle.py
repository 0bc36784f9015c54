"""TEST INFRASTRUCTURE (CPU restatement, numpy + scipy.ndimage): the pre-processing of
cet_pick/utils/loader.py (quantize :16-25, load_rec :27-88, preprocess :90-121) for reconstructions
(is_tilt=False).  Only tests/ may import this.  Pinned by tests/golden/preproc_*.npz, produced by the unmodified
reference (tests/golden/make_golden_pre.py)."""
from __future__ import annotations

import math

import numpy as np
from scipy.ndimage import gaussian_filter


def quantize(x, mi=-2.5, ma=2, dtype=np.uint8):
    """loader.py:16-25."""
    if mi is None:
        mi = x.min()
    if ma is None:
        ma = x.max()
    r = ma - mi
    x = 255 * (x - mi) / r
    x = np.clip(x, 0, 255)
    return np.round(x).astype(dtype)


def load_rec(rec: np.ndarray, order="xyz", compress=False):
    """loader.py:27-88 on the array mrcfile would return; float64 (z', x, y)."""
    if order in ("xzy", "xyz", "yxz"):
        if order == "xzy":
            rec = np.swapaxes(rec, 2, 1)
        if order == "yxz":
            rec = np.swapaxes(rec, 1, 0)
        x, y, z = rec.shape
        if compress:
            out = np.zeros([math.ceil(z / 2), x, y])
            for j, i in enumerate(range(0, z, 2)):
                out[j] = np.max(rec[:, :, i:i + 2], axis=-1)
        else:
            out = np.zeros([z, x, y])
            for i in range(z):
                out[i] = rec[:, :, i]
    elif order == "zxy":
        z, x, y = rec.shape
        if compress:
            out = np.zeros([z // 2, x, y])
            for j, i in enumerate(range(0, z, 2)):
                out[j] = np.max(rec[i:i + 2], axis=0)
        else:
            out = np.zeros([z, x, y])
            for i in range(z):
                out[i] = rec[i]
    else:
        raise UnboundLocalError(order)
    return (out - np.mean(out)) / np.std(out)


def preprocess(mrc: np.ndarray, denoise=0):
    """loader.py:90-121, reconstruction branch."""
    if denoise > 0:
        im = gaussian_filter(mrc, sigma=denoise)
        im = (im - im.mean()) / im.std()
        im = quantize(im, mi=-3, ma=3)
    else:
        im = (mrc - mrc.mean()) / mrc.std()
        im = quantize(im)
    return (im - np.min(im)) / (np.max(im) - np.min(im))
