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


def load_rec(rec: np.ndarray, order="xyz", compress=False, is_tilt=False):
    """loader.py:27-88 on the array mrcfile would return; float64 (z', x, y).  is_tilt: every slice is z-scored on
    its own IN THE DTYPE OF THE FILE (numpy float32 statistics for a mode-2 MRC) before it lands in the float64
    buffer (:48-50,56-58); otherwise the float64 volume is z-scored as a whole (:59-60)."""
    def norm(sl):
        return (sl - sl.mean()) / sl.std() if is_tilt else sl

    if order in ("xzy", "xyz", "yxz"):
        if order == "xzy":
            rec = np.swapaxes(rec, 2, 1)
        if order == "yxz":
            rec = np.swapaxes(rec, 1, 0)
        x, y, z = rec.shape
        if compress:
            out = np.zeros([math.ceil(z / 2), x, y])
            for j, i in enumerate(range(0, z, 2)):
                out[j] = norm(np.max(rec[:, :, i:i + 2], axis=-1))
        else:
            out = np.zeros([z, x, y])
            for i in range(z):
                out[i] = norm(rec[:, :, i])
    elif order == "zxy":
        z, x, y = rec.shape
        if compress:
            out = np.zeros([z // 2, x, y])
            for j, i in enumerate(range(0, z, 2)):
                out[j] = norm(np.max(rec[i:i + 2], axis=0))
        else:
            out = np.zeros([z, x, y])
            for i in range(z):
                out[i] = norm(rec[i])
    else:
        raise UnboundLocalError(order)
    if is_tilt:
        return out
    return (out - np.mean(out)) / np.std(out)


def preprocess_tilt(mrc: np.ndarray, denoise=0):
    """loader.py:92-100,108-116, tilt branch; cv2.normalize(NORM_MINMAX, CV_32F) restated as float32 scale/shift."""
    out = []
    for sli in mrc:
        dd = gaussian_filter(sli, sigma=denoise) if denoise > 0 else sli
        dd = (dd - dd.mean()) / dd.std()
        q = quantize(dd)
        lo, hi = int(q.min()), int(q.max())
        scale = 1.0 / (hi - lo) if hi > lo else 0.0
        out.append((q.astype(np.float32) * np.float32(scale) + np.float32(-lo * scale)).astype(np.float32))
    return np.asarray(out)


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
