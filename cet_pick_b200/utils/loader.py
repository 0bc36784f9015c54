"""Mirror of the pre-processing functions of cet_pick/utils/loader.py (quantize :16-25, load_rec :27-88,
preprocess :90-121, load_tomos_from_list :165-173) for reconstructed tomograms (`is_tilt=False`, what the
refinement step uses) and for tilt series (`is_tilt=True`: per-slice statistics).  Same names, arguments and float64 arithmetic; the volume lives on the GPU and every step
is a kernel of csrc/preproc.cu, so the results are CUDA tensors (float64, like the reference's numpy arrays;
`load_tomos_from_list(..., dtype=torch.float32)` gives the detector's input type directly, which is what the
reference's datasets produce with `.astype(np.float32)`, datasets/particle_moco.py:176-178)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib
from .mrcio import read_mrc

_DTYPE_CODE = {torch.float32: 0, torch.int16: 1, torch.int8: 3, torch.float64: 4}


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("cet_pick_b200.utils.loader: no CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_device(a):
    """numpy / tensor -> contiguous CUDA tensor of a dtype the gather kernel reads, plus its dtype code"""
    if isinstance(a, np.ndarray):
        if a.dtype == np.uint16:          # torch has no uint16 arithmetic: ship the bits as int16, code 2 reads them unsigned
            t = torch.from_numpy(np.array(a, copy=True).view(np.int16)).to(_dev())
            return t, 2
        a = np.ascontiguousarray(a)
        a = torch.from_numpy(a if a.flags.writeable else a.copy())      # memory-mapped / frombuffer arrays are read-only
    t = a if a.is_cuda else a.to(_dev())
    if t.dtype not in _DTYPE_CODE:
        t = t.to(torch.float64)
    return t.contiguous(), _DTYPE_CODE[t.dtype]


def _stats(x: torch.Tensor) -> torch.Tensor:
    L = _lib.lib()
    nb = C.c_size_t(0)
    _lib.check(L.cetpick_pre_stats_workspace_bytes(C.byref(nb)), "pre_stats_workspace_bytes")
    ws = torch.empty(nb.value, dtype=torch.uint8, device=x.device)
    st = torch.empty(2, dtype=torch.float64, device=x.device)
    _lib.check(L.cetpick_pre_mean_std_f64(x.data_ptr(), x.numel(), st.data_ptr(), ws.data_ptr(), ws.numel(),
                                          _lib.stream_ptr()), "pre_mean_std")
    return st


def _zscore_(x: torch.Tensor) -> torch.Tensor:
    st = _stats(x)
    _lib.check(_lib.lib().cetpick_pre_zscore_f64(x.data_ptr(), x.numel(), st.data_ptr(), _lib.stream_ptr()), "pre_zscore")
    return x


def quantize(x, mi=-2.5, ma=2, dtype=torch.uint8):
    """loader.py:16-25 on a float64 CUDA tensor -> uint8 levels."""
    if dtype not in (torch.uint8, np.uint8):
        raise NotImplementedError("quantize: only uint8 levels (the reference's default) are implemented")
    x, _ = _to_device(x)
    x = x.to(torch.float64)
    if mi is None:
        mi = float(x.min().item())
    if ma is None:
        ma = float(x.max().item())
    q = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    _lib.check(_lib.lib().cetpick_pre_quantize_u8(x.data_ptr(), x.numel(), float(mi), float(ma), q.data_ptr(),
                                                  _lib.stream_ptr()), "pre_quantize")
    return q


def load_rec(path, order="xyz", compress=False, is_tilt=False):
    """loader.py:27-88.  `path` may also be the (nz,ny,nx) array `mrcfile` would return.  -> float64 CUDA tensor.
    is_tilt: every output slice is z-scored on its own (:48-49,56-57) instead of the volume as a whole (:59-60)."""
    rec = read_mrc(path) if isinstance(path, (str, bytes)) or hasattr(path, "__fspath__") else path
    t, code = _to_device(rec)
    if t.dim() != 3:
        raise ValueError(f"load_rec: expected a 3-D volume, got {tuple(t.shape)}")
    n0, n1, n2 = t.shape
    s0, s1, s2 = n1 * n2, n2, 1                       # element strides of the stored array
    if order in ("xzy", "xyz", "yxz"):
        shape, st = [n0, n1, n2], [s0, s1, s2]
        if order == "xzy":                            # np.swapaxes(rec, 2, 1)
            shape[1], shape[2], st[1], st[2] = shape[2], shape[1], st[2], st[1]
        if order == "yxz":                            # np.swapaxes(rec, 1, 0)
            shape[0], shape[1], st[0], st[1] = shape[1], shape[0], st[1], st[0]
        A, B, Z = shape                               # x, y, z = rec.shape; slices rec[:, :, i]
        sa, sb, sz = st
        J = -(-Z // 2) if compress else Z             # math.ceil(z / 2)
    elif order == "zxy":
        Z, A, B = n0, n1, n2                          # z, x, y = rec.shape; slices rec[i]
        sz, sa, sb = s0, s1, s2
        J = Z // 2 if compress else Z
        if compress and Z % 2:
            raise IndexError("load_rec(order='zxy', compress=True) needs an even number of slices "
                             "(the reference overruns its int(z//2)-slice buffer, loader.py:64-78)")
    else:
        raise UnboundLocalError(f"load_rec: unknown order {order!r} (the reference leaves new_slices unbound)")
    out = torch.empty((J, A, B), dtype=torch.float64, device=t.device)
    _lib.check(_lib.lib().cetpick_pre_gather_f64(t.data_ptr(), code, A, B, J, sa, sb, sz, Z, 1 if compress else 0,
                                                 out.data_ptr(), _lib.stream_ptr()), "pre_gather")
    if is_tilt:
        for j in range(J):                            # (new_slice - new_slice.mean()) / new_slice.std()
            _zscore_(out[j])
        return out
    return _zscore_(out)                              # (new_slices - mean) / std, loader.py:59-60,85-86


def gaussian_kernel1d(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) with radius = int(truncate*sigma + 0.5)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return phi / phi.sum(), radius


def gaussian_filter(x: torch.Tensor, sigma) -> torch.Tensor:
    """scipy.ndimage.gaussian_filter(x, sigma) (mode='reflect', truncate=4) of a float64 (n0,n1,n2) CUDA tensor."""
    if float(sigma) <= 1e-15:
        return x.clone()
    w, radius = gaussian_kernel1d(sigma)
    w = np.ascontiguousarray(w, dtype=np.float64)
    a, b = x, torch.empty_like(x)
    if x.dim() == 2:                                  # a single image: filter its two axes
        n0, (n1, n2), axes = 1, x.shape, (1, 2)
    else:
        (n0, n1, n2), axes = x.shape, (0, 1, 2)
    for axis in axes:
        _lib.check(_lib.lib().cetpick_pre_gauss1d_f64(a.data_ptr(), b.data_ptr(), n0, n1, n2, axis,
                                                      w.ctypes.data_as(C.c_void_p), radius, _lib.stream_ptr()),
                   "pre_gauss1d")
        a, b = b, (torch.empty_like(x) if a is x else a)      # never overwrite the caller's tensor
    return a


def preprocess(mrc, denoise=0, is_tilt=False, dtype=torch.float64):
    """loader.py:90-121 (reconstruction branch): [Gaussian sigma=denoise] -> z-score -> 256 levels on [-3,3]
    (or [-2.5,2] without denoising) -> min-max to [0,1].  -> CUDA tensor of `dtype` (float64 like the reference)."""
    x, _ = _to_device(mrc)
    x = x.to(torch.float64)
    if x.dim() != 3:
        raise ValueError(f"preprocess: expected a 3-D volume, got {tuple(x.shape)}")
    if is_tilt:
        # :92-100,108-116: per slice [2-D Gaussian] -> z-score -> quantize(-2.5, 2) -> cv2.normalize(NORM_MINMAX, CV_32F)
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        mm = torch.empty(2, dtype=torch.int32, device=x.device)
        for j in range(x.shape[0]):
            dd = gaussian_filter(x[j], denoise) if denoise > 0 else x[j].clone()
            _zscore_(dd)
            q = quantize(dd)
            _lib.check(_lib.lib().cetpick_pre_minmax_normalize(q.data_ptr(), q.numel(), mm.data_ptr(), out[j].data_ptr(),
                                                               0, _lib.stream_ptr()), "pre_minmax_normalize")
        return out                                    # float32 like np.asarray of the CV_32F slices
    if denoise > 0:
        im = gaussian_filter(x, denoise)
        mi, ma = -3.0, 3.0
    else:
        im = x.clone()
        mi, ma = -2.5, 2.0
    _zscore_(im)
    q = quantize(im, mi=mi, ma=ma)
    out = torch.empty(q.shape, dtype=dtype, device=q.device)
    mm = torch.empty(2, dtype=torch.int32, device=q.device)
    if dtype not in (torch.float64, torch.float32):
        raise ValueError("preprocess: dtype must be float64 or float32")
    _lib.check(_lib.lib().cetpick_pre_minmax_normalize(q.data_ptr(), q.numel(), mm.data_ptr(), out.data_ptr(),
                                                       1 if dtype == torch.float64 else 0, _lib.stream_ptr()),
               "pre_minmax_normalize")
    return out


def preprocess_levels(mrc, denoise=0):
    """`preprocess` (reconstruction branch) stopped one step early: -> (levels, level_values) with `levels` a uint8
    CUDA tensor (q - q.min(), loader.py:106,120) and `level_values[k]` the float32 value the reference's float32
    input holds for level k, float32((q - min) / (max - min)).  `level_values[levels]` == preprocess(...) cast to
    float32 (datasets cast with astype(np.float32)); the detector takes the pair as is (TomoConvUNet.level_values)."""
    x, _ = _to_device(mrc)
    x = x.to(torch.float64)
    if x.dim() != 3:
        raise ValueError(f"preprocess_levels: expected a 3-D volume, got {tuple(x.shape)}")
    if denoise > 0:
        im, mi, ma = gaussian_filter(x, denoise), -3.0, 3.0
    else:
        im, mi, ma = x.clone(), -2.5, 2.0
    _zscore_(im)
    q = quantize(im, mi=mi, ma=ma)
    lo, hi = (int(v) for v in torch.aminmax(q))
    if hi == lo:
        raise ZeroDivisionError("preprocess_levels: constant volume (the reference divides by max - min = 0)")
    lv = np.zeros(256, dtype=np.float32)
    lv[:hi - lo + 1] = (np.arange(hi - lo + 1, dtype=np.float64) / np.float64(hi - lo)).astype(np.float32)
    return q - lo, lv


def load_tomos_from_list(names, paths, order="xzy", compress=False, denoise=0, tilt=False, dtype=torch.float64):
    """loader.py:165-173."""
    images = {}
    for name, path in zip(names, paths):
        im = load_rec(path, order=order, compress=compress, is_tilt=tilt)
        images[name] = preprocess(im, denoise=denoise, is_tilt=tilt, dtype=dtype)
    return images


def load_tomos_from_list_nopre(names, paths, order="xzy", compress=False, denoise=False, tilt=False):
    """loader.py:175-180."""
    return {name: load_rec(path, order=order, compress=compress, is_tilt=tilt) for name, path in zip(names, paths)}
