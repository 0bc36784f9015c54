"""TEST / BENCH INFRASTRUCTURE (not part of the product package): seeded synthetic inputs for the localisation hot
path (tomograms, heat-maps, checkpoints).

Everything is a pure function of (seed, element index) through one 64-bit counter hash
(splitmix64 finaliser), written once for numpy (host, tests, oracle) and once for torch
(device-side generation inside bench.py, so config-2's 64 GiB of tomograms never cross PCIe
unless the H2D copy is the thing being measured).  Both produce bit-identical values.

Shapes follow SURVEY.md section 8(d):
  * tomogram  T(D,H,W,seed)  : float32 k/255 levels, like cet_pick/utils/loader.py:117-120
  * heat-map  Hm(D,H,W,seed) : tie-free fp32 values in (0,1)  (decode parity / config 3)
  * weights   W1(seed)       : every tensor of the unet_N state_dict (Appendix A keys),
                               xavier-like scale, non-trivial BatchNorm statistics
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

_M64 = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15
_C1 = 0xBF58476D1CE4E5B9
_C2 = 0x94D049BB133111EB


# --------------------------------------------------------------------------- numpy
def mix64_np(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = x.astype(np.uint64) + np.uint64(_GOLD)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(_C1)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(_C2)
        x = x ^ (x >> np.uint64(31))
    return x


def _stream_np(seed: int, n: int, offset: int = 0) -> np.ndarray:
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return mix64_np(idx + np.uint64((seed * 0xD1342543DE82EF95) & _M64))


def uniform_np(seed: int, n: int, offset: int = 0) -> np.ndarray:
    """float32 uniforms in [0,1) with 24 random bits (exact in fp32)."""
    return ((_stream_np(seed, n, offset) >> np.uint64(40)).astype(np.float32)
            * np.float32(1.0 / (1 << 24)))


def tomogram_np(D: int, H: int, W: int, seed: int = 0) -> np.ndarray:
    """T(D,H,W,seed): (D,H,W) float32 in {k/255}.  Sum of four 6-bit fields (bell-shaped)."""
    bits = _stream_np(seed, D * H * W)
    k = ((bits >> np.uint64(8)) & np.uint64(63)) + ((bits >> np.uint64(20)) & np.uint64(63)) \
        + ((bits >> np.uint64(32)) & np.uint64(63)) + ((bits >> np.uint64(44)) & np.uint64(63))
    return (k.astype(np.float32) / np.float32(255.0)).reshape(D, H, W)


def heatmap_tiefree_np(D: int, H: int, W: int, seed: int = 0) -> np.ndarray:
    """Hm(i): all D*H*W values distinct, in (0,1): fp32 bit pattern 0x3F7FFFFF - perm(i).

    perm is a bijection on [0, 2^b), b = ceil(log2 N): odd multiplier + xorshift + odd multiplier.
    """
    n = D * H * W
    b = max(1, (n - 1).bit_length())
    assert b <= 29, "tie-free pattern needs <= 2^29 voxels"
    mask = np.uint64((1 << b) - 1)
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (i * np.uint64(0x9E3779B1 | 1) + np.uint64(seed * 2654435761 & _M64)) & mask
        x = x ^ (x >> np.uint64((b + 1) // 2))
        x = (x * np.uint64(0x85EBCA6B | 1)) & mask
        x = x ^ (x >> np.uint64((b + 1) // 2))
    bits = (np.uint64(0x3F7FFFFF) - x).astype(np.uint32)
    return bits.view(np.float32).reshape(D, H, W)


def heatmap_peaks_np(D: int, H: int, W: int, n_peaks: int, seed: int = 0,
                     sigma: float = 2.0) -> np.ndarray:
    """Hm(ii): sum of Gaussian peaks -> sigmoid -> clamp [1e-4, 1-1e-4]; big floor plateau."""
    u = uniform_np(seed, 4 * n_peaks).reshape(n_peaks, 4)
    logit = np.full((D, H, W), -12.0, np.float32)
    r = int(math.ceil(3 * sigma))
    for cz, cy, cx, a in u:
        z0, y0, x0 = int(cz * D), int(cy * H), int(cx * W)
        zs = slice(max(0, z0 - r), min(D, z0 + r + 1))
        ys = slice(max(0, y0 - r), min(H, y0 + r + 1))
        xs = slice(max(0, x0 - r), min(W, x0 + r + 1))
        zz, yy, xx = np.meshgrid(np.arange(zs.start, zs.stop), np.arange(ys.start, ys.stop),
                                 np.arange(xs.start, xs.stop), indexing="ij")
        g = np.exp(-((zz - z0) ** 2 + (yy - y0) ** 2 + (xx - x0) ** 2) / (2 * sigma * sigma))
        logit[zs, ys, xs] += (np.float32(8.0 + 10.0 * a) * g).astype(np.float32)
    hm = (1.0 / (1.0 + np.exp(-logit.astype(np.float64)))).astype(np.float32)
    return np.clip(hm, np.float32(1e-4), np.float32(1 - 1e-4))


# --------------------------------------------------------------------------- weights
def unet_param_shapes(n_blocks: int = 4, heads=None, head_conv: int = 32) -> "OrderedDict[str, tuple]":
    """state_dict keys/shapes of TomoConvUNet (reference unet_small.py:30-61, unet.py:807-848)."""
    heads = heads if heads is not None else {"hm": 1, "proj": 32}
    sh: "OrderedDict[str, tuple]" = OrderedDict()

    def bn(prefix, c):
        sh[prefix + ".weight"] = (c,)
        sh[prefix + ".bias"] = (c,)
        sh[prefix + ".running_mean"] = (c,)
        sh[prefix + ".running_var"] = (c,)
        sh[prefix + ".num_batches_tracked"] = ()

    sh["conv1.weight"] = (16, 1, 7, 7)
    bn("bn1", 16)
    outs = 16
    for i in range(n_blocks):
        ins = 16 if i == 0 else outs
        outs = 32 * (2 ** i)
        p = f"unet.down_convs.{i}"
        sh[p + ".conv1.weight"] = (outs, ins, 3, 3)
        sh[p + ".conv2.weight"] = (outs, outs, 3, 3)
        bn(p + ".norm0", outs)
        bn(p + ".norm1", outs)
    for i in range(n_blocks - 1):
        ins = outs
        outs = ins // 2
        p = f"unet.up_convs.{i}"
        sh[p + ".upconv.weight"] = (ins, outs, 2, 2)
        sh[p + ".upconv.bias"] = (outs,)
        sh[p + ".conv1.weight"] = (outs, 2 * outs, 3, 3)
        sh[p + ".conv2.weight"] = (outs, outs, 3, 3)
        bn(p + ".norm0", outs)
        bn(p + ".norm1", outs)
        bn(p + ".norm2", outs)
    sh["unet.conv_final.weight"] = (32, outs, 1, 1)
    sh["unet.conv_final.bias"] = (32,)
    sh["feature_head.0.weight"] = (head_conv, 32, 3, 3, 3)
    sh["feature_head.2.weight"] = (head_conv, head_conv, 3, 3, 3)
    for h, c in heads.items():
        sh[h + ".weight"] = (c, head_conv, 3, 1, 1)
    return sh


def unet_state_dict_np(seed: int = 317, n_blocks: int = 4, heads=None, head_conv: int = 32,
                       gain: float = 1.0) -> "OrderedDict[str, np.ndarray]":
    """W1(seed): deterministic, non-degenerate weights for every key (SURVEY.md 8(d) 'W1').

    conv / convT weights: uniform with the xavier variance 2/(fan_in+fan_out) (times gain);
    conv biases ~ U(-0.1,0.1); BN weight ~ U(0.5,1.5), bias ~ U(-0.2,0.2),
    running_mean ~ U(-0.2,0.2), running_var ~ U(0.5,1.5).
    """
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    off = 0
    for name, shape in unet_param_shapes(n_blocks, heads, head_conv).items():
        n = int(np.prod(shape)) if shape else 1
        u = uniform_np(seed, n, off)
        off += n + 7
        if name.endswith("num_batches_tracked"):
            sd[name] = np.array(1, dtype=np.int64)
            continue
        if name.endswith("running_var") or (name.endswith(".weight") and len(shape) == 1):
            v = np.float32(0.5) + u
        elif name.endswith("running_mean") or (name.endswith(".bias") and
                                               ("norm" in name or name.startswith("bn"))):
            v = (u - np.float32(0.5)) * np.float32(0.4)
        elif name.endswith(".bias"):
            v = (u - np.float32(0.5)) * np.float32(0.2)
        else:
            rf = int(np.prod(shape[2:]))
            fan_in, fan_out = shape[1] * rf, shape[0] * rf
            bound = gain * math.sqrt(6.0 / (fan_in + fan_out))
            v = (u - np.float32(0.5)) * np.float32(2.0 * bound)
        sd[name] = v.astype(np.float32).reshape(shape)
    return sd


# --------------------------------------------------------------------------- torch (device)
def _mix64_t(x):
    import torch  # int64 wrapping arithmetic; logical shifts emulated by masking
    def lsr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)
    def c(v):  # python int -> signed 64-bit constant
        v &= _M64
        return v - (1 << 64) if v >= (1 << 63) else v
    x = x + c(_GOLD)
    x = (x ^ lsr(x, 30)) * c(_C1)
    x = (x ^ lsr(x, 27)) * c(_C2)
    return x ^ lsr(x, 31)


def tomogram_torch(D: int, H: int, W: int, seed: int = 0, device="cuda", out=None):
    """Device-side twin of tomogram_np (bit-identical), generated plane by plane."""
    import torch
    if out is None:
        out = torch.empty((D, H, W), dtype=torch.float32, device=device)
    s = (seed * 0xD1342543DE82EF95) & _M64
    s = s - (1 << 64) if s >= (1 << 63) else s
    flat = out.view(-1)
    n = D * H * W
    step = 1 << 26
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        bits = _mix64_t(torch.arange(lo, hi, dtype=torch.int64, device=device) + s)
        k = ((bits >> 8) & 63) + ((bits >> 20) & 63) + ((bits >> 32) & 63) + ((bits >> 44) & 63)
        flat[lo:hi] = k.to(torch.float32) / 255.0
        del bits, k
    return out


def heatmap_tiefree_torch(D: int, H: int, W: int, seed: int = 0, device="cuda"):
    """Device-side twin of heatmap_tiefree_np (bit-identical)."""
    import torch
    n = D * H * W
    b = max(1, (n - 1).bit_length())
    assert b <= 29
    mask = (1 << b) - 1
    sh = (b + 1) // 2
    out = torch.empty(n, dtype=torch.float32, device=device)
    step = 1 << 24
    for lo in range(0, n, step):
        i = torch.arange(lo, min(n, lo + step), dtype=torch.int64, device=device)
        x = (i * (0x9E3779B1 | 1) + (seed * 2654435761 & _M64 & mask)) & mask
        x = x ^ (x >> sh)
        x = (x * (0x85EBCA6B | 1)) & mask
        x = x ^ (x >> sh)
        out[lo:lo + i.numel()] = (0x3F7FFFFF - x).to(torch.int32).view(torch.float32)
    return out.view(D, H, W)


def unet_state_dict_torch(seed: int = 317, n_blocks: int = 4, heads=None, head_conv: int = 32):
    import torch
    return OrderedDict((k, torch.from_numpy(np.array(v)) if v.ndim else torch.tensor(int(v)))
                       for k, v in unet_state_dict_np(seed, n_blocks, heads, head_conv).items())


def fullres_stub_model():
    """A detector stand-in for the semiclass tile path (detectors/tomo_det_classify.py): maps (B,D,H,W)
    to [{'hm': (B,1,D,H,W)}] at the INPUT resolution with a one-voxel receptive field halo, using only
    separately-rounded elementwise fp32 operations so CPU and CUDA results are bit-identical."""
    import torch
    import torch.nn.functional as F

    class FullResStub(torch.nn.Module):
        def forward(self, x):
            xp = F.pad(x, (1, 1, 1, 1, 1, 1))
            y = x * 6.0 - 3.0
            y = y + 0.5 * xp[:, :-2, 1:-1, 1:-1]
            y = y + 0.25 * xp[:, 1:-1, 2:, 1:-1]
            return [{"hm": y[:, None]}]

    return FullResStub()


# --------------------------------------------------------------------------- SimSiam 3-D encoder (oracle groundwork)
def simsiam3d_param_shapes(layers=(2, 2, 2), heads=("proj", "pred"), two_d=False, out_dim=256):
    """state_dict keys and shapes of cet_pick/models/networks/simsiam_model.py:159-235 `TomoResClassifier`
    (BasicBlock, three stages; arch `simsiam3d_18` / `simsiam_18`), in registration order; with two_d, of
    simsiam_model_2d.py:617-664 `TomoResClassifier2D` (arch `simsiam2d_18`: 3x3 conv1, no 3-D layer, width out_dim)."""
    sh = OrderedDict()
    od = out_dim if two_d else 256

    def bn(prefix, c, affine=True):
        if affine:
            sh[prefix + ".weight"] = (c,)
            sh[prefix + ".bias"] = (c,)
        sh[prefix + ".running_mean"] = (c,)
        sh[prefix + ".running_var"] = (c,)
        sh[prefix + ".num_batches_tracked"] = ()

    sh["conv1.weight"] = (64, 1, 3, 3) if two_d else (64, 1, 7, 7)
    bn("bn1", 64)
    inpl = 64
    for li, (planes, nblk) in enumerate(zip((64, 128, 256), layers), start=1):
        for b in range(nblk):
            p = f"layer{li}.{b}"
            stride_or_widen = (b == 0) and (li > 1 or inpl != planes)
            sh[p + ".conv1.weight"] = (planes, inpl if b == 0 else planes, 3, 3)
            bn(p + ".bn1", planes)
            sh[p + ".conv2.weight"] = (planes, planes, 3, 3)
            bn(p + ".bn2", planes)
            if stride_or_widen:
                sh[p + ".downsample.0.weight"] = (planes, inpl, 1, 1)
        inpl = planes
    if not two_d:
        sh["feature_3d.0.weight"] = (256, 256, 3, 3, 3)
        bn("feature_3d.1", 256)
    sh["fc.weight"] = (od, 256)
    sh["fc.bias"] = (od,)
    if "proj" in heads:
        for i in (0, 3, 6):
            sh[f"proj.{i}.weight"] = (od, od)
            bn(f"proj.{i + 1}", od, affine=(i != 6))
    if "pred" in heads:
        sh["pred.0.weight"] = (od, od)
        bn("pred.1", od)
        sh["pred.3.weight"] = (od, od)
        sh["pred.3.bias"] = (od,)
    return sh


def simsiam2d_state_dict_torch(seed: int = 6, layers=(2, 2, 2), heads=("proj", "pred"), out_dim=128):
    """Seeded weights of `TomoResClassifier2D` (simsiam_model_2d.py:617-664), head width out_dim = head_conv."""
    return simsiam3d_state_dict_torch(seed, layers, heads, two_d=True, out_dim=out_dim)


def simsiam3d_state_dict_torch(seed: int = 5, layers=(2, 2, 2), heads=("proj", "pred"), two_d=False, out_dim=256):
    """Seeded non-degenerate weights for the shapes above (BN statistics away from identity), as torch tensors."""
    import torch
    shapes = simsiam3d_param_shapes(layers, heads, two_d, out_dim)
    sd, off = OrderedDict(), 0
    for name, shape in shapes.items():
        n = int(np.prod(shape)) if shape else 1
        u = uniform_np(seed, n, off)
        off += n + 7
        prefix = name.rsplit(".", 1)[0]
        is_bn = (prefix + ".running_mean") in shapes
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.tensor(1)
            continue
        if name.endswith("running_var") or (is_bn and name.endswith(".weight")):
            v = np.float32(0.5) + u
        elif name.endswith("running_mean") or (is_bn and name.endswith(".bias")):
            v = (u - np.float32(0.5)) * np.float32(0.4)
        elif name.endswith(".bias"):
            v = (u - np.float32(0.5)) * np.float32(0.2)
        else:
            rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
            bound = math.sqrt(6.0 / (shape[1] * rf + shape[0] * rf))
            v = (u - np.float32(0.5)) * np.float32(2.0 * bound)
        sd[name] = torch.from_numpy(v.astype(np.float32).reshape(shape).copy())
    return sd
