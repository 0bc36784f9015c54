"""Host-side mirror of cet_pick/models/networks/unet_small.py (TomoConvUNet) for the default
detector.  The nn.Module below is ONLY a container with the reference's parameter names, shapes,
registration order and initialisers (so state_dicts and seeded random init are interchangeable
with the reference); its forward() runs entirely in libcetpick_sm100a.so (csrc/unet.cu,
csrc/conv_tc.cu).  There is no PyTorch/CPU forward path.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from ... import _lib


class _DownParams(nn.Module):
    """Parameters of unet.py:198-249 DownConv(dim=2, normalization='batch', full_norm=True)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False)
        self.norm0 = nn.BatchNorm2d(cout)
        self.norm1 = nn.BatchNorm2d(cout)


class _UpParams(nn.Module):
    """Parameters of unet.py:319-399 UpConv(merge_mode='concat', up_mode='transpose', dim=2)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.upconv = nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2)
        self.conv1 = nn.Conv2d(2 * cout, cout, 3, padding=1, bias=False)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1, bias=False)
        self.norm0 = nn.BatchNorm2d(cout)
        self.norm1 = nn.BatchNorm2d(cout)
        self.norm2 = nn.BatchNorm2d(cout)


class _UNetParams(nn.Module):
    """Parameters of unet.py:538-859 UNet(16, 32, n_blocks, start_filts=32, dim=2)."""

    def __init__(self, n_blocks):
        super().__init__()
        self.down_convs = nn.ModuleList()
        self.up_convs = nn.ModuleList()
        outs = 16
        for i in range(n_blocks):
            ins = 16 if i == 0 else outs
            outs = 32 * (2 ** i)
            self.down_convs.append(_DownParams(ins, outs))
        for i in range(n_blocks - 1):
            ins = outs
            outs = ins // 2
            self.up_convs.append(_UpParams(ins, outs))
        self.conv_final = nn.Conv2d(outs, 32, kernel_size=1)
        self.apply(self._weight_init)          # unet.py:850-859

    @staticmethod
    def _weight_init(m):
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            nn.init.xavier_normal_(m.weight)
            if getattr(m, "bias") is not None:
                nn.init.constant_(m.bias, 0)


class TomoConvUNet(nn.Module):
    """unet_small.py:30-97.  forward(x) -> [ {'hm': (B,1,D,h,w), 'proj': (B,C,D,h,w)} ]."""

    def __init__(self, n_blocks, heads, head_conv, last_k=3):
        super().__init__()
        self.heads = heads
        self.n_blocks = n_blocks
        self.head_conv = head_conv
        self.conv1 = nn.Conv2d(1, 16, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(16)
        self.unet = _UNetParams(n_blocks)
        self.feature_head = nn.Sequential(
            nn.Conv3d(32, head_conv, (3, 3, 3), dilation=(1, 4, 4), padding=(1, 4, 4), bias=False),
            nn.Identity(),   # ReLU slot (keeps the reference's index 2 for the second conv)
            nn.Conv3d(head_conv, head_conv, (3, 3, 3), dilation=(1, 4, 4), padding=(1, 4, 4), bias=False),
            nn.Identity())
        for m in self.feature_head:                              # fill_fc_weights, unet_small.py:17-28
            if isinstance(m, nn.Conv3d):
                nn.init.normal_(m.weight, std=0.001)
        for head, classes in self.heads.items():
            fc = nn.Conv3d(head_conv, classes, kernel_size=(3, 1, 1), padding=(1, 0, 0), bias=False)
            nn.init.normal_(fc.weight, std=0.001)
            setattr(self, head, fc)
        # z-slab streaming (north_star "sliding-slab scheduler"; the reference forwards the whole volume,
        # tomo_det.py:26): None = whole volume, int = core slices per slab, "auto" = largest slab whose
        # workspace fits in free device memory.  The 2-D trunk is per-slice and the 3-D head reaches
        # +-3 slices (feature_head.0, feature_head.2, hm/proj: one slice each), so every slab is
        # forwarded with a 3-slice recompute halo and only its core slices are kept: exact.
        self.slab_z = None
        # Quantised input: a uint8 tensor is read as the 256 levels utils/loader.py's preprocess ends with; level k
        # stands for the float32 value level_values[k] (default k/255, i.e. levels spanning 0..255).  Lossless
        # and a quarter of the host->device bytes of the float32 volume the reference ships.
        self.level_values = None
        # absolute z of plane 0 of the tensors passed to forward() (a z-shard of a larger volume, shard.py)
        self.z_origin = 0
        # operand precision of the tensor-core convolutions: "bf16" (default, heat-map within 1e-2 of the fp32 reference)
        # or "tf32" (within 1e-4; generic kernel, about half the throughput ceiling).  north_star's two tolerances.
        self.precision = "bf16"
        self.compute_proj = "proj" in heads      # detectors switch this off (they never read 'proj')
        self.fuse_sigmoid = False                # TomodetDetector fuses _sigmoid into the hm epilogue
        self._plan = None
        self._plan_key = None
        self._ws = None

    # ------------------------------------------------------------------ plan management
    def _param_key(self):
        return (self.precision,) + tuple((k, v.data_ptr(), v._version) for k, v in self.state_dict().items())

    def _destroy_plan(self):
        if self._plan is not None:
            _lib.lib().cetpick_unet_destroy(self._plan)
            self._plan = None

    def __del__(self):
        try:
            self._destroy_plan()
        except Exception:
            pass

    def plan(self):
        """(Re)build the device plan when the parameters changed: BN folding + bf16 packing."""
        key = self._param_key()
        if self._plan is not None and key == self._plan_key:
            return self._plan
        self._destroy_plan()
        if self.heads.get("hm", 1) != 1:
            raise NotImplementedError("cet_pick_b200: the hm head must have one class (opts.py:286-289)")
        L = _lib.lib()
        h = C.c_void_p()
        _lib.check(L.cetpick_unet_create(C.byref(h), self.n_blocks, self.head_conv,
                                         int(self.heads.get("proj", 0))), "cetpick_unet_create")
        for k, v in self.state_dict().items():
            if k.endswith("num_batches_tracked"):
                continue
            t = v.detach().to("cpu", torch.float32).contiguous()
            _lib.check(L.cetpick_unet_set_param(h, k.encode(), t.data_ptr(), t.numel()), f"set_param({k})")
        if self.precision not in ("bf16", "tf32"):
            raise ValueError("TomoConvUNet.precision must be 'bf16' or 'tf32'")
        _lib.check(L.cetpick_unet_set_precision(h, 1 if self.precision == "tf32" else 0), "cetpick_unet_set_precision")
        _lib.check(L.cetpick_unet_finalize(h), "cetpick_unet_finalize")
        self._plan, self._plan_key = h, key
        return h

    # ------------------------------------------------------------------ forward
    HEAD_HALO = 3

    def _slab_depth(self, plan, d, h, w, device):
        """core slices per slab (>= 1), or d when the whole volume is forwarded at once"""
        if self.slab_z is None:
            return d
        if self.slab_z != "auto":
            return max(1, min(int(self.slab_z), d))
        L = _lib.lib()
        free, _ = torch.cuda.mem_get_info(device)
        budget = int(0.8 * free) + (self._ws.numel() if self._ws is not None and self._ws.device == device else 0)
        nb = C.c_size_t(0)
        _lib.check(L.cetpick_unet_workspace_bytes(plan, d, h, w, 0, C.byref(nb)), "unet_workspace_bytes")
        if nb.value <= budget:
            return d
        _lib.check(L.cetpick_unet_workspace_bytes(plan, 1, h, w, 0, C.byref(nb)), "unet_workspace_bytes")
        per_slice = max(1, nb.value)                       # the workspace is linear in the slab depth
        depth = budget // per_slice - 2 * self.HEAD_HALO
        if depth < 1:
            raise RuntimeError("TomoConvUNet: not enough device memory for a one-slice slab")
        return int(min(depth, d))

    def forward(self, x):
        _lib.require_cuda(x, "TomoConvUNet.forward")
        if x.dim() > 4:
            x = x.squeeze()                                   # unet_small.py:64-65
        if x.dim() == 3:
            x = x.unsqueeze(0)
        b, d, h, w = x.shape
        lut = None
        if x.dtype == torch.uint8:
            import numpy as np
            lv = self.level_values
            lut = np.ascontiguousarray((np.arange(256, dtype=np.float64) / 255.0).astype(np.float32) if lv is None
                                       else np.asarray(lv, dtype=np.float32))
            if lut.shape != (256,) or lut[0] != 0:
                raise ValueError("TomoConvUNet: level_values must be 256 float32 values with level_values[0] == 0")
            if w % 16 or self.precision == "tf32":   # the uint8 TMA stem is BF16-only / needs 16-byte rows: expand on the device
                x, lut = torch.from_numpy(lut).to(x.device)[x.long()], None
            else:
                x = x.contiguous()
        if lut is None:
            x = x.to(torch.float32).contiguous()
        plan = self.plan()
        L = _lib.lib()
        oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        slab = self._slab_depth(plan, d, h, w, x.device)
        halo = self.HEAD_HALO if slab < d else 0
        dmax = min(d, slab + 2 * halo)
        nbytes = C.c_size_t(0)
        _lib.check(L.cetpick_unet_workspace_bytes(plan, dmax, h, w, 0, C.byref(nbytes)), "unet_workspace_bytes")
        if self._ws is None or self._ws.numel() < nbytes.value or self._ws.device != x.device:
            self._ws = None
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=x.device)
        hm = torch.empty((b, 1, d, oh, ow), dtype=torch.float32, device=x.device)
        pc = int(self.heads.get("proj", 0))
        proj = torch.empty((b, pc, d, oh, ow), dtype=torch.float32, device=x.device) \
            if (self.compute_proj and pc > 0) else None
        self.last_launches = 0
        self.last_slabs = 0
        sig = 1 if self.fuse_sigmoid else 0

        def run(src, depth, hm_dst, proj_dst, z_origin=0):
            pj = proj_dst.data_ptr() if proj_dst is not None else None
            z_origin += int(self.z_origin)
            _lib.check(L.cetpick_unet_forward_slab(plan, None if lut is not None else src.data_ptr(),
                                                   src.data_ptr() if lut is not None else None,
                                                   lut.ctypes.data_as(C.c_void_p) if lut is not None else None,
                                                   depth, h, w, z_origin, hm_dst.data_ptr(), sig, pj,
                                                   self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr()),
                       "cetpick_unet_forward_slab")
            self.last_launches += L.cetpick_last_launch_count()
            self.last_slabs += 1

        for i in range(b):
            if slab >= d:
                run(x[i], d, hm[i], proj[i] if proj is not None else None)
                continue
            hm_s = torch.empty((dmax, oh, ow), dtype=torch.float32, device=x.device)
            pj_s = torch.empty((pc, dmax, oh, ow), dtype=torch.float32, device=x.device) if proj is not None else None
            for z0 in range(0, d, slab):
                z1 = min(d, z0 + slab)
                lo, hi = max(0, z0 - halo), min(d, z1 + halo)
                # (pc, depth, oh, ow) output of a slab is contiguous only for the slab's own depth
                pj_v = pj_s.view(-1)[:pc * (hi - lo) * oh * ow].view(pc, hi - lo, oh, ow) if pj_s is not None else None
                run(x[i, lo:hi], hi - lo, hm_s, pj_v, z_origin=lo)
                hm[i, 0, z0:z1].copy_(hm_s[z0 - lo:z0 - lo + (z1 - z0)])
                if pj_v is not None:
                    proj[i, :, z0:z1].copy_(pj_v[:, z0 - lo:z0 - lo + (z1 - z0)])
        ret = {"hm": hm}
        if proj is not None:
            ret["proj"] = proj
        return [ret]


def get_tomo_unet_small(num_layers, heads, head_conv=32, last_k=3, local_path=None):
    """unet_small.py:189-193."""
    return TomoConvUNet(num_layers, heads, head_conv, last_k)
