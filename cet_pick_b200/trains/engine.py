"""Training-mode forward + backward of the detector on the device (BASELINE.json configs[4]).

What `ModelWithLoss.forward` + `loss.backward()` + `optimizer.step()` of cet_pick/trains/base_trainer.py:135-155,484-489 do
for `TomoConvUNet` (cet_pick/models/networks/unet_small.py:63-97; UNet blocks unet.py:198-249,319-399,861-886) under
`TomoCRSemiLoss` without `--contrastive` (trains/tomo_cr_semi_trainer.py:43-60,101-104: loss = PULoss(_sigmoid(hm), gt)):

    DetectorTrainer(model).forward_backward(x, gt, tau)   # gradients accumulate into the flat bucket (.grad views)
    DetectorTrainer.step(lr)                              # all-reduce of the bucket (if distributed) + fused Adam

Every layer runs in csrc/train_net.cu (fp32, batch-statistics BatchNorm with running-stat updates); this module only owns
the activation buffers and the order of the launches.  The (D,C,h,w) <-> (1,C,D,h,w) permutes of the reference and the
channel concat of the up blocks are strides / channel offsets of the same buffers.

Not covered: `--contrastive` (UnbiasedConLoss is a (2N)^2 matrix with N = 524 288 voxels at 128^3), the
`proj` head therefore gets no gradient, exactly as in the reference when the flag is off.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from .. import _lib
from . import step as _step


class Geom(C.Structure):
    """include/cetpick.h `cetpick_conv_geom`."""
    _fields_ = [("N", C.c_int), ("Cin", C.c_int), ("Cout", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Ho", C.c_int),
                ("Wo", C.c_int), ("xs_n", C.c_longlong), ("xs_c", C.c_longlong), ("ys_n", C.c_longlong), ("ys_c", C.c_longlong),
                ("kz", C.c_int), ("ky", C.c_int), ("kx", C.c_int), ("dz", C.c_int), ("dy", C.c_int), ("dx", C.c_int),
                ("pz", C.c_int), ("py", C.c_int), ("px", C.c_int), ("stride", C.c_int), ("zdepth", C.c_int)]


@dataclass
class View:
    """C channels of a (N, Ctot, H, W) fp32 buffer starting at channel c0: dense rows, (slice, channel) strides."""
    t: torch.Tensor
    N: int
    C: int
    H: int
    W: int
    c0: int = 0

    @property
    def sn(self):
        return self.t.shape[1] * self.H * self.W

    @property
    def sc(self):
        return self.H * self.W

    @property
    def ptr(self):
        return self.t.data_ptr() + 4 * self.c0 * self.H * self.W

    def chan(self, c0, c):
        return View(self.t, self.N, c, self.H, self.W, self.c0 + c0)

    def tensor(self):
        return self.t[:, self.c0:self.c0 + self.C]


def new_view(N, Cc, H, W, device, zero=False):
    f = torch.zeros if zero else torch.empty
    return View(f((N, Cc, H, W), dtype=torch.float32, device=device), N, Cc, H, W)


@dataclass
class ConvSpec:
    k: tuple            # (kz, ky, kx)
    dil: tuple = (1, 1, 1)
    pad: tuple = (0, 0, 0)
    stride: int = 1


CONV3 = ConvSpec((1, 3, 3), (1, 1, 1), (0, 1, 1))
CONV1 = ConvSpec((1, 1, 1))
STEM = ConvSpec((1, 7, 7), (1, 1, 1), (0, 3, 3), 2)
HEAD3D = ConvSpec((3, 3, 3), (1, 4, 4), (1, 4, 4))
HM = ConvSpec((3, 1, 1), (1, 1, 1), (1, 0, 0))
UP_DGRAD = ConvSpec((1, 2, 2), (1, 1, 1), (0, 0, 0), 2)


def _geom(x: View, y: View, spec: ConvSpec, zdepth: int, cin=None, cout=None) -> Geom:
    g = Geom()
    g.N, g.Cin, g.Cout = x.N, cin if cin is not None else x.C, cout if cout is not None else y.C
    g.H, g.W, g.Ho, g.Wo = x.H, x.W, y.H, y.W
    g.xs_n, g.xs_c, g.ys_n, g.ys_c = x.sn, x.sc, y.sn, y.sc
    (g.kz, g.ky, g.kx), (g.dz, g.dy, g.dx), (g.pz, g.py, g.px) = spec.k, spec.dil, spec.pad
    g.stride, g.zdepth = spec.stride, zdepth
    return g


class Ops:
    """Thin typed wrappers over the C entry points of csrc/train_net.cu (all on the current stream)."""

    def __init__(self, device, c_max=512):
        self.L = _lib.lib()
        self.device = device
        n = C.c_size_t(0)
        _lib.check(self.L.cetpick_train_net_workspace_bytes(c_max, C.byref(n)), "cetpick_train_net_workspace_bytes")
        self._ws = torch.empty(n.value + 256, dtype=torch.uint8, device=device)
        self.ws_ptr = (self._ws.data_ptr() + 255) // 256 * 256
        self.ws_bytes = n.value
        self.launches = 0

    def _ck(self, rc, what):
        _lib.check(rc, what)
        self.launches += int(self.L.cetpick_last_launch_count())

    def conv(self, x: View, w, bias, y: View, spec: ConvSpec, zdepth, accumulate=False, relu=False):
        g = _geom(x, y, spec, zdepth)
        flags = (1 if accumulate else 0) | (2 if relu else 0)
        self._ck(self.L.cetpick_train_conv_f32(x.ptr, w.data_ptr(), bias.data_ptr() if bias is not None else None, y.ptr,
                                               C.byref(g), flags, _lib.stream_ptr()), "cetpick_train_conv_f32")

    def dgrad(self, dy: View, w, dx: View, spec: ConvSpec, zdepth, accumulate=False):
        """Gradient w.r.t. the input of a stride-1 conv with weights w [Cout][Cin][taps]."""
        assert spec.stride == 1
        cout, cin = w.shape[0], w.shape[1]
        taps = spec.k[0] * spec.k[1] * spec.k[2]
        wt = torch.empty((cin, cout, taps), dtype=torch.float32, device=w.device)
        self._ck(self.L.cetpick_train_flip_weights_f32(w.data_ptr(), wt.data_ptr(), cout, cin, taps, _lib.stream_ptr()),
                 "cetpick_train_flip_weights_f32")
        pad = tuple((k - 1) * d - p for k, d, p in zip(spec.k, spec.dil, spec.pad))
        self.conv(dy, wt, None, dx, ConvSpec(spec.k, spec.dil, pad, 1), zdepth, accumulate)

    def wgrad(self, x: View, dy: View, dw, spec: ConvSpec, zdepth):
        g = _geom(x, dy, spec, zdepth)
        self._ck(self.L.cetpick_train_conv_wgrad_f32(x.ptr, dy.ptr, dw.data_ptr(), C.byref(g), _lib.stream_ptr()),
                 "cetpick_train_conv_wgrad_f32")

    def upconv(self, x: View, w, bias, y: View):
        g = _geom(x, y, ConvSpec((1, 2, 2), (1, 1, 1), (0, 0, 0), 2), 1)
        self._ck(self.L.cetpick_train_upconv_f32(x.ptr, w.data_ptr(), bias.data_ptr() if bias is not None else None, y.ptr,
                                                 C.byref(g), _lib.stream_ptr()), "cetpick_train_upconv_f32")

    def bn(self, x: View, y: View, gamma, beta, rmean, rvar, eps=1e-5, momentum=0.1, relu=True):
        save = torch.empty((2, x.C), dtype=torch.float32, device=self.device)
        self._ck(self.L.cetpick_train_bn_f32(x.ptr, x.sn, x.sc, y.ptr, y.sn, y.sc, gamma.data_ptr(), beta.data_ptr(),
                                             rmean.data_ptr() if rmean is not None else None,
                                             rvar.data_ptr() if rvar is not None else None, save[0].data_ptr(), save[1].data_ptr(),
                                             x.N, x.C, x.H * x.W, eps, momentum, int(relu), self.ws_ptr, self.ws_bytes,
                                             _lib.stream_ptr()), "cetpick_train_bn_f32")
        return save

    def bn_bwd(self, x: View, y: View, dy: View, dx: View, gamma, save, dgamma, dbeta, relu=True):
        assert (y.sn, y.sc) == (dy.sn, dy.sc)
        self._ck(self.L.cetpick_train_bn_bwd_f32(x.ptr, x.sn, x.sc, y.ptr, dy.ptr, y.sn, y.sc, dx.ptr, dx.sn, dx.sc, gamma.data_ptr(),
                                                 save[0].data_ptr(), save[1].data_ptr(),
                                                 dgamma.data_ptr() if dgamma is not None else None,
                                                 dbeta.data_ptr() if dbeta is not None else None, x.N, x.C, x.H * x.W, int(relu),
                                                 self.ws_ptr, self.ws_bytes, _lib.stream_ptr()), "cetpick_train_bn_bwd_f32")

    def channel_sum(self, x: View, out, accumulate=True):
        self._ck(self.L.cetpick_train_channel_sum_f32(x.ptr, x.sn, x.sc, out.data_ptr(), x.N, x.C, x.H * x.W, int(accumulate),
                                                      self.ws_ptr, self.ws_bytes, _lib.stream_ptr()), "cetpick_train_channel_sum_f32")

    def pool(self, x: View, y: View):
        self._ck(self.L.cetpick_train_pool_f32(x.ptr, x.sn, x.sc, y.ptr, y.sn, y.sc, x.N, x.C, x.H, x.W, _lib.stream_ptr()),
                 "cetpick_train_pool_f32")

    def pool_bwd(self, x: View, dy: View, dx: View, accumulate):
        self._ck(self.L.cetpick_train_pool_bwd_f32(x.ptr, x.sn, x.sc, dy.ptr, dy.sn, dy.sc, dx.ptr, dx.sn, dx.sc, x.N, x.C, x.H, x.W,
                                                   int(accumulate), _lib.stream_ptr()), "cetpick_train_pool_bwd_f32")

    def relu_bwd(self, y: View, dy: View, dx: View):
        assert y.C == y.t.shape[1] and dy.C == dy.t.shape[1] and dx.C == dx.t.shape[1]
        self._ck(self.L.cetpick_train_relu_bwd_f32(y.ptr, dy.ptr, dx.ptr, y.t.numel(), _lib.stream_ptr()), "cetpick_train_relu_bwd_f32")


class DetectorTrainer:
    """Owns the flat parameter bucket of a `TomoConvUNet` and runs its training step on the device."""

    def __init__(self, model: torch.nn.Module, tau: float = 0.01, beta: float = 0.0):
        self.model = model
        self.P = dict(model.named_parameters())
        self.B = dict(model.named_buffers())
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise _lib.CetpickError(-4, "DetectorTrainer", "the training step runs on a CUDA device only (no CPU fallback)")
        self.device = dev
        self.bucket = _step.FlatBucket(model)
        self.P = dict(model.named_parameters())            # .data / .grad are views into the bucket now
        self.nb = 0
        while f"unet.down_convs.{self.nb}.conv1.weight" in self.P:
            self.nb += 1
        self.tau, self.beta = tau, beta
        self.ops = Ops(dev)
        self.stats = {}

    # ------------------------------------------------------------------------------------------------ helpers
    def _w(self, name):
        return self.P[name].data

    def _g(self, name):
        return self.P[name].grad

    def _bn(self, x, y, p, relu=True):
        return self.ops.bn(x, y, self._w(p + ".weight"), self._w(p + ".bias"), self.B[p + ".running_mean"], self.B[p + ".running_var"],
                           relu=relu)

    def _bn_bwd(self, x, y, dy, dx, p, save, relu=True):
        self.ops.bn_bwd(x, y, dy, dx, self._w(p + ".weight"), save, self._g(p + ".weight"), self._g(p + ".bias"), relu=relu)

    # ------------------------------------------------------------------------------------------------ forward
    def _forward(self, x: torch.Tensor):
        """Training-mode forward (batch-statistics BatchNorm, running statistics updated); keeps every activation."""
        o, nb, dev = self.ops, self.nb, self.device
        if x.dim() > 4:
            x = x.squeeze()
        if x.dim() == 3:
            x = x[None]
        b, d, h, w = x.shape
        N = b * d
        xin = View(x.contiguous().float().view(N, 1, h, w), N, 1, h, w)
        h1, w1 = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        z0 = new_view(N, 16, h1, w1, dev)
        a0 = new_view(N, 16, h1, w1, dev)
        o.conv(xin, self._w("conv1.weight"), None, z0, STEM, 1)
        s_stem = self._bn(z0, a0, "bn1")
        chans = [32 << i for i in range(nb)]
        sizes = [(h1, w1)]
        for i in range(1, nb):
            sizes.append(((sizes[-1][0] + 1) // 2, (sizes[-1][1] + 1) // 2))
        # merged buffers of the up blocks: [0, C) = upsampled path, [C, 2C) = encoder output of the same level
        merged = [new_view(N, 2 * chans[i], *sizes[i], dev) for i in range(nb - 1)]
        down = []
        cur = a0
        for i in range(nb):
            p = f"unet.down_convs.{i}"
            Cc, (hh, ww) = chans[i], sizes[i]
            c1, a1, c2 = new_view(N, Cc, hh, ww, dev), new_view(N, Cc, hh, ww, dev), new_view(N, Cc, hh, ww, dev)
            a2 = merged[i].chan(Cc, Cc) if i < nb - 1 else new_view(N, Cc, hh, ww, dev)
            o.conv(cur, self._w(p + ".conv1.weight"), None, c1, CONV3, 1)
            s0 = self._bn(c1, a1, p + ".norm0")
            o.conv(a1, self._w(p + ".conv2.weight"), None, c2, CONV3, 1)
            s1 = self._bn(c2, a2, p + ".norm1")
            rec = dict(inp=cur, c1=c1, a1=a1, c2=c2, a2=a2, s0=s0, s1=s1)
            if i < nb - 1:
                pooled = new_view(N, Cc, *sizes[i + 1], dev)
                o.pool(a2, pooled)
                cur = pooled
            else:
                cur = a2
            down.append(rec)
        ups = []
        for j in range(nb - 1):
            p = f"unet.up_convs.{j}"
            lvl = nb - 2 - j
            Cc, (hh, ww) = chans[lvl], sizes[lvl]
            u = new_view(N, Cc, hh, ww, dev)
            o.upconv(cur, self._w(p + ".upconv.weight"), self._w(p + ".upconv.bias"), u)
            M = merged[lvl]
            s0 = self._bn(u, M.chan(0, Cc), p + ".norm0")
            c1, a1, c2, a2 = (new_view(N, Cc, hh, ww, dev) for _ in range(4))
            o.conv(M, self._w(p + ".conv1.weight"), None, c1, CONV3, 1)
            s1 = self._bn(c1, a1, p + ".norm1")
            o.conv(a1, self._w(p + ".conv2.weight"), None, c2, CONV3, 1)
            s2 = self._bn(c2, a2, p + ".norm2")
            ups.append(dict(inp=cur, u=u, M=M, c1=c1, a1=a1, c2=c2, a2=a2, s0=s0, s1=s1, s2=s2, lvl=lvl))
            cur = a2
        trunk_out = cur
        f = new_view(N, 32, h1, w1, dev)
        o.conv(trunk_out, self._w("unet.conv_final.weight"), self._w("unet.conv_final.bias"), f, CONV1, 1)
        hc = self._w("feature_head.0.weight").shape[0]
        r0, r1 = new_view(N, hc, h1, w1, dev), new_view(N, hc, h1, w1, dev)
        o.conv(f, self._w("feature_head.0.weight"), None, r0, HEAD3D, d, relu=True)
        o.conv(r0, self._w("feature_head.2.weight"), None, r1, HEAD3D, d, relu=True)
        hm = new_view(N, 1, h1, w1, dev)
        o.conv(r1, self._w("hm.weight"), None, hm, HM, d)
        return dict(b=b, d=d, N=N, h1=h1, w1=w1, xin=xin, z0=z0, a0=a0, s_stem=s_stem, chans=chans, sizes=sizes, down=down, ups=ups,
                    trunk_out=trunk_out, f=f, r0=r0, r1=r1, hc=hc, logits=hm.t.view(b, 1, d, h1, w1))

    def forward_train(self, x: torch.Tensor) -> torch.Tensor:
        """`model.train(); model(x)[-1]['hm']` -- the logits only (the reference's second, augmented view without
        `--contrastive`: it moves the BatchNorm running statistics and nothing else, base_trainer.py:143-145)."""
        _lib.require_cuda(x, "DetectorTrainer.forward_train")
        return self._forward(x)["logits"]

    # ------------------------------------------------------------------------------------------------ forward + backward
    def forward_backward(self, x: torch.Tensor, gt: torch.Tensor, x_aug: torch.Tensor = None, want_logits: bool = False):
        """x: (b, d, h, w) fp32 crops on the device, gt: (b, 1, d, h/2, w/2)-shaped labels (1 / soft / -1, loss.py:255-263).
        Parameter gradients of loss = PULoss(tau)(_sigmoid(hm), gt) are ADDED to the bucket.  x_aug: the augmented view
        the reference also pushes through the model in training mode.  -> loss (0-dim tensor)."""
        _lib.require_cuda(x, "DetectorTrainer.forward_backward")
        o, nb, dev = self.ops, self.nb, self.device
        o.launches = 0
        A = self._forward(x)
        if x_aug is not None:
            self._forward(x_aug)
        b, d, N, h1, w1, hc = A["b"], A["d"], A["N"], A["h1"], A["w1"], A["hc"]
        xin, z0, a0, s_stem, chans, sizes, down, ups = (A[k] for k in ("xin", "z0", "a0", "s_stem", "chans", "sizes", "down", "ups"))
        trunk_out, f, r0, r1, logits = A["trunk_out"], A["f"], A["r0"], A["r1"], A["logits"]

        # ---- loss (tomo_cr_semi_trainer.py:52-60) and its gradient w.r.t. the logits
        loss, dlog, st = _step.pu_loss(logits, gt, self.tau, self.beta)
        self.stats = {"loss": loss, "pos_risk": st[1], "neg_risk": st[2], "n_pos": st[3]}
        d_hm = View(dlog.view(N, 1, h1, w1), N, 1, h1, w1)

        # ---- backward: 3-D head
        o.wgrad(r1, d_hm, self._g("hm.weight"), HM, d)
        d_r1 = new_view(N, hc, h1, w1, dev)
        o.dgrad(d_hm, self._w("hm.weight"), d_r1, HM, d)
        o.relu_bwd(r1, d_r1, d_r1)
        o.wgrad(r0, d_r1, self._g("feature_head.2.weight"), HEAD3D, d)
        d_r0 = new_view(N, hc, h1, w1, dev)
        o.dgrad(d_r1, self._w("feature_head.2.weight"), d_r0, HEAD3D, d)
        o.relu_bwd(r0, d_r0, d_r0)
        o.wgrad(f, d_r0, self._g("feature_head.0.weight"), HEAD3D, d)
        d_f = new_view(N, 32, h1, w1, dev)
        o.dgrad(d_r0, self._w("feature_head.0.weight"), d_f, HEAD3D, d)
        del d_r0, d_r1
        # ---- conv_final
        o.wgrad(trunk_out, d_f, self._g("unet.conv_final.weight"), CONV1, 1)
        o.channel_sum(d_f, self._g("unet.conv_final.bias"))
        d_cur = new_view(N, trunk_out.C, trunk_out.H, trunk_out.W, dev)
        o.dgrad(d_f, self._w("unet.conv_final.weight"), d_cur, CONV1, 1)
        del d_f
        # ---- up blocks, last to first
        d_merged = [None] * (nb - 1)
        for j in range(nb - 2, -1, -1):
            p = f"unet.up_convs.{j}"
            r = ups[j]
            Cc, (hh, ww), lvl = chans[r["lvl"]], sizes[r["lvl"]], r["lvl"]
            d_c = new_view(N, Cc, hh, ww, dev)
            self._bn_bwd(r["c2"], r["a2"], d_cur, d_c, p + ".norm2", r["s2"])
            o.wgrad(r["a1"], d_c, self._g(p + ".conv2.weight"), CONV3, 1)
            d_a1 = new_view(N, Cc, hh, ww, dev)
            o.dgrad(d_c, self._w(p + ".conv2.weight"), d_a1, CONV3, 1)
            self._bn_bwd(r["c1"], r["a1"], d_a1, d_c, p + ".norm1", r["s1"])
            o.wgrad(r["M"], d_c, self._g(p + ".conv1.weight"), CONV3, 1)
            dM = new_view(N, 2 * Cc, hh, ww, dev)
            o.dgrad(d_c, self._w(p + ".conv1.weight"), dM, CONV3, 1)
            d_merged[lvl] = dM
            d_u = d_a1                                        # reuse
            self._bn_bwd(r["u"], r["M"].chan(0, Cc), dM.chan(0, Cc), d_u, p + ".norm0", r["s0"])
            o.channel_sum(d_u, self._g(p + ".upconv.bias"))
            # transposed-conv weight gradient: dw[ci][co][a][b] = sum in[ci][y][x] * d_u[co][2y+a][2x+b]
            o.wgrad(d_u, r["inp"], self._g(p + ".upconv.weight"), UP_DGRAD, 1)
            d_cur = new_view(N, r["inp"].C, r["inp"].H, r["inp"].W, dev)
            o.conv(d_u, self._w(p + ".upconv.weight"), None, d_cur, UP_DGRAD, 1)
        # ---- down blocks, last to first
        for i in range(nb - 1, -1, -1):
            p = f"unet.down_convs.{i}"
            r = down[i]
            Cc, (hh, ww) = chans[i], sizes[i]
            if i < nb - 1:
                d_a2 = d_merged[i].chan(Cc, Cc)               # skip-connection gradient ...
                o.pool_bwd(r["a2"], d_cur, d_a2, accumulate=True)    # ... plus the pooled path
            else:
                d_a2 = d_cur
            d_c = new_view(N, Cc, hh, ww, dev)
            self._bn_bwd(r["c2"], r["a2"], d_a2, d_c, p + ".norm1", r["s1"])
            o.wgrad(r["a1"], d_c, self._g(p + ".conv2.weight"), CONV3, 1)
            d_a1 = new_view(N, Cc, hh, ww, dev)
            o.dgrad(d_c, self._w(p + ".conv2.weight"), d_a1, CONV3, 1)
            self._bn_bwd(r["c1"], r["a1"], d_a1, d_c, p + ".norm0", r["s0"])
            o.wgrad(r["inp"], d_c, self._g(p + ".conv1.weight"), CONV3, 1)
            d_cur = new_view(N, r["inp"].C, r["inp"].H, r["inp"].W, dev)
            o.dgrad(d_c, self._w(p + ".conv1.weight"), d_cur, CONV3, 1)
        # ---- stem
        d_z0 = new_view(N, 16, h1, w1, dev)
        self._bn_bwd(z0, a0, d_cur, d_z0, "bn1", s_stem)
        o.wgrad(xin, d_z0, self._g("conv1.weight"), STEM, 1)
        self.stats["launches"] = o.launches
        return (loss, logits) if want_logits else loss

    # ------------------------------------------------------------------------------------------------ optimiser
    def zero_grad(self):
        self.bucket.zero_grad()

    def step(self, lr: float, group=None):
        """Gradient all-reduce over the data-parallel ranks (one flat bucket) + torch.optim.Adam step (main.py:55)."""
        scale = _step.allreduce_gradients(self.bucket, group=group, average=False)
        self.bucket.adam_step(lr, grad_scale=scale)
        if hasattr(self.model, "_destroy_plan"):
            self.model._destroy_plan()                        # the inference plan packs the weights: rebuild it lazily
