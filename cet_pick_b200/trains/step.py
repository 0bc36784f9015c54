"""Device-side pieces of the refinement training step around the network (reference
cet_pick/trains/base_trainer.py:135-155,446-552, trains/tomo_cr_semi_trainer.py:43-112, main.py:55):

  * `TomoCRSemiLoss`-equivalent loss without `--contrastive`: `_sigmoid` + PULoss and its gradient w.r.t. the logits
    (csrc/train.cu), plus the ConsistencyLoss MSE;
  * `FlatBucket`: every parameter of a module viewed inside ONE flat fp32 buffer (values, gradients, Adam moments), so
    that the data-parallel gradient exchange is a single all-reduce of 7.97 MB for unet_4 (latency-bound: one bucket,
    not DDP's per-layer buckets) and the optimiser is one fused kernel;
  * `allreduce_gradients`: that all-reduce (NCCL on CUDA tensors, gloo on CPU tensors), averaged over the ranks.

The forward / backward of the U-Net in training mode (batch-statistics BatchNorm, dgrad / wgrad kernels) is NOT built:
gradients have to come from elsewhere (tests feed reference gradients); see DESIGN.md."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from .. import _lib

_ws_cache = {}


def _ws(device):
    n = C.c_size_t(0)
    _lib.check(_lib.lib().cetpick_train_workspace_bytes(C.byref(n)), "cetpick_train_workspace_bytes")
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    if key not in _ws_cache:
        _ws_cache[key] = torch.empty(n.value + 256, dtype=torch.uint8, device=device)
    ws = _ws_cache[key]
    ptr = (ws.data_ptr() + 255) // 256 * 256
    return ws, ptr, ws.numel() - (ptr - ws.data_ptr())


def pu_loss(logits: torch.Tensor, gt: torch.Tensor, tau: float, beta: float = 0.0, apply_sigmoid: bool = True,
            want_grad: bool = True, grad_scale: float = 1.0):
    """loss.py:255-325 `PULoss(tau)(pred, gt)` with pred = `_sigmoid(logits)` (models/utils.py:167-169) when
    apply_sigmoid.  -> (loss 0-dim tensor, grad w.r.t. `logits` or None, stats tensor [loss, pos_risk, neg_risk, n_pos]).
    Raises ValueError like the reference when no voxel is a labelled positive (loss.py:275-276)."""
    _lib.require_cuda(logits, "pu_loss")
    x = logits.contiguous().float()
    g = gt.to(x.device).contiguous().float()
    if x.numel() != g.numel():
        raise ValueError("pu_loss: prediction and target must have the same number of elements")
    out = torch.empty(4, dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    ws, ptr, nbytes = _ws(x.device)
    _lib.check(_lib.lib().cetpick_pu_loss_f32(x.data_ptr(), g.data_ptr(), x.numel(), int(apply_sigmoid), float(tau), float(beta),
                                              out.data_ptr(), grad.data_ptr() if grad is not None else None, float(grad_scale),
                                              ptr, nbytes, _lib.stream_ptr()), "cetpick_pu_loss_f32")
    if float(out[3]) == 0:
        raise ValueError("Num of true positive is zero")
    return out[0], (grad.view_as(logits) if grad is not None else None), out


def consistency_loss(a: torch.Tensor, b: torch.Tensor, want_grad: bool = True, grad_scale: float = 1.0):
    """loss.py:701-715 `ConsistencyLoss` = mse_loss(a, b) -> (loss, grad w.r.t. a or None)."""
    _lib.require_cuda(a, "consistency_loss")
    x, y = a.contiguous().float(), b.to(a.device).contiguous().float()
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    ws, ptr, nbytes = _ws(x.device)
    _lib.check(_lib.lib().cetpick_mse_loss_f32(x.data_ptr(), y.data_ptr(), x.numel(), out.data_ptr(),
                                               grad.data_ptr() if grad is not None else None, float(grad_scale), ptr, nbytes,
                                               _lib.stream_ptr()), "cetpick_mse_loss_f32")
    return out[0], (grad.view_as(a) if grad is not None else None)


class FlatBucket:
    """All parameters of `module` re-homed into one flat fp32 tensor (`.params`); `.grads`, `.exp_avg`, `.exp_avg_sq`
    have the same layout.  `param.data` / `param.grad` of every parameter become views into the flat tensors, so
    whatever produces gradients writes straight into the bucket."""

    def __init__(self, module: torch.nn.Module):
        ps = [p for p in module.parameters() if p.requires_grad]
        if not ps:
            raise ValueError("FlatBucket: the module has no trainable parameters")
        dev = ps[0].device
        self.numel = sum(p.numel() for p in ps)
        self.params = torch.empty(self.numel, dtype=torch.float32, device=dev)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.step_count = 0
        off = 0
        self.slices = []
        for p in ps:
            n = p.numel()
            self.params[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.params[off:off + n].view_as(p.data)
            p.grad = self.grads[off:off + n].view_as(p.data)
            self.slices.append((off, n))
            off += n

    def zero_grad(self):
        self.grads.zero_()

    def adam_step(self, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, grad_scale: float = 1.0):
        """torch.optim.Adam(model.parameters(), lr) of main.py:55, one fused kernel over the bucket (CUDA only)."""
        _lib.require_cuda(self.params, "FlatBucket.adam_step")
        self.step_count += 1
        _lib.check(_lib.lib().cetpick_adam_step_f32(self.params.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(),
                                                    self.exp_avg_sq.data_ptr(), self.numel, float(lr), float(betas[0]),
                                                    float(betas[1]), float(eps), float(weight_decay), self.step_count,
                                                    float(grad_scale), _lib.stream_ptr()), "cetpick_adam_step_f32")


def allreduce_gradients(bucket: FlatBucket, group=None, average: bool = True):
    """The data-parallel exchange of the training step: ONE all-reduce of the flat gradient bucket (DDP's result with a
    single bucket; base_trainer.py:229-238 wraps the model in DistributedDataParallel).  Returns the scale still to be
    applied (1.0 when averaged here; pass 1/world as `grad_scale` of adam_step to fold the division into the optimiser)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(bucket.grads, op=dist.ReduceOp.SUM, group=group)
    if average:
        bucket.grads.div_(world)
        return 1.0
    return 1.0 / world
