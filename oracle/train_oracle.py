"""TEST INFRASTRUCTURE (CPU restatement, torch fp32, differentiable): the loss functions of the refinement training
step (BASELINE.json configs[4]; cet_pick/trains/tomo_cr_semi_trainer.py:17-112 combines them):
  `_pu_neg_loss` / `PULoss`  cet_pick/models/loss.py:255-325   (non-negative positive-unlabeled focal loss)
  `_neg_loss` / `FocalLoss`  cet_pick/models/loss.py:378-437   (CornerNet focal loss, validation criterion)
  `ConsistencyLoss`          cet_pick/models/loss.py:701-715   (MSE between the two views)
Groundwork for SURVEY 8f-4: no CUDA path uses or mirrors it yet.  Pinned by tests/golden/train_losses.npz, values AND
gradients produced by the unmodified reference (tests/golden/make_golden_train.py)."""
from __future__ import annotations

import torch


def pu_focal_loss(pred, gt, tau, beta=0.0):
    """loss.py:255-308.  gt: 1 = labelled positive, (-1,1) open = soft positive (Gaussian shoulder), -1 = unlabelled."""
    pred, gt = pred.squeeze(), gt.squeeze()
    pos = gt.eq(1).float()
    soft = (gt.gt(-1).float() == gt.lt(1).float()).float()
    unl = gt.eq(-1).float()
    n_pos, n_soft, n_unl = pos.sum(), soft.sum(), unl.sum()
    if n_pos == 0:
        raise ValueError("Num of true positive is zero")
    lp, ln = torch.log(pred), torch.log(1 - pred)
    pos_term = -(lp * (1 - pred) ** 2 * pos).sum() / n_pos
    neg_pos_term = -(ln * pred ** 2 * pos).sum() / n_pos
    if n_soft > 0:
        pos_term = pos_term - (ln * pred ** 2 * (1 - gt) ** 4 * soft).sum() / n_soft
        neg_pos_term = neg_pos_term - (lp * (1 - pred) ** 2 * gt ** 4 * soft).sum() / n_soft
    pos_risk = pos_term * tau
    unl_risk = -(pred ** 2 * ln * unl).sum() / n_unl
    neg_risk = -tau * neg_pos_term + unl_risk
    return pos_risk if neg_risk < -beta else pos_risk + neg_risk


def focal_loss(pred, gt):
    """loss.py:378-411."""
    gt = gt.unsqueeze(0)
    pos = gt.eq(1).float()
    neg = (gt.gt(-1).float() == gt.lt(1).float()).float()
    pos_loss = (torch.log(pred) * (1 - pred) ** 2 * pos).sum()
    neg_loss = (torch.log(1 - pred) * pred ** 2 * (1 - gt) ** 4 * neg).sum()
    n_pos = pos.sum()
    return -neg_loss if n_pos == 0 else -(pos_loss + neg_loss) / n_pos


def consistency_loss(a, b):
    """loss.py:701-715."""
    return torch.nn.functional.mse_loss(a, b)


def training_step(x, gt, sd, tau, beta=0.0, param_names=None):
    """One training forward + backward of the detector as trains/base_trainer.py:135-155,484-489 runs it for
    `TomoCRSemiLoss` without `--contrastive` (tomo_cr_semi_trainer.py:52-60,101-104): model in train mode (batch-statistics
    BatchNorm, running statistics updated in `sd` IN PLACE), loss = PULoss(tau)(_sigmoid(hm), gt), autograd gradients.
    x: (b,d,h,w); gt: same number of elements as hm.  -> (loss, {name: grad}, hm logits)."""
    from . import unet_oracle as uo
    names = param_names or [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k and "num_batches" not in k]
    leaves = {k: sd[k].detach().clone().requires_grad_(True) for k in names}
    sdl = dict(sd)
    sdl.update(leaves)
    out = uo.forward(x, sdl, want_proj=False, train=True)
    hm = out["hm"]
    loss = pu_focal_loss(uo.sigmoid_clamp(hm), gt.reshape(hm.shape), tau, beta)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    return loss.detach(), {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, grads)}, hm.detach()
