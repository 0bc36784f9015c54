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
