"""Mirror of cet_pick/models/model.py for the hot path: create_model (:65-70), load_model
(:195-251), save_model (:283-296).  The default detector family ('unet_N') and the exploration-step embedding
networks ('simsiam_N' / 'simsiam3d_N', 'simsiam2d_N') are built; the other arch keys of the reference factory are out of scope (SURVEY.md section 2, rows 10 / Surprise 3)."""
from __future__ import annotations

import torch

from .networks.simsiam_model import get_simsiam_net_small
from .networks.simsiam_model_2d import get_simsiam2d_net_small
from .networks.unet_small import get_tomo_unet_small

_model_factory = {
    "unet": get_tomo_unet_small,
    "simsiam": get_simsiam_net_small,        # model.py:42-43: 'simsiam' and 'simsiam3d' are the same factory
    "simsiam3d": get_simsiam_net_small,
    "simsiam2d": get_simsiam2d_net_small,    # model.py:45
}


def create_model(arch, heads, head_conv, last_k=0, local_path=None):
    num_layers = int(arch[arch.find("_") + 1:]) if "_" in arch else 0
    arch = arch[:arch.find("_")] if "_" in arch else arch
    if arch not in _model_factory:
        raise KeyError(f"arch '{arch}' is outside cet_pick_b200's hot path (unet_N, simsiam[3d]_N and simsiam2d_N are built)")
    get_model = _model_factory[arch]
    return get_model(num_layers=num_layers, heads=heads, head_conv=head_conv, last_k=last_k,
                     local_path=local_path)


_HINT = ("the checkpoint does not cover the whole model: check --arch (and the head sizes) against the ones it "
         "was trained with")


def _without_data_parallel_prefix(sd):
    """keys saved from nn.DataParallel carry a leading 'module.' (but 'module_list...' is a real name)"""
    return {(k[7:] if k.startswith("module") and not k.startswith("module_list") else k): v for k, v in sd.items()}


def _reconcile(loaded, own):
    """The reference's tolerant load (model.py:207-230): tensors whose shape does not fit, and tensors the
    checkpoint lacks, keep the model's own values; extra tensors are dropped by the non-strict load."""
    merged = dict(loaded)
    for k, v in loaded.items():
        if k not in own:
            print(f"Drop parameter {k}: {_HINT}")
        elif v.shape != own[k].shape:
            print(f"Skip loading parameter {k}: model has {tuple(own[k].shape)}, checkpoint has {tuple(v.shape)}; {_HINT}")
            merged[k] = own[k]
    for k, v in own.items():
        if k not in merged:
            print(f"No param {k}: {_HINT}")
            merged[k] = v
    return merged


def load_model(model, model_path, optimizer=None, resume=False, lr=None, lr_step=None, model_only=False):
    """Checkpoint contract of the reference (model.py:195-251): a dict with 'epoch' and 'state_dict' (+ 'optimizer');
    returns the model, or (model, optimizer, start_epoch) when an optimizer is passed and `model_only` is false."""
    ckpt = torch.load(model_path, map_location="cpu")
    print(f"Loaded {model_path}, epoch {ckpt['epoch']}")
    model.load_state_dict(_reconcile(_without_data_parallel_prefix(ckpt["state_dict"]), model.state_dict()), strict=False)
    start_epoch = 0
    if optimizer is not None and resume:
        if "optimizer" in ckpt:
            optimizer.load_state_dict(ckpt["optimizer"])
            start_epoch = ckpt["epoch"]
            # the step schedule is replayed: x0.1 for every milestone already passed (:237-243)
            start_lr = lr
            for _ in (step for step in lr_step if start_epoch >= step):
                start_lr *= 0.1                       # repeated product, like the reference (same float)
            for group in optimizer.param_groups:
                group["lr"] = start_lr
            print("Resumed optimizer with start lr", start_lr)
        else:
            print("No optimizer parameters in checkpoint.")
    if optimizer is not None and not model_only:
        return model, optimizer, start_epoch
    return model


def save_model(path, epoch, model, optimizer=None):
    state_dict = model.module.state_dict() if isinstance(model, torch.nn.DataParallel) else model.state_dict()
    data = {"epoch": epoch, "state_dict": state_dict}
    if optimizer is not None:
        data["optimizer"] = optimizer.state_dict()
    torch.save(data, path)
