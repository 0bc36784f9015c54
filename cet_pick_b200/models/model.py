"""Mirror of cet_pick/models/model.py for the hot path: create_model (:65-70), load_model
(:195-251), save_model (:283-296).  Only the default detector family ('unet_N') is built; the other
20 arch keys of the reference factory are out of scope (SURVEY.md section 2, rows 10 / Surprise 3)."""
from __future__ import annotations

import torch

from .networks.unet_small import get_tomo_unet_small

_model_factory = {
    "unet": get_tomo_unet_small,
}


def create_model(arch, heads, head_conv, last_k=0, local_path=None):
    num_layers = int(arch[arch.find("_") + 1:]) if "_" in arch else 0
    arch = arch[:arch.find("_")] if "_" in arch else arch
    if arch not in _model_factory:
        raise KeyError(f"arch '{arch}' is outside cet_pick_b200's hot path (only unet_N is built)")
    get_model = _model_factory[arch]
    return get_model(num_layers=num_layers, heads=heads, head_conv=head_conv, last_k=last_k,
                     local_path=local_path)


def load_model(model, model_path, optimizer=None, resume=False, lr=None, lr_step=None, model_only=False):
    """Same checkpoint contract as the reference: dict with 'epoch' and 'state_dict'; a leading
    'module.' is stripped; shape mismatches and missing keys fall back to the model's own values."""
    start_epoch = 0
    checkpoint = torch.load(model_path, map_location=lambda storage, loc: storage)
    print("Loaded {}, epoch {}".format(model_path, checkpoint["epoch"]))
    state_dict_ = checkpoint["state_dict"]
    state_dict = {}
    for k in state_dict_:
        if k.startswith("module") and not k.startswith("module_list"):
            state_dict[k[7:]] = state_dict_[k]
        else:
            state_dict[k] = state_dict_[k]
    model_state_dict = model.state_dict()
    msg = ("If you see this, your model does not fully load the pre-trained weight. Please make sure "
           "you have correctly specified --arch xxx or set the correct --num_classes for your own dataset.")
    for k in state_dict:
        if k in model_state_dict:
            if state_dict[k].shape != model_state_dict[k].shape:
                print("Skip loading parameter {}, required shape{}, loaded shape{}. {}".format(
                    k, model_state_dict[k].shape, state_dict[k].shape, msg))
                state_dict[k] = model_state_dict[k]
        else:
            print("Drop parameter {}.".format(k) + msg)
    for k in model_state_dict:
        if k not in state_dict:
            print("No param {}.".format(k) + msg)
            state_dict[k] = model_state_dict[k]
    model.load_state_dict(state_dict, strict=False)

    if optimizer is not None and resume:
        if "optimizer" in checkpoint:
            optimizer.load_state_dict(checkpoint["optimizer"])
            start_epoch = checkpoint["epoch"]
            start_lr = lr
            for step in lr_step:
                if start_epoch >= step:
                    start_lr *= 0.1
            for param_group in optimizer.param_groups:
                param_group["lr"] = start_lr
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
