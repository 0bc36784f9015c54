"""Exploration-step embedding inference from the command line, with the arguments of cet_pick/simsiam_test_hm_3d.py:136-195:

    python -m cet_pick_b200.simsiam_test_hm_3d simsiam3d --arch simsiam3d_18 --load_model model.pth --bbox 32 \\
        --dog 2.5,5 --gauss 0.8 --compress --test_img_txt test_images.txt --exp_id run

Every tomogram of the tab-separated list goes MRC -> GPU pre-processing (utils/loader.py) -> difference-of-Gaussians
candidate generator (utils/image.get_potential_coords_pyramid) -> z-slab sums around the kept candidates
(utils/image.extract_subvols): the `test` split of the reference's dataset
(datasets/tomo_pre_proj_angle_select_new3d_vol.py:181-246).  The patches are quantised and normalised like
PrefetchDatasetProj (simsiam_test_hm_3d.py:31-61: ToPILImage -> ToTensor -> Normalize(mean, std of all patches)), embedded
in batches of 256 by `model.forward_test` (the SimSiam encoder in libcetpick_sm100a.so) and written to
`<save_dir>/all_output_info.npz` with the reference's keys: proj, pred, name, coords, subvol."""
from __future__ import annotations

import os

import numpy as np
import torch

from .models.model import create_model, load_model
from .opts import opts
from .test import read_image_list, write_opt_file
from .utils import loader
from .utils.image import extract_subvols, get_potential_coords_pyramid

BATCH = 256            # simsiam_test_hm_3d.py:151: the DataLoader's batch size, not opt.batch_size
SLAB = 3               # :150: Dataset(opt, split, (3, opt.bbox, opt.bbox), sigma1=opt.dog)


def keep_candidates(positions, shape, bbox):
    """datasets/tomo_pre_proj_angle_select_new3d_vol.py:207: candidates whose window stays inside the plane
    (x strictly, y inclusively, margins bbox // 1.8 as the reference's float floor division).  positions: (n, 3) x, y, z."""
    D, H, W = shape
    positions = np.asarray(positions).reshape(-1, 3)
    mx, my = bbox // 1.8, bbox // 1.8
    x, y = positions[:, 0], positions[:, 1]
    return positions[(x > mx) & (x < W - mx) & (y >= my) & (y <= H - my)]


def candidate_patches(rec, opt):
    """one tomogram -> (patches (n, 1, bbox, bbox) float32 on the device, coords (n, 3) int32 x, y, z)"""
    _, positions = get_potential_coords_pyramid(rec, sigmas=opt.dog)
    coords = keep_candidates(positions, tuple(rec.shape), opt.bbox).astype(np.int32)
    patches = extract_subvols(rec, coords, [SLAB, opt.bbox, opt.bbox])
    return patches, coords


def normalise_patches(patches, mean, std):
    """simsiam_test_hm_3d.py:44-51 on a batch: ToPILImage (x 255, truncated to a byte) -> ToTensor (/ 255) -> Normalize"""
    q = patches.mul(255).to(torch.uint8).to(torch.float32).div(255)
    return q.sub_(mean).div_(std)


def embed(model, patches, coords, names):
    """the loop of :162-176 over one list of patches; -> dict with the reference's keys"""
    mean, std = patches.mean(), patches.std()            # :243-244: over every element of every patch, unbiased
    proj, pred, subvol = [], [], []
    for i in range(0, patches.shape[0], BATCH):
        x = normalise_patches(patches[i:i + BATCH], mean, std)
        ret = model.forward_test(x)
        proj.append(ret["proj"].detach().cpu().numpy())
        pred.append(ret["pred"].detach().cpu().numpy())
        subvol.append(x.cpu().numpy())
    return {"proj": np.concatenate(proj, axis=0), "pred": np.concatenate(pred, axis=0), "name": np.asarray(names),
            "coords": np.asarray(coords), "subvol": np.concatenate(subvol, axis=0)}


def test(opt):
    first_gpu = int(str(getattr(opt, "gpus_str", "0")).split(",")[0] or 0)
    if first_gpu >= 0:
        torch.cuda.set_device(first_gpu)                  # the reference's CUDA_VISIBLE_DEVICES = opt.gpus_str (:137)
    print(opt)
    write_opt_file(opt)
    model = create_model(opt.arch, opt.heads, opt.head_conv)
    model = load_model(model, opt.load_model)
    model = model.cuda().eval()
    path = opt.test_img_txt if os.path.isabs(opt.test_img_txt) else os.path.join(opt.data_dir, opt.test_img_txt)
    patches, coords, names = [], [], []
    for name, rec_path in read_image_list(path):
        rec = loader.load_tomos_from_list([name], [rec_path], compress=opt.compress, denoise=opt.gauss)[name]
        p, c = candidate_patches(rec, opt)
        patches.append(p)
        coords.append(c)
        names += [name] * len(c)
        print(f"{name}: {len(c)} candidates")
    patches = torch.cat(patches, dim=0)
    if patches.shape[0] == 0:
        raise RuntimeError("no candidate passed the filters (the reference fails on the empty stack as well)")
    out = embed(model, patches, np.concatenate(coords, axis=0), names)
    out_file = os.path.join(opt.save_dir, "all_output_info.npz")
    print("opt.save_dir", opt.save_dir)
    np.savez(out_file, **out)
    return out_file


if __name__ == "__main__":
    test(opts().init())
