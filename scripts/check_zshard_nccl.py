"""torchrun --nproc-per-node N scripts/check_zshard_nccl.py : one oversized volume, z-slabs over N GPUs (NCCL).
Every rank forwards only its slab (+ recompute halo).  Two exchanges are checked: all_gather of heat-map slabs with the
decode everywhere, and local decode + all_gather of K candidates per rank + merge-select.
Checks against the single-GPU whole-volume result on rank 0's device: heat-map and picks must be identical."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synthdata as synth                                   # noqa: E402
from cet_pick_b200.models.decode import tomo_decode               # noqa: E402
from cet_pick_b200.models.model import create_model               # noqa: E402
from cet_pick_b200.shard import decode_z_sharded, forward_z_sharded, slab_range     # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    D, H, W, K = 37, 128, 160, 200
    m = create_model("unet_4", {"hm": 1, "proj": 32}, 32, last_k=3)
    m.load_state_dict(synth.unet_state_dict_torch(317, 4))
    m = m.to(dev).eval()
    m.compute_proj, m.fuse_sigmoid = False, True
    vol = synth.tomogram_torch(D, H, W, seed=5, device=dev)        # every rank could load just its slab; here it is synthetic
    loaded = []

    def slab_fn(lo, hi):
        loaded.append((lo, hi))
        return vol[lo:hi]

    hm = forward_z_sharded(lambda s, lo: (setattr(m, "z_origin", lo), m(s[None])[-1]["hm"][0, 0], setattr(m, "z_origin", 0))[1], slab_fn, D)
    dets = tomo_decode(hm[None, None].contiguous(), kernel=3, K=K)
    whole = m(vol[None])[-1]["hm"][0, 0]
    ref = tomo_decode(whole[None, None].contiguous(), kernel=3, K=K)
    z0, z1, lo, hi = slab_range(D, rank, world)
    ok = bool((hm - whole).abs().max().item() <= 2e-7) and torch.equal(dets, ref) and loaded == [(lo, hi)]
    # SURVEY 8e proper: local decode of the own core slices, all_gather of K candidates per rank, merge-select
    fwd = lambda s, lo: (setattr(m, "z_origin", lo), m(s[None])[-1]["hm"][0, 0], setattr(m, "z_origin", 0))[1]
    dets2 = decode_z_sharded(fwd, lambda a, b: vol[a:b], D, K, 3)
    ok = ok and torch.equal(dets2.view(torch.int32), ref.view(torch.int32))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"z-sharded forward over {world} GPUs: slab of rank 0 = [{lo},{hi}) of {D}; "
              f"max |hm - whole| = {(hm - whole).abs().max().item():.2e}; picks identical = {torch.equal(dets, ref)}; candidate-merge picks identical = {torch.equal(dets2.view(torch.int32), ref.view(torch.int32))}; "
              f"{'OK' if flag.item() == 1 else 'MISMATCH'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
