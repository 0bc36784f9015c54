"""Multi-GPU sharding of the localisation hot path (one process per GPU).

Tomograms are independent (the reference loops them serially, cet_pick/test.py:82-85), so the list is
split by tomogram over the ranks and there is NO data-path collective; the only exchange is the
gather of the (n_local, K, 5) pick tensors.  Works over NCCL (CUDA tensors) and gloo (CPU tensors)."""
from __future__ import annotations

import torch
import torch.distributed as dist

HEAD_HALO = 3      # z reach of the 3-D head (feature_head.0, feature_head.2, hm: one slice each, unet_small.py:39-61)


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous share [first, first + count) of `n_items` for `rank`; the first n % world ranks get one more."""
    if world <= 0 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_items, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def gather_picks(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All ranks call this with their (count_r, K, 5) picks (shard_range order); every rank gets the
    (n_items, K, 5) tensor in tomogram order.  Unequal shares are padded to the largest share for the
    single all_gather and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    K, C = local.shape[1], local.shape[2]
    max_count = -(-n_items // world)
    buf = local.new_zeros((max_count, K, C))
    buf[:local.shape[0]] = local
    out = local.new_empty((world * max_count, K, C))
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    parts = []
    for r in range(world):
        _, cnt = shard_range(n_items, r, world)
        parts.append(out[r * max_count:r * max_count + cnt])
    return torch.cat(parts, 0)


def slab_range(depth: int, rank: int, world: int, halo: int = HEAD_HALO):
    """z-slab of ONE oversized volume for `rank`: core slices [z0, z1) it owns and the slices [lo, hi) it has
    to forward (core + recompute halo, clipped at the true volume boundary where zero padding is exact)."""
    z0, cnt = shard_range(depth, rank, world)
    z1 = z0 + cnt
    return z0, z1, max(0, z0 - halo), min(depth, z1 + halo)


def forward_z_sharded(forward_fn, volume_slab_fn, depth: int, group=None, halo: int = HEAD_HALO):
    """Heat-map of one volume whose z-slabs are spread over the ranks (the 2-D trunk is
    per-slice, the head needs +-3 slices, so each rank recomputes a 3-slice halo instead of exchanging features).

    volume_slab_fn(lo, hi) -> this rank's input slices (hi-lo, H, W) (only those are ever loaded);
    forward_fn(slab[, lo]) -> heat-map slices (hi-lo, h, w) of that slab; a two-argument function also receives the
                              absolute z of the slab's first slice (TomoConvUNet.z_origin: the head then adds its
                              partial sums in the whole-volume order and the slab result is bit-identical).
    Every rank returns the full (depth, h, w) heat-map: ONE all_gather of the core slabs (padded to the
    largest share), after which decode runs on identical data everywhere (picks bit-identical to 1 GPU)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    z0, z1, lo, hi = slab_range(depth, rank, world, halo)
    if z1 > z0:
        import inspect
        two = len(inspect.signature(forward_fn).parameters) >= 2
        hm_slab = forward_fn(volume_slab_fn(lo, hi), lo) if two else forward_fn(volume_slab_fn(lo, hi))
        core = hm_slab[z0 - lo:z0 - lo + (z1 - z0)].contiguous()
    else:
        core = None
    if world == 1:
        return core
    # shapes: every rank needs (h, w); ranks without slices learn them from the gathered meta
    meta = torch.tensor([core.shape[1], core.shape[2]] if core is not None else [0, 0], dtype=torch.int64,
                        device=core.device if core is not None else _default_device(group))
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    h, w = (int(v) for v in max((m for m in metas), key=lambda m: int(m[0])).tolist())
    max_cnt = -(-depth // world)
    dev = meta.device
    buf = torch.zeros((max_cnt, h, w), dtype=torch.float32, device=dev)
    if core is not None:
        buf[:z1 - z0] = core
    out = torch.empty((world * max_cnt, h, w), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = []
    for r in range(world):
        _, cnt = shard_range(depth, r, world)
        parts.append(out[r * max_cnt:r * max_cnt + cnt])
    return torch.cat(parts, 0)


def _default_device(group=None):
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


NMS_HALO = 1       # the (3,k,k) NMS window reaches one more slice in z (decode.py:27-33)


def merge_topk(scores: torch.Tensor, inds: torch.Tensor, K: int):
    """K best of the gathered candidates in the canonical order (score descending, linear index ascending) -- the
    order csrc/decode.cu uses.  scores / inds: 1-D tensors (CPU or CUDA).  Returns (scores[K], inds[K])."""
    o = torch.argsort(inds, stable=True)
    o = o[torch.argsort(scores[o], descending=True, stable=True)]
    o = o[:K]
    return scores[o], inds[o]


def local_candidates(forward_fn, volume_slab_fn, depth: int, K: int, kernel: int, rank: int, world: int):
    """This rank's K best NMS survivors among its own core slices of the volume: (scores[K], global linear indices[K]).
    The slab is forwarded with a (3 + 1)-slice recompute halo: 3 for the head, 1 more so that the 3x3x3 NMS of the
    core slices sees real neighbours (of the forwarded range core +- 4, the slices core +- 1 are exact)."""
    from .models.decode import _nms, _topk
    z0, z1, lo, hi = slab_range(depth, rank, world, HEAD_HALO + NMS_HALO)
    hm = forward_fn(volume_slab_fn(lo, hi), lo)                     # heat-map slices [lo, hi)
    h, w = int(hm.shape[1]), int(hm.shape[2])
    n_lo, n_hi = max(0, z0 - NMS_HALO), min(depth, z1 + NMS_HALO)
    sub = hm[n_lo - lo:n_hi - lo].contiguous()[None, None]
    nm = _nms(sub, kernel)
    nm[0, 0, :z0 - n_lo] = 0                                        # halo slices only lend their values to the windows
    nm[0, 0, z1 - n_lo:] = 0
    kk = min(K, sub.numel())
    sc, _, _, _, li = _topk(nm, kk)
    sc, li = sc.reshape(-1), li.reshape(-1) + n_lo * h * w          # local linear index -> global linear index
    if kk < K:
        sc = torch.cat([sc, sc.new_zeros(K - kk)])
        li = torch.cat([li, li.new_zeros(K - kk)])
    return sc.contiguous(), li.contiguous(), h, w


def gather_merge_candidates(sc: torch.Tensor, li: torch.Tensor, K: int, group=None):
    """the one exchange of the z-sharded decode: all_gather of every rank's (K scores, K global indices), then the
    merge-select in the canonical order.  Works over NCCL (CUDA tensors) and gloo (CPU tensors)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world > 1:
        gs = torch.empty(world * K, dtype=sc.dtype, device=sc.device)
        gi = torch.empty(world * K, dtype=li.dtype, device=li.device)
        dist.all_gather_into_tensor(gs, sc.contiguous(), group=group)
        dist.all_gather_into_tensor(gi, li.contiguous(), group=group)
    else:
        gs, gi = sc, li
    return merge_topk(gs, gi, K)


def rows_from_indices(scores: torch.Tensor, inds: torch.Tensor, depth: int, h: int, w: int) -> torch.Tensor:
    """(1, K, 5) pick rows of (score, global linear index) pairs with the reference's fp32 index arithmetic"""
    from . import _lib
    K = scores.numel()
    dets = torch.empty((1, K, 5), dtype=torch.float32, device=scores.device)
    _lib.check(_lib.lib().cetpick_rows_from_indices_f32(scores.contiguous().data_ptr(), inds.contiguous().data_ptr(), K,
                                                        depth, h, w, dets.data_ptr(), _lib.stream_ptr()), "rows_from_indices")
    return dets


def decode_z_sharded(forward_fn, volume_slab_fn, depth: int, K: int, kernel: int = 3, group=None):
    """SURVEY.md section 8e for ONE oversized volume: every rank forwards its z-slab with a recompute halo, decodes its
    OWN core slices (local top-K of the NMS map), and only K candidates per rank travel: one all_gather of (K scores,
    K global indices), then the merge-select.  No heat-map leaves its GPU.  -> (1, K, 5) picks, identical on every
    rank and bit-identical to the single-GPU tomo_decode of the whole volume (rows whose score is 0 excepted: filler
    when fewer than K peaks exist, whose indices torch.topk leaves unspecified).

    forward_fn(slab, lo) -> (hi-lo, h, w) float32 heat-map slices (sigmoid applied); volume_slab_fn(lo, hi) -> input."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    sc, li, h, w = local_candidates(forward_fn, volume_slab_fn, depth, K, kernel, rank, world)
    ms, mi = gather_merge_candidates(sc, li, K, group)
    return rows_from_indices(ms, mi, depth, h, w)
