"""Frame sharding across the GPUs of one box (one process per GPU, torch.distributed).

Frames (and subjects) are independent optimisation problems in the reference (a fresh SMPLify
and a fresh Adam per frame: smplify/body_fitting.py:82-91, apps/genebody_fitting.py:184), so the
fit needs no collective: rank r owns the contiguous range [r*ceil(F/G), ...) and the only
communication is the final gather of the fitted parameters (NCCL on GPUs, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def frame_range(n_frames, rank, world):
    """Contiguous, balanced split: the first (n_frames % world) ranks get one extra frame."""
    base, rem = divmod(int(n_frames), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_frames(local, n_frames, group=None):
    """All-gather per-frame rows [n_local, ...] from every rank into [n_frames, ...] in frame order.
    Ragged ranges are padded to the largest shard for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [frame_range(n_frames, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    assert local.shape[0] == hi - lo, 'rank %d holds %d rows, expected %d' % (rank, local.shape[0], hi - lo)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:hi - lo] = local
    out = torch.empty((world * nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return torch.cat([out[r * nmax: r * nmax + (sizes[r][1] - sizes[r][0])] for r in range(world)], dim=0)
