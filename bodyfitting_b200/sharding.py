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


def exchange_halo(first_row, last_row, group=None):
    """Boundary-frame halo exchange of the temporal term: every rank sends its first frame's parameter row to
    rank-1 and its last to rank+1 and receives theirs.  Returns (prev_row | None, next_row | None); None at the
    two ends of the sequence.  NCCL send/recv on GPU tensors (NVLink within a box), gloo on CPU tensors."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None, None
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    first, last = first_row.contiguous(), last_row.contiguous()
    prev = torch.empty_like(first) if rank > 0 else None
    nxt = torch.empty_like(last) if rank + 1 < world else None
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, first, rank - 1, group), dist.P2POp(dist.irecv, prev, rank - 1, group)]
    if rank + 1 < world:
        ops += [dist.P2POp(dist.isend, last, rank + 1, group), dist.P2POp(dist.irecv, nxt, rank + 1, group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return prev, nxt


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


class HaloLink(object):
    """NVLink halo of the temporal term (BASELINE config 4): this rank's halo buffer plus the neighbouring ranks' buffers
    mapped into this process by CUDA IPC.  The optimiser kernel of rank r stores its first / last updated parameter row
    straight into the buffers of ranks r-1 / r+1 and raises a flag; their temporal kernel of the next iteration spins on
    it (csrc/bf_pack.cuh).  No host work per iteration, so the coupled fit is one CUDA graph like the uncoupled one.
    ``exchange_halo`` above stays as the tested host-driven fallback.

    ``HaloLink(group)`` is collective over the process group (one all_gather_object of the 64-byte IPC handles);
    ``HaloLink.local_chain(n)`` wires n links inside ONE process (shards of one GPU, used by the single-GPU test)."""

    def __init__(self, group=None, _local=None):
        import ctypes as C
        from . import _lib
        L = _lib.lib()
        self._L, self._opened = L, []
        buf = C.c_void_p()
        handle = C.create_string_buffer(L.bf_halo_handle_bytes())
        _lib.check(L.bf_halo_alloc(C.byref(buf), handle), 'bf_halo_alloc')
        self.buf, self.peer_prev, self.peer_next = buf.value, None, None
        self.rank, self.world = 0, 1
        self._group, self._group_ready = group, False
        if _local is not None:
            return
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        for attr, peer in (('peer_prev', self.rank - 1), ('peer_next', self.rank + 1)):
            if 0 <= peer < self.world:
                p = C.c_void_p()
                _lib.check(L.bf_halo_open(C.create_string_buffer(handles[peer], len(handles[peer])), C.byref(p)), 'bf_halo_open')
                setattr(self, attr, p.value)
                self._opened.append(p.value)
        dist.barrier(group=group)                       # every rank has mapped its neighbours before anyone stores into them
        self._group_ready = True

    def align(self):
        """Host barrier over the group before the FIRST run of a session, so that no rank's boundary warps spin while a
        neighbour is still preparing its model on the host (the steady state needs no host synchronisation)."""
        if self._group_ready:
            dist.barrier(group=self._group)

    @classmethod
    def local_chain(cls, n):
        links = [cls(_local=True) for _ in range(n)]
        for i, l in enumerate(links):
            l.rank, l.world = i, n
            l.peer_prev = links[i - 1].buf if i > 0 else None
            l.peer_next = links[i + 1].buf if i + 1 < n else None
        return links

    def close(self):
        for p in self._opened:
            self._L.bf_halo_close(p)
        self._opened = []
        if self.buf:
            self._L.bf_halo_free(self.buf)
            self.buf = None
