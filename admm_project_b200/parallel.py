"""Host-side plumbing of the row-sharded solvers: one process per GPU, torch.distributed only as
the bootstrap channel (it carries the 128-byte NCCL id and gathers result shards); the data-path
collective -- ONE allreduce of [D'r ; D'dz ; D'u ; 8 scalars] per iteration -- is issued by
libadmm_b200 itself on the handle's stream.

Replaces the reference's PCT pool (gcp / parfor, admm.m:343-408; unwrappedadmm.m:45-74).  The row
partition is the reference's own balancing rule (errorcheck.m:249-259)."""
from __future__ import annotations

import numpy as np

from .engine import slicemaker


def dist_info():
    """(rank, world) of the default torch.distributed group, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def row_range(m, rank, world):
    """Rows [lo, hi) of `rank`: first mod(m, w) ranks get floor(m/w)+1 rows (errorcheck.m:249-259)."""
    sizes = slicemaker(m, world)
    lo = int(sum(sizes[:rank]))
    return lo, lo + int(sizes[rank])


def _device_key(engine):
    """(host, GPU uuid) of the engine's device: two ranks with the same key share one GPU."""
    import socket
    import torch
    try:
        uuid = str(torch.cuda.get_device_properties(engine.device).uuid)
    except Exception:  # older torch without .uuid
        uuid = "cuda:%d" % engine.device
    return socket.gethostname(), uuid


def attach_comm(engine):
    """Connect the engine to the ranks of the default torch.distributed group (which only carries the bootstrap
    bytes, over gloo or nccl).  One rank per GPU: rank 0 makes the NCCL id, everybody receives it
    (admm_b200_comm_init: NCCL for the Gram, peer mailboxes for the per-iteration messages).  When several ranks
    SHARE a device -- NCCL refuses that -- or ADMM_B200_TRANSPORT=ipc is set, the mailbox-only transport is used:
    the ranks exchange the CUDA IPC handles of their mailboxes here (admm_b200_comm_ipc_export / _attach)."""
    import os
    import torch.distributed as dist
    rank, world = dist_info()
    if world == 1:
        return rank, world
    if engine.nranks == world and engine.rank == rank:
        return rank, world
    keys = [None] * world
    dist.all_gather_object(keys, _device_key(engine))
    shared = len(set(keys)) < world
    if shared or os.environ.get("ADMM_B200_TRANSPORT", "") == "ipc":
        handles = [None] * world
        dist.all_gather_object(handles, engine.comm_ipc_export(rank, world))
        engine.comm_ipc_attach(handles)
        dist.barrier()      # nobody stores into a mailbox before every rank has mapped all of them
        return rank, world
    payload = [engine.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    engine.comm_init(rank, world, payload[0])
    return rank, world


def gather_rows(local, m_total):
    """Concatenate the per-rank row shards of a vector (or the rows x K history of one) in rank order."""
    rank, world = dist_info()
    if world == 1:
        return local
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(local))
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == m_total, (out.shape, m_total)
    return out


def shared_draw(draw):
    """Run `draw()` (a function that consumes the NumPy global RNG) on rank 0 and hand its result to every rank:
    a random init drawn per rank would make the global z0 / u0 a patchwork of unrelated streams."""
    rank, world = dist_info()
    if world == 1:
        return draw()
    import torch.distributed as dist
    payload = [draw() if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    return payload[0]
