"""Multi-GPU plumbing: particles are sharded by contiguous or interleaved index sets, one process per GPU; the per-deposit
exchange is one all-reduce (sum) of the raw rho mesh over NCCL/NVLink (`gloo` on CPU for the host-logic tests).
Everything after the sum (ghost fold, neutralisation, Poisson, energy) runs redundantly and identically on
every rank, so no broadcast is needed.
"""
from __future__ import annotations

import numpy as np


def shard_range(nbpart_global: int, rank: int, world_size: int) -> tuple[int, int]:
    """contiguous slice [lo, hi) of the global particle index space owned by `rank`"""
    lo = nbpart_global * rank // world_size
    hi = nbpart_global * (rank + 1) // world_size
    return lo, hi


def interleaved_shard(nbpart_global: int, rank: int, world_size: int) -> tuple[int, int, int]:
    """(first, stride, count): `rank` owns the global particle indices first, first+stride, ... (count of them).
    For loads whose particle properties depend on the index -- the Landau load assigns |v| by index, src/landau.jl:35 --
    interleaving gives every rank the same mix; contiguous ranges (shard_range) give each rank a velocity band."""
    count = (nbpart_global - rank + world_size - 1) // world_size if nbpart_global > rank else 0
    return rank, world_size, count


class _CudaView:
    """expose a raw device pointer through __cuda_array_interface__ so torch can alias it without a copy"""

    def __init__(self, ptr: int, count: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                         "strides": None}


def nccl_unique_id() -> bytes:
    """128-byte NCCL unique id (uapic_nccl_unique_id): create on one rank, hand to the others"""
    import ctypes as C
    from ._lib import check, lib
    buf = C.create_string_buffer(128)
    check(lib().uapic_nccl_unique_id(buf))
    return buf.raw


def attach_nccl(session, group=None):
    """the native multi-GPU path: the library owns an NCCL communicator and reduces by itself (no Python in the step).
    torch.distributed (any backend) is used ONCE, to hand rank 0's unique id to the other ranks."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    session.init_nccl(box[0], world, rank)          # Session (2D UA path) and Session3D (fortran/uapic3d.f90) both have it


def attach_peer_exchange(session, group=None):
    """the fused path: no collective in the step.  The ranks' deposit meshes are summed inside the field-solve kernel straight
    out of each other's memory over NVLink (uapic_session_init_peers).  torch.distributed is used once, to gather the 64-byte
    IPC handles; Session.close() then synchronises the ranks before anything is unmapped."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    handles = [None] * world
    dist.all_gather_object(handles, session.peer_handle(), group=group)
    session.init_peers(b"".join(handles), world, rank)
    session._peer_group = group


def attach_torch_allreduce(session, group=None):
    """make `session` sum its raw rho mesh over the ranks of `group` with torch.distributed (NCCL), through the C ABI's
    callback hook.  The collective is issued on the stream handle the library passes (the session's stream), whatever
    torch's current stream is."""
    import torch
    import torch.distributed as dist

    cache = {}
    device = torch.device("cuda", torch.cuda.current_device())

    def fn(ptr, count, dtype, stream):
        key = (ptr, count, dtype)
        t = cache.get(key)
        if t is None:
            t = torch.as_tensor(_CudaView(ptr, count, "<f8" if dtype == 0 else "<i8"), device=device)
            cache[key] = t
        if stream and stream != torch.cuda.current_stream(device).cuda_stream:
            with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=device)):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return 0

    session.set_allreduce(fn)
    return fn


def allreduce_host(array: np.ndarray, group=None) -> np.ndarray:
    """sum a host mesh over ranks (gloo): used by the CPU tests of the sharding logic"""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(array))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.numpy().reshape(array.shape)
