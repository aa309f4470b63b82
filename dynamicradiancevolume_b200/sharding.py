"""Multi-GPU host logic for the sharded gather (not in the reference, which is single-GPU).

Cache entries are independent units: after the replicated, deterministic allocation every rank holds the same
cell-ordered entry list; rank r lights the contiguous range ``shard_range(count, r, world)`` (64-entry aligned, so
a shard is a run of (cascade, brick) keys) and the lit SH payload is exchanged so that every rank can apply.

Two exchange paths:
  * fused (default on NVLink boxes): ``connect_peers`` maps every rank's entries buffer into every other rank
    (CUDA IPC); ``drv_light_caches`` then stores each finished entry to all peers from inside the gather
    epilogue. ``barrier`` (a one-word all-reduce) orders allocation / lighting / apply across ranks.
  * collective: ``exchange_entries`` broadcasts each rank's range with torch.distributed (NCCL on GPUs; gloo on
    CPU for the tests).
"""
from typing import List, Tuple

from . import _lib  # noqa: F401
from .renderer import shard_range


def shard_ranges(count: int, world: int) -> List[Tuple[int, int]]:
    return [shard_range(count, r, world) for r in range(world)]


def exchange_entries(entries, count: int, world: int, group=None):
    """All-gather of the lit ranges: after the call rows [0, count) of ``entries`` ([max, stride/4] float32,
    device tensor for NCCL / CPU tensor for gloo) are complete on every rank. Each rank must have filled its own
    range. Ranges are uneven (64-entry granularity), hence one broadcast per non-empty range."""
    import torch.distributed as dist
    for r, (b, e) in enumerate(shard_ranges(count, world)):
        if e > b:
            dist.broadcast(entries[b:e], src=r, group=group)
    return entries


def connect_peers(ctx, rank: int, world: int, group=None):
    """Fused path set-up: shard the context and map all peers' entries buffers (CUDA IPC over NVLink)."""
    import torch.distributed as dist
    ctx.set_shard(rank, world)
    handles = [None] * world
    dist.all_gather_object(handles, ctx.export_entries_ipc(), group=group)
    for r, h in enumerate(handles):
        if r != rank:
            ctx.import_peer_entries(r, h)


def connect_image_gather(ctx, rank: int, world: int, group=None):
    """Fused image gather set-up: rank 0 exports its context-owned RGBA16F target, everybody else maps it, so that
    ``drv_draw_frame(..., DRV_FRAME_GATHER_IMAGE)`` can store every rank's band there over NVLink."""
    import torch.distributed as dist
    handles = [None] * world
    dist.all_gather_object(handles, ctx.export_hdr_ipc() if rank == 0 else None, group=group)
    if rank != 0:
        ctx.import_peer_hdr(0, handles[0])


def barrier(word, group=None, ctx=None):
    """Stream-ordered cross-GPU barrier. With ``ctx`` (peers connected): flags in NVLink peer memory
    (``drv_peer_barrier``, a few microseconds); otherwise an NCCL all-reduce of one int32 on the current stream."""
    if ctx is not None:
        ctx.peer_barrier()
        return
    import torch.distributed as dist
    dist.all_reduce(word, group=group)


class SharedHostImage:
    """A host image every rank of one node can copy into with cudaMemcpyAsync: a POSIX shared-memory segment mapped
    by all ranks and page-locked (cudaHostRegister) in each process. Rank r copies its band of rows device -> host
    into it directly, so the end-to-end frame has no single-rank read-back funnel."""

    def __init__(self, name, shape, dtype, rank, world, group=None):
        import numpy as np
        import torch
        import torch.distributed as dist
        from multiprocessing import shared_memory
        self.rank = rank
        nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        if rank == 0:
            try:
                shared_memory.SharedMemory(name=name).unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=nbytes)
        if world > 1:
            dist.barrier(group=group)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name)
            try:  # only the creating rank may unlink the segment (Python's tracker would do it at exit of any rank)
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        buf = np.ndarray((nbytes,), dtype=np.uint8, buffer=self.shm.buf)
        self.tensor = torch.from_numpy(buf).view(dtype).view(*shape)
        self.registered = False
        if torch.cuda.is_available():  # page-lock the mapping in THIS process so async copies can target it
            err = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 0)
            if int(err) != 0:
                raise RuntimeError("cudaHostRegister failed: %s" % err)
            self.registered = True
        self._nbytes = nbytes
        if world > 1:
            dist.barrier(group=group)

    def close(self):
        import torch
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
        del self.tensor
        self.shm.close()
        if self.rank == 0:
            self.shm.unlink()
