"""Row-block sharding of C|D across ranks: one process per GPU, every rank stages
the same A|B (replicated), eliminates the rows r with r % world == rank, and the reduced rows are
all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) before the replicated echelon.

The C engine calls back into `make_allgather(...)` with device pointers; the functions below are
also used on host tensors by the gloo tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(n_bot: int, rank: int, world: int) -> range:
    """Rows of C|D owned by `rank` (block-cyclic with block 1: costs are skewed along the row order)."""
    return range(rank, n_bot, world)


def rows_per_rank(n_bot: int, world: int) -> int:
    """Padded block height: every rank contributes the same number of rows to the all-gather."""
    return (n_bot + world - 1) // world


def gathered_row_of(r: int, n_bot: int, world: int) -> int:
    """Index of original row r inside the gathered [world * rows_per_rank] block."""
    return (r % world) * rows_per_rank(n_bot, world) + r // world


class _DevMem:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def make_allgather(group=None):
    """Callback for LinalgEngine.set_allgather: all-gather `nbytes` per rank between device buffers,
    ordered on the engine's stream."""
    def gather(send_ptr, recv_ptr, nbytes, stream_ptr):
        world = dist.get_world_size(group)
        send = torch.as_tensor(_DevMem(send_ptr, nbytes), device="cuda")
        recv = torch.as_tensor(_DevMem(recv_ptr, nbytes * world), device="cuda")
        ext = torch.cuda.ExternalStream(stream_ptr)
        with torch.cuda.stream(ext):
            dist.all_gather_into_tensor(recv, send, group=group)
    return gather


def make_broadcast(group=None):
    """Callback for LinalgEngine.set_broadcast: in-place broadcast of a device buffer from `root`
    (NCCL over NVLink), ordered on the engine's stream."""
    def bcast(buf_ptr, nbytes, root, stream_ptr):
        buf = torch.as_tensor(_DevMem(buf_ptr, nbytes), device="cuda")
        ext = torch.cuda.ExternalStream(stream_ptr)
        with torch.cuda.stream(ext):
            dist.broadcast(buf, src=root, group=group)
    return bcast


def allgather_rows(local_block: torch.Tensor, group=None) -> torch.Tensor:
    """Host/any-backend variant: local_block is [rows_per_rank, n_nonpiv]; returns the gathered
    [world * rows_per_rank, n_nonpiv] block in the layout the engine's echelon reads."""
    world = dist.get_world_size(group)
    out = [torch.empty_like(local_block) for _ in range(world)]
    dist.all_gather(out, local_block.contiguous(), group=group)
    return torch.cat(out, dim=0)
