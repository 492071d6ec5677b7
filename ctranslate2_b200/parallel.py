"""Host side of tensor parallelism: rank discovery and the one-time exchange of CUDA IPC handles.

The reference bootstraps with MPI (ScopedMPISetter, src/devices.cc:141-217: MPI_Comm_rank/size, MPI_Bcast of the
ncclUniqueId).  Here the launcher is torchrun / torch.distributed (one process per GPU) and the only thing the ranks
exchange on the host is one 64-byte cudaIpcMemHandle each; all data-path communication happens inside the kernels
over NVLink peer memory (csrc/kernels/tp_rows.cu)."""
from __future__ import annotations

from typing import List, Optional, Tuple


def default_rank_and_size(rank: Optional[int], size: Optional[int], group=None) -> Tuple[int, int]:
    if rank is not None and size is not None:
        return int(rank), int(size)
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("tensor_parallel=True needs torch.distributed to be initialised (torchrun), or explicit "
                           "tp_rank / tp_size")
    return dist.get_rank(group), dist.get_world_size(group)


def exchange_handles(mine: bytes, rank: int, size: int, group=None) -> List[bytes]:
    """All-gather of the per-rank handle, in rank order.  Works on any backend (gloo on CPU, nccl on GPU): the payload
    is a Python object."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("exchange_handles needs an initialised torch.distributed process group")
    if dist.get_world_size(group) != size or dist.get_rank(group) != rank:
        raise ValueError("tp_rank / tp_size do not match the process group")
    out: List[Optional[bytes]] = [None] * size
    dist.all_gather_object(out, mine, group=group)
    if any(h is None or len(h) != len(mine) for h in out):
        raise RuntimeError("handle exchange returned a malformed handle")
    if out[rank] != mine:
        raise RuntimeError("handle exchange is not in rank order")
    return [bytes(h) for h in out]


def shard_range(total: int, rank: int, size: int) -> Tuple[int, int]:
    """[begin, end) of a dimension split evenly over the ranks (src/models/model.cc:662-743)."""
    if total % size:
        raise ValueError("dimension %d is not divisible by %d ranks" % (total, size))
    per = total // size
    return rank * per, (rank + 1) * per
