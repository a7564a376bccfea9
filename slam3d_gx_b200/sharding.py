"""Multi-GPU plumbing of the registration path: frame pairs are independent units (reference
src/GraphicEnd.cpp:729-761 runs the loop-closure candidates one after another in a `for`), so a batch is
block-partitioned over the ranks with no data-path exchange; the only collective is an all-gather of the
fixed-size result records.  Works with any torch.distributed backend (nccl on GPUs, gloo on CPU)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


def partition(n_items: int, world: int, rank: int) -> range:
    """Contiguous block partition: item i belongs to rank floor(i * world / n_items)."""
    lo = (rank * n_items + world - 1) // world
    hi = ((rank + 1) * n_items + world - 1) // world
    return range(lo, hi)


def records_to_bytes(results) -> np.ndarray:
    """ctypes array of s3d_result (or list of them) -> uint8 array of len(results) * RESULT_BYTES."""
    n = len(results)
    out = np.empty(n * _abi.RESULT_BYTES, np.uint8)
    for i in range(n):
        out[i * _abi.RESULT_BYTES:(i + 1) * _abi.RESULT_BYTES] = np.frombuffer(bytes(results[i]), dtype=np.uint8)
    return out


def bytes_to_records(buf: np.ndarray):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    n = buf.size // _abi.RESULT_BYTES
    arr = (_abi.Result * n)()
    C.memmove(arr, buf.ctypes.data, n * _abi.RESULT_BYTES)
    return [_abi.result_to_dict(arr[i]) for i in range(n)]


def gather_results(local_results, n_total: int, dist, device="cpu"):
    """All-gather the ranks' result records into the global pair order.  `local_results` are the records of
    partition(n_total, world, rank), in order.  Uneven shards are padded to the largest shard."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    per = max(len(partition(n_total, world, r)) for r in range(world))
    send = torch.zeros(per * _abi.RESULT_BYTES, dtype=torch.uint8)
    mine = records_to_bytes(local_results)
    send[:mine.size] = torch.from_numpy(mine)
    send = send.to(device)
    recv = torch.empty(world * per * _abi.RESULT_BYTES, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy()
    out = []
    for r in range(world):
        k = len(partition(n_total, world, r))
        out.extend(bytes_to_records(recv[r * per * _abi.RESULT_BYTES:(r * per + k) * _abi.RESULT_BYTES]))
    return out
