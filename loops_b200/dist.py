"""Row-partitioned multi-GPU SpMV (BASELINE.json configs[4]).

The reference is single-GPU; this is new design fixed by the north star:
contiguous row ranges across the ranks of one box, GLOBAL column ids in every
shard, and exactly one collective per SpMV -- an all-gather of the dense x
shards (NCCL over NVLink when the tensors are on GPUs; gloo in the CPU tests of
the host logic). y stays sharded the same way x is, so the output of one SpMV is
the next one's x shard.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def row_range(rows: int, rank: int, world: int) -> tuple[int, int]:
    """Rank r owns rows [rows*r/world, rows*(r+1)/world)."""
    return (rows * rank) // world, (rows * (rank + 1)) // world


def shard_csr(off: np.ndarray, idx: np.ndarray, val: np.ndarray, rank: int, world: int):
    """Local CSR of one rank: offsets rebased to 0, column ids untouched."""
    rows = len(off) - 1
    r0, r1 = row_range(rows, rank, world)
    a, b = int(off[r0]), int(off[r1])
    return (off[r0:r1 + 1] - off[r0]).astype(np.int32), idx[a:b], val[a:b]


def nnz_imbalance(off: np.ndarray, world: int) -> float:
    """max over ranks of local nnz / mean local nnz (1.0 = perfectly even)."""
    rows = len(off) - 1
    per = [int(off[row_range(rows, r, world)[1]] - off[row_range(rows, r, world)[0]]) for r in range(world)]
    return max(per) / (sum(per) / world) if sum(per) else 1.0


def split_columns(off: torch.Tensor, idx: torch.Tensor, val: torch.Tensor, c0: int, c1: int):
    """Split a shard by column range: entries with c0 <= col < c1 (the columns of the
    rank's OWN x shard, rebased to 0) and the rest (global column ids). CSR order is
    kept inside both parts, so  y = A_own @ x_shard + A_rest @ x_full  and the first
    product does not have to wait for the all-gather. Works on CPU and CUDA tensors."""
    rows = off.numel() - 1
    nnz = idx.numel()
    own = (idx >= c0) & (idx < c1)
    rowid = torch.repeat_interleave(torch.arange(rows, device=idx.device), (off[1:] - off[:-1]).long(),
                                    output_size=nnz)
    cnt = torch.zeros(rows, dtype=torch.int64, device=idx.device).index_add_(0, rowid, own.long())
    own_off = torch.zeros(rows + 1, dtype=torch.int64, device=idx.device)
    torch.cumsum(cnt, 0, out=own_off[1:])
    rest_off = off.long() - own_off
    return ((own_off.to(torch.int32), (idx[own] - c0).to(torch.int32), val[own]),
            (rest_off.to(torch.int32), idx[~own].contiguous(), val[~own].contiguous()))


class DistSpMV:
    """y_shard = A_shard @ allgather(x_shard). `local` is this rank's csr_t
    (rows = its row range, cols = global); `x_full` is a persistent buffer."""

    def __init__(self, local, rows_global: int, group=None):
        self.local = local
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rows_global = rows_global
        if rows_global % self.world:
            raise ValueError("equal-sized shards need rows % world == 0 (plain all_gather)")
        dev = local.values.device
        self.x_full = torch.empty(local.cols, dtype=torch.float32, device=dev)

    def gather_x(self, x_shard: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            self.x_full.copy_(x_shard)
        else:
            dist.all_gather_into_tensor(self.x_full, x_shard, group=self.group)
        return self.x_full

    def __call__(self, x_shard: torch.Tensor, y_shard: torch.Tensor, sync: bool = False):
        from .algorithms import spmv
        self.gather_x(x_shard)
        spmv.merge_path_flat(self.local, self.x_full, y_shard, sync=sync)
        return y_shard
