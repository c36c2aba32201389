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


def ring_blocks(world: int, rank: int, groups) -> list[int]:
    """block_of_chunk[c] for ``loopsb_dist_create``'s column blocks: block 0 = the rank's
    own chunk, block g = the chunks of ranks rank+k for the ring shifts k of group g."""
    if sum(groups) != world - 1 or any(g < 1 for g in groups):
        raise ValueError("group sizes must be >= 1 and add up to world - 1")
    out = [0] * world
    k = 1
    for g, size in enumerate(groups):
        for _ in range(size):
            out[(rank + k) % world] = g + 1
            k += 1
    return out


def split_column_blocks(off: np.ndarray, idx: np.ndarray, val: np.ndarray, chunk_cols: int, block_of_chunk):
    """Host statement of ``loopsb_csr_split_columns_*``: the CSR blocks (offsets, indices,
    values) a shard is cut into, CSR order kept, GLOBAL column ids. Test / reference helper."""
    blk = np.asarray(block_of_chunk)[np.asarray(idx) // chunk_cols]
    rows = len(off) - 1
    rowid = np.repeat(np.arange(rows), np.diff(off))
    out = []
    for b in range(int(max(block_of_chunk)) + 1):
        m = blk == b
        o = np.zeros(rows + 1, np.int64)
        np.add.at(o, rowid[m] + 1, 1)
        out.append((np.cumsum(o).astype(np.int32), np.asarray(idx)[m], np.asarray(val)[m]))
    return out


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


def default_groups(world: int):
    """How the all-gather is phased by default (chunks per phase after the rank's own
    chunk, ring order): measured on B200 (profiles/): the phases hide behind the SpMV of
    the block before them once every block holds at least ~1/4 of the shard."""
    import os
    env = os.environ.get("LOOPSB_DIST_GROUPS")
    if env is not None:
        env = env.strip()
        return [] if env in ("", "0", "none") else [int(t) for t in env.split(",")]
    return {1: [], 2: [1], 4: [1, 2], 8: []}.get(world, [])


class DistPlan:
    """``loopsb_dist_*`` (include/loopsb.h): the multi-GPU step behind the C ABI.
    NCCL is driven by libloopsb200.so itself; torch.distributed (or any launcher) is
    only used to hand rank 0's NCCL id to the other ranks."""

    def __init__(self, local, world: int, rank: int, unique_id: bytes | None = None, groups=None, stream=None):
        import ctypes as C
        from . import _lib
        self._lib = lib = _lib.load()
        self.local = local            # keeps the arrays alive (borrowed when the shard is not split)
        self.world, self.rank = int(world), int(rank)
        if groups is None:
            groups = default_groups(world)
        self.groups = list(groups)
        g = (C.c_int32 * max(len(self.groups), 1))(*self.groups)
        h = C.c_void_p()
        idbuf = C.create_string_buffer(unique_id, _lib.DIST_ID_BYTES) if unique_id is not None else None
        _lib.check(lib.loopsb_dist_create(C.byref(h), idbuf, self.world, self.rank, local.rows, local.cols,
                                          local.nnzs, _lib.ptr(local.offsets), _lib.ptr(local.indices),
                                          _lib.ptr(local.values), g if self.groups else None, len(self.groups),
                                          _lib.stream_ptr(stream)), "loopsb_dist_create")
        self.handle = h

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from . import _lib
        buf = C.create_string_buffer(_lib.DIST_ID_BYTES)
        _lib.check(_lib.load().loopsb_dist_unique_id(buf), "loopsb_dist_unique_id")
        return buf.raw

    @classmethod
    def from_process_group(cls, local, groups=None, stream=None, group=None):
        """Rendezvous over an initialised torch.distributed group: rank 0 makes the id."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        box = [cls.unique_id() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(box, src=0, group=group)
        return cls(local, world, rank, box[0] if world > 1 else None, groups, stream)

    def __call__(self, x_shard: torch.Tensor, y_shard: torch.Tensor, stream=None):
        from . import _lib
        _lib.check(self._lib.loopsb_dist_spmv(self.handle, _lib.ptr(x_shard), _lib.ptr(y_shard),
                                              _lib.stream_ptr(stream)), "loopsb_dist_spmv")
        return y_shard

    def info(self) -> dict:
        import ctypes as C
        from . import _lib
        i = _lib.DistInfo()
        _lib.check(self._lib.loopsb_dist_info(self.handle, C.byref(i)), "loopsb_dist_info")
        d = {n: int(getattr(i, n)) for n in ("world", "rank", "local_rows", "num_cols", "num_blocks",
                                             "nccl_version", "local_nnz", "bytes", "graphs_cached")}
        d["transport"] = {0: "one ncclAllGather", 1: "nccl send/recv phases",
                          2: "copy-engine pulls over CUDA IPC (stream memory ops)"}[int(i.transport)]
        d["block_nnz"] = [int(v) for v in i.block_nnz][: d["num_blocks"]]
        d["groups"] = self.groups
        return d

    def x_full(self, cols: int) -> torch.Tensor:
        """The gathered x of the last step, copied into a new tensor."""
        import ctypes as C
        from . import _lib
        p = C.c_void_p()
        _lib.check(self._lib.loopsb_dist_x_full(self.handle, C.byref(p)), "loopsb_dist_x_full")
        out = torch.empty(cols, dtype=torch.float32, device=self.local.values.device)
        rt = C.CDLL("libcudart.so.12")
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        torch.cuda.synchronize()
        if rt.cudaMemcpy(out.data_ptr(), p, cols * 4, 3) != 0:     # cudaMemcpyDeviceToDevice
            raise RuntimeError("cudaMemcpy of the gathered x failed")
        return out

    def probe(self, on: bool = True):
        from . import _lib
        _lib.check(self._lib.loopsb_dist_probe(self.handle, 1 if on else 0), "loopsb_dist_probe")

    def probe_read(self) -> dict:
        import ctypes as C
        from . import _lib
        c, k = C.c_float(), C.c_float()
        blk = (C.c_float * 8)()
        _lib.check(self._lib.loopsb_dist_probe_read(self.handle, C.byref(c), C.byref(k), blk, 8),
                   "loopsb_dist_probe_read")
        nb = self.info()["num_blocks"]
        return {"comm_ms": float(c.value), "kernel_ms": float(k.value), "block_ms": [float(v) for v in blk][:nb]}

    def close(self):
        if getattr(self, "handle", None):
            self._lib.loopsb_dist_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
