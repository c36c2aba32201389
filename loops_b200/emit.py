"""Schedule index-stream emission through ``loopsb_emit_schedule`` (parity
instrument; see include/loopsb.h for the record format)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


@dataclass
class Stream:
    visitor: np.ndarray
    step: np.ndarray
    tile: np.ndarray
    visits: np.ndarray
    commit: np.ndarray | None = None        # work_oriented
    map: np.ndarray | None = None           # work_oriented [N,4]
    dense_tile: np.ndarray | None = None    # merge_path_flat [M*TPB*IPT]
    dense_atom: np.ndarray | None = None
    dense_emit: np.ndarray | None = None
    thread_start: np.ndarray | None = None  # merge_path_flat [M*TPB,2]


def work_oriented_grid() -> int:
    g = C.c_int32()
    _lib.check(_lib.load().loopsb_work_oriented_grid(C.byref(g)), "loopsb_work_oriented_grid")
    return g.value


def emit_schedule(layout, schedule: int, grid_blocks: int = 0, threads_per_block: int = 128,
                  items_per_thread: int = 8) -> Stream:
    lib = _lib.load()
    dev = torch.device("cuda")
    A = layout.num_atoms()
    T = layout.num_tiles()
    mk = lambda n, fill: torch.full((max(int(n), 1),), fill, dtype=torch.int32, device=dev)
    visitor, step, tile, visits = mk(A, -1), mk(A, -1), mk(A, -1), mk(A, 0)
    extra_a = extra_b = d_tile = d_atom = d_emit = None
    dense_len = 0
    if schedule == _lib.SCHED_WORK_ORIENTED:
        extra_a, extra_b = mk(A, -1), mk(grid_blocks * 128 * 4, -1)
    if schedule == _lib.SCHED_MERGE_PATH_FLAT:
        I = threads_per_block * items_per_thread
        M = (T + A + I - 1) // I
        dense_len = M * I
        d_tile, d_atom, d_emit = mk(dense_len, -1), mk(dense_len, -1), mk(dense_len, -1)
        extra_b = mk(M * threads_per_block * 2, -1)
    desc = layout.desc()
    _lib.check(lib.loopsb_emit_schedule(C.byref(desc), schedule, grid_blocks, threads_per_block,
                                        items_per_thread, _lib.ptr(visitor), _lib.ptr(step),
                                        _lib.ptr(tile), _lib.ptr(visits), _lib.ptr(extra_a),
                                        _lib.ptr(extra_b), _lib.ptr(d_tile), _lib.ptr(d_atom),
                                        _lib.ptr(d_emit), dense_len, _lib.stream_ptr()),
               "loopsb_emit_schedule")
    h = lambda t, n: None if t is None else t.cpu().numpy()[: int(n)]
    s = Stream(h(visitor, A), h(step, A), h(tile, A), h(visits, A))
    if schedule == _lib.SCHED_WORK_ORIENTED:
        s.commit = h(extra_a, A)
        s.map = h(extra_b, grid_blocks * 128 * 4).reshape(-1, 4)
    if schedule == _lib.SCHED_MERGE_PATH_FLAT:
        s.dense_tile, s.dense_atom, s.dense_emit = h(d_tile, dense_len), h(d_atom, dense_len), h(d_emit, dense_len)
        s.thread_start = h(extra_b, (dense_len // items_per_thread) * 2).reshape(-1, 2)
    return s
