"""``loops::algorithms::spmv::*`` host entry points (reference
include/loops/algorithms/spmv/*.cuh), same names and argument meaning:

    merge_path_flat(csr, x, y, stream)   -> timer_t   merge_path_flat.cuh:96-139
    work_oriented(csr, x, y, stream)     -> None      work_oriented.cuh:102-120
    thread_mapped(csr, x, y, stream)     -> None      thread_mapped.cuh:69-91
    group_mapped(csr, x, y, stream)      -> None      group_mapped.cuh:72-104
    coo_thread_mapped(coo, x, y, stream) -> timer_t   coo_thread_mapped.cuh:61-89
    ell_thread_mapped(ell, x, y, stream) -> None      ell_thread_mapped.cuh:52-76
    ell_merge_path(ell, x, y, stream)    -> timer_t   ell_merge_path.cuh:76-126
    bcsr_thread_mapped(bcsr, x, y, stream) -> timer_t bcsr_thread_mapped.cuh:88-123
    csc_thread_mapped(csc, x, y, stream) -> timer_t   csc_thread_mapped.cuh:54-84
    dia_thread_mapped(dia, x, y, stream) -> timer_t   dia_thread_mapped.cuh:65-98
    flat_partitioned(csr, x, y, stream, K=8) -> timer_t  flat_partitioned.cuh:73-107
    original(csr, x, y, stream)          -> None      original.cuh:55-72

Each call goes straight through the C ABI (``loopsb_spmv_f32`` ...) into the
sm_100a kernels; like the reference wrappers they synchronise the stream before
returning unless ``sync=False`` is passed (benchmark loops, multi-GPU overlap).
Differences from the reference, both relaxations: ``y`` need not be zeroed
beforehand (it is fully overwritten), and the merge-path preprocess is cached
on the container instead of being rebuilt per call.
"""
from __future__ import annotations

import torch

from .. import _lib
from ..container import bcsr_t, coo_t, csc_t, csr_t, dia_t, ell_t


class timer_t:
    """CUDA-event stopwatch with the reference's accessor names (util/timer.hxx)."""

    def __init__(self, stream):
        self._stream = stream
        self._a = torch.cuda.Event(enable_timing=True)
        self._b = torch.cuda.Event(enable_timing=True)

    def start(self):
        self._a.record(self._stream)

    def stop(self):
        self._b.record(self._stream)
        self._b.synchronize()

    def milliseconds(self) -> float:
        return self._a.elapsed_time(self._b)

    def seconds(self) -> float:
        return self.milliseconds() / 1e3


def _check_vec(name, t, n):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32
            and t.is_contiguous() and t.numel() >= n):
        raise ValueError(f"{name} must be a contiguous CUDA float32 tensor with >= {n} elements")


def _run(container, schedule, values, cols, rows_idx, x, y, nrows, ncols, stream, sync, timed, tiled=None):
    lib = _lib.load()
    stream = stream or torch.cuda.current_stream()
    _check_vec("x", x, ncols)
    _check_vec("y", y, nrows)
    plan = container.plan(schedule, stream, tiled) if isinstance(container, csr_t) else container.plan(schedule, stream)
    timer = timer_t(stream) if (timed and sync) else None
    if timer:
        timer.start()
    _lib.check(lib.loopsb_spmv_f32(plan.handle, _lib.ptr(values), _lib.ptr(cols), _lib.ptr(rows_idx),
                                   _lib.ptr(x), _lib.ptr(y), nrows, ncols, _lib.stream_ptr(stream)),
               "loopsb_spmv_f32")
    if timer:
        timer.stop()
    elif sync:
        stream.synchronize()
    return timer


def _csr(csr: csr_t, schedule, x, y, stream, sync, timed=False, tiled=None):
    return _run(csr, schedule, csr.values, csr.indices, None, x, y, csr.rows, csr.cols, stream, sync, timed,
                tiled)


def merge_path_flat(csr: csr_t, x, y, stream=None, sync=True, tiled=None):
    """``tiled``: see ``csr_t.plan`` (None = LOOPSB_TILED / cost model, True =
    force the band-tiled kernel, False = the plain CSR merge-path kernel)."""
    return _csr(csr, _lib.SCHED_MERGE_PATH_FLAT, x, y, stream, sync, timed=True, tiled=tiled)


def work_oriented(csr: csr_t, x, y, stream=None, sync=True):
    _csr(csr, _lib.SCHED_WORK_ORIENTED, x, y, stream, sync)


def thread_mapped(csr: csr_t, x, y, stream=None, sync=True):
    _csr(csr, _lib.SCHED_THREAD_MAPPED, x, y, stream, sync)


def group_mapped(csr: csr_t, x, y, stream=None, sync=True):
    _csr(csr, _lib.SCHED_GROUP_MAPPED, x, y, stream, sync)


def coo_thread_mapped(coo: coo_t, x, y, stream=None, sync=True):
    return _run(coo, _lib.SCHED_THREAD_MAPPED, coo.values, coo.col_indices, coo.row_indices, x, y,
                coo.rows, coo.cols, stream, sync, timed=True)


def ell_thread_mapped(ell: ell_t, x, y, stream=None, sync=True):
    _run(ell, _lib.SCHED_THREAD_MAPPED, ell.values, ell.indices, None, x, y, ell.rows, ell.cols,
         stream, sync, timed=False)


def ell_merge_path(ell: ell_t, x, y, stream=None, sync=True):
    return _run(ell, _lib.SCHED_MERGE_PATH_FLAT, ell.values, ell.indices, None, x, y, ell.rows,
                ell.cols, stream, sync, timed=True)


def original(csr: csr_t, x, y, stream=None, sync=True):
    """The plain one-thread-per-row kernel of the reference (original.cuh): the
    same sequential-per-row arithmetic as thread_mapped, hence the same kernel."""
    _csr(csr, _lib.SCHED_THREAD_MAPPED, x, y, stream, sync)


def csc_thread_mapped(csc: csc_t, x, y, stream=None, sync=True):
    return _run(csc, _lib.SCHED_THREAD_MAPPED, csc.values, csc.indices, None, x, y, csc.rows, csc.cols,
                stream, sync, timed=True)


def flat_partitioned(csr: csr_t, x, y, stream=None, sync=True, K: int = 8):
    lib = _lib.load()
    stream = stream or torch.cuda.current_stream()
    _check_vec("x", x, csr.cols)
    _check_vec("y", y, csr.rows)
    plan = csr.flat_plan(K, stream)
    timer = timer_t(stream) if sync else None
    if timer:
        timer.start()
    _lib.check(lib.loopsb_spmv_f32(plan.handle, _lib.ptr(csr.values), _lib.ptr(csr.indices), None, _lib.ptr(x),
                                   _lib.ptr(y), csr.rows, csr.cols, _lib.stream_ptr(stream)), "loopsb_spmv_f32")
    if timer:
        timer.stop()
    return timer


def dia_thread_mapped(dia: dia_t, x, y, stream=None, sync=True):
    lib = _lib.load()
    stream = stream or torch.cuda.current_stream()
    _check_vec("x", x, dia.cols)
    _check_vec("y", y, dia.rows)
    timer = timer_t(stream) if sync else None
    if timer:
        timer.start()
    _lib.check(lib.loopsb_spmv_dia_f32(dia.rows, dia.cols, dia.stride, dia.num_diagonals,
                                       _lib.ptr(dia.diag_offsets), _lib.ptr(dia.values), _lib.ptr(x), _lib.ptr(y),
                                       _lib.stream_ptr(stream)), "loopsb_spmv_dia_f32")
    if timer:
        timer.stop()
    return timer


def bcsr_thread_mapped(bcsr: bcsr_t, x, y, stream=None, sync=True, repack=False):
    """fp32 blocks: one thread per block-row (the reference kernel's map).
    bf16 4x4 blocks: the tcgen05 tensor-core kernel (BASELINE config 4); its plan
    keeps a packed copy of the blocks (made on the first call and re-made when torch's
    version counters show an in-place write to ``bcsr.values`` / ``block_col_indices``;
    pass ``repack=True`` after a write torch cannot see, e.g. from a raw CUDA kernel).
    ``x`` must already be padded to ``num_block_cols * C`` (bcsr.padded_x)."""
    lib = _lib.load()
    stream = stream or torch.cuda.current_stream()
    n_pad = bcsr.num_block_cols * bcsr.C
    timer = timer_t(stream) if sync else None
    if timer:
        timer.start()
    if bcsr.values.dtype == torch.float32:
        _check_vec("x", x, n_pad)
        _check_vec("y", y, bcsr.rows)
        d = bcsr.layout().desc()
        import ctypes as C
        _lib.check(lib.loopsb_spmv_bcsr_f32(bcsr.R, bcsr.C, C.byref(d), _lib.ptr(bcsr.values),
                                            _lib.ptr(bcsr.block_col_indices), _lib.ptr(x), _lib.ptr(y),
                                            bcsr.rows, _lib.stream_ptr(stream)), "loopsb_spmv_bcsr_f32")
    elif bcsr.values.dtype == torch.bfloat16:
        if not (bcsr.R == 4 and bcsr.C == 4):
            raise _lib.LoopsbError(_lib.ERR_UNSUPPORTED, "bcsr_thread_mapped", "bf16 path is 4x4 only")
        if not (x.dtype == torch.bfloat16 and x.is_cuda and x.numel() >= n_pad):
            raise ValueError("x must be a CUDA bfloat16 tensor padded to num_block_cols*4")
        _check_vec("y", y, bcsr.rows)
        plan = bcsr.plan(_lib.SCHED_THREAD_MAPPED, stream)
        # once per matrix: the plan's packed copy of the block values (TMA-fed A tiles);
        # LOOPSB_BCSR_PACKED=0 keeps the kernel that reads the BCSR value array directly
        import os
        # (re-made by itself when torch saw an in-place write to the arrays since the last pack)
        key = (bcsr.values.data_ptr(), bcsr.values._version, bcsr.block_col_indices.data_ptr(),
               bcsr.block_col_indices._version)
        if os.environ.get("LOOPSB_BCSR_PACKED", "1") != "0" and \
                (repack or getattr(plan, "_packed_key", None) != key):
            _lib.check(lib.loopsb_plan_pack_bcsr4x4(plan.handle, _lib.ptr(bcsr.values),
                                                    _lib.ptr(bcsr.block_col_indices), _lib.stream_ptr(stream)),
                       "loopsb_plan_pack_bcsr4x4")
            plan._packed_key = key
        _lib.check(lib.loopsb_spmv_bcsr4x4_bf16(plan.handle, _lib.ptr(bcsr.values),
                                                _lib.ptr(bcsr.block_col_indices), _lib.ptr(x),
                                                _lib.ptr(y), bcsr.rows, _lib.stream_ptr(stream)),
                   "loopsb_spmv_bcsr4x4_bf16")
    else:
        raise _lib.LoopsbError(_lib.ERR_UNSUPPORTED, "bcsr_thread_mapped", f"dtype {bcsr.values.dtype}")
    if timer:
        timer.stop()
    return timer


# ---- the (schedule x layout) cells of BASELINE configs[2] without a kernel in the reference
# tree: schedule::setup<scheme, ..., layout> + the format's per-atom body (SURVEY 8 a17;
# loops_b200/csrc/spmv_generic.cu). Same argument meaning as the in-tree entry points.
def coo_group_mapped(coo: coo_t, x, y, stream=None, sync=True):
    return _run(coo, _lib.SCHED_GROUP_MAPPED, coo.values, coo.col_indices, coo.row_indices, x, y, coo.rows, coo.cols,
                stream, sync, timed=True)


def coo_work_oriented(coo: coo_t, x, y, stream=None, sync=True):
    return _run(coo, _lib.SCHED_WORK_ORIENTED, coo.values, coo.col_indices, coo.row_indices, x, y, coo.rows, coo.cols,
                stream, sync, timed=True)


def coo_merge_path(coo: coo_t, x, y, stream=None, sync=True):
    return _run(coo, _lib.SCHED_MERGE_PATH_FLAT, coo.values, coo.col_indices, coo.row_indices, x, y, coo.rows,
                coo.cols, stream, sync, timed=True)


def ell_group_mapped(ell: ell_t, x, y, stream=None, sync=True):
    return _run(ell, _lib.SCHED_GROUP_MAPPED, ell.values, ell.indices, None, x, y, ell.rows, ell.cols, stream, sync,
                timed=True)


def ell_work_oriented(ell: ell_t, x, y, stream=None, sync=True):
    return _run(ell, _lib.SCHED_WORK_ORIENTED, ell.values, ell.indices, None, x, y, ell.rows, ell.cols, stream, sync,
                timed=True)


# every cell of the {thread_mapped, group_mapped, work_oriented, merge_path_flat} x {csr, coo, ell} grid
CELLS = {
    ("csr", "thread_mapped"): thread_mapped, ("csr", "group_mapped"): group_mapped,
    ("csr", "work_oriented"): work_oriented, ("csr", "merge_path_flat"): merge_path_flat,
    ("coo", "thread_mapped"): coo_thread_mapped, ("coo", "group_mapped"): coo_group_mapped,
    ("coo", "work_oriented"): coo_work_oriented, ("coo", "merge_path_flat"): coo_merge_path,
    ("ell", "thread_mapped"): ell_thread_mapped, ("ell", "group_mapped"): ell_group_mapped,
    ("ell", "work_oriented"): ell_work_oriented, ("ell", "merge_path_flat"): ell_merge_path,
}

BY_NAME = {
    "merge_path_flat": merge_path_flat,
    "work_oriented": work_oriented,
    "thread_mapped": thread_mapped,
    "group_mapped": group_mapped,
}


def select_schedule(csr: csr_t, max_degree: int | None = None) -> str:
    """Name of the schedule ``loopsb_select_schedule`` picks for this matrix
    (thresholds measured by tools/heuristic_sweep.py). ``max_degree``
    None = compute it on the device when the matrix is there, else leave it unknown."""
    import ctypes as C
    if max_degree is None:
        if csr.values.is_cuda and csr.rows:
            from ..convert import csr_max_degree
            max_degree = csr_max_degree(csr)
        else:
            max_degree = -1
    out = C.c_int32(0)
    _lib.check(_lib.load().loopsb_select_schedule(csr.rows, csr.cols, csr.nnzs, int(max_degree), C.byref(out)),
               "loopsb_select_schedule")
    return {v: k for k, v in _lib.SCHEDULE_NAMES.items()}[int(out.value)]


def automatic(csr: csr_t, x, y, stream=None, sync=True):
    """SpMV with the schedule ``select_schedule`` picks (the choice is cached on the container)."""
    name = getattr(csr, "_auto_schedule", None)
    if name is None:
        name = csr._auto_schedule = select_schedule(csr)
    return BY_NAME[name](csr, x, y, stream=stream, sync=sync)


def spmv_f64(schedule: str, offsets, indices, values, x, y, rows: int, cols: int, stream=None, sync=True):
    """fp64 CSR SpMV (the reference builds its examples for double too).
    Tensors: offsets/indices int32, values/x/y float64, all on the device.
    ``thread_mapped`` is bit-equal to reference::spmv<double>; the other three
    schedule names share one merge-path kernel (loops_b200/csrc/spmv_f64.cu)."""
    import torch
    from ..layout import csr as csr_layout
    for name, t, dt in (("offsets", offsets, torch.int32), ("indices", indices, torch.int32),
                        ("values", values, torch.float64), ("x", x, torch.float64), ("y", y, torch.float64)):
        if t.dtype != dt or not t.is_cuda or not t.is_contiguous():
            raise ValueError(f"{name}: expected a contiguous {dt} device tensor")
    if offsets.numel() != rows + 1 or x.numel() < cols or y.numel() < rows:
        raise ValueError("shape mismatch")
    lay = csr_layout(offsets, rows, int(indices.numel())).desc()
    import ctypes as C
    _lib.check(_lib.load().loopsb_spmv_f64(C.byref(lay), _lib.SCHEDULE_NAMES[schedule], _lib.ptr(values),
                                           _lib.ptr(indices), _lib.ptr(x), _lib.ptr(y), rows, cols,
                                           _lib.stream_ptr(stream)), "loopsb_spmv_f64")
    if sync:
        torch.cuda.current_stream().synchronize() if stream is None else stream.synchronize()
