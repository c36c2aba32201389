"""ctypes binding of the C ABI in include/loopsb.h (libloopsb200.so).

The shared object is built in-tree by ``__graft_entry__.build()`` /
``make -C loops_b200/csrc``. There is no fallback of any kind: if the library
is missing, ``load()`` raises, and every compute entry point returns an error
status when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libloopsb200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_ALLOC = range(5)

# reference schedule::algorithms_t order (schedule.hxx:26-32)
SCHED_MERGE_PATH_FLAT, SCHED_WORK_ORIENTED, SCHED_THREAD_MAPPED, SCHED_GROUP_MAPPED = range(4)
SCHEDULE_NAMES = {
    "merge_path_flat": SCHED_MERGE_PATH_FLAT,
    "work_oriented": SCHED_WORK_ORIENTED,
    "thread_mapped": SCHED_THREAD_MAPPED,
    "group_mapped": SCHED_GROUP_MAPPED,
}
(LAYOUT_CSR, LAYOUT_COO, LAYOUT_ELL, LAYOUT_BCSR, LAYOUT_CSC, LAYOUT_DIA,
 LAYOUT_FLAT) = range(7)


class LoopsbError(RuntimeError):
    def __init__(self, status: int, where: str, detail: str):
        self.status = status
        super().__init__(f"{where}: status {status} ({detail})")


class LayoutDesc(C.Structure):
    """loopsb_layout_t"""
    _fields_ = [
        ("kind", C.c_int32),
        ("num_tiles", C.c_int32),
        ("num_atoms", C.c_int32),
        ("pitch", C.c_int32),
        ("offsets", C.c_void_p),
    ]


class PlanInfo(C.Structure):
    """loopsb_plan_info_t"""
    _fields_ = [
        ("schedule", C.c_int32),
        ("layout_kind", C.c_int32),
        ("threads_per_block", C.c_int32),
        ("items_per_thread", C.c_int32),
        ("num_merge_tiles", C.c_int64),
        ("grid_blocks", C.c_int32),
        ("cta_threads", C.c_int32),
        ("launches_per_spmv", C.c_int32),
        ("smem_bytes", C.c_int32),
        ("workspace_bytes", C.c_int64),
    ]


class TiledInfo(C.Structure):
    """loopsb_tiled_info_t"""
    _fields_ = [(n, C.c_int32) for n in ("nb", "q", "warps", "cb", "xb", "es", "rb", "rw", "cq", "nband",
                                         "grid_blocks", "cta_threads", "smem_bytes", "long_steps")] + \
               [(n, C.c_int64) for n in ("total_steps", "real_entries", "pad_entries", "flagged_entries",
                                         "flagged_steps", "bytes")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class DistInfo(C.Structure):
    """loopsb_dist_info_t"""
    _fields_ = [(n, C.c_int32) for n in ("world", "rank", "local_rows", "num_cols", "num_blocks", "nccl_version",
                                         "transport", "graphs_cached")] + \
               [("local_nnz", C.c_int64), ("block_nnz", C.c_int64 * 8), ("bytes", C.c_int64)]


TILE_FORCE = 1
DIST_ID_BYTES = 128

# every symbol include/loopsb.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "loopsb_version": (C.c_int, []),
    "loopsb_status_string": (C.c_char_p, [C.c_int]),
    "loopsb_last_error": (C.c_char_p, []),
    "loopsb_device_info": (C.c_int, [C.POINTER(C.c_int32)] * 3),
    "loopsb_plan_create": (C.c_int, [C.POINTER(_P), C.POINTER(LayoutDesc), C.c_int, _P]),
    "loopsb_plan_destroy": (C.c_int, [_P]),
    "loopsb_plan_info": (C.c_int, [_P, C.POINTER(PlanInfo)]),
    "loopsb_plan_hint_x_bytes": (C.c_int, [_P, C.c_int64]),
    "loopsb_plan_invalidate": (C.c_int, [_P]),
    "loopsb_plan_tile_breakeven": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int64)]),
    "loopsb_plan_merge_coords_host": (C.c_int, [_P, _P, C.c_int64]),
    "loopsb_plan_debug_phases_host": (C.c_int, [_P, _P, C.c_int64]),
    "loopsb_plan_probe_begin": (C.c_int, [_P, C.c_int32]),
    "loopsb_plan_probe_collect": (C.c_int, [_P, _P, C.c_int32, C.POINTER(C.c_int32)]),
    "loopsb_plan_tile_csr": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "loopsb_plan_untile": (C.c_int, [_P]),
    "loopsb_plan_tiled_download": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "loopsb_plan_tiled_info": (C.c_int, [_P, C.POINTER(TiledInfo)]),
    "loopsb_tiled_image_build_host": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, C.POINTER(C.c_int32 * 6),
                                                C.POINTER(_P)]),
    "loopsb_tiled_image_info": (C.c_int, [_P, C.POINTER(TiledInfo)]),
    "loopsb_tiled_image_arrays": (C.c_int, [_P] + [C.POINTER(_P)] * 6),
    "loopsb_tiled_image_free": (C.c_int, [_P]),
    "loopsb_spmv_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "loopsb_spmv_bcsr_f32": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(LayoutDesc), _P, _P, _P, _P, C.c_int32, _P]),
    "loopsb_spmv_dia_f32": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P, _P, _P]),
    "loopsb_spmm_csr_f32": (C.c_int, [C.POINTER(LayoutDesc), _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "loopsb_plan_pack_bcsr4x4": (C.c_int, [_P, _P, _P, _P]),
    "loopsb_spmv_bcsr4x4_bf16": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, _P]),
    "loopsb_spmv_csr_host_f32": (C.c_int, [C.c_int, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, C.POINTER(C.c_float)]),
    "loopsb_emit_schedule": (C.c_int, [C.POINTER(LayoutDesc), C.c_int, C.c_int32, C.c_int32, C.c_int32,
                                       _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int64, _P]),
    "loopsb_work_oriented_grid": (C.c_int, [C.POINTER(C.c_int32)]),
    "loopsb_spmv_f64": (C.c_int, [C.POINTER(LayoutDesc), C.c_int, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "loopsb_spmv_layout_f64": (C.c_int, [C.POINTER(LayoutDesc), C.c_int, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "loopsb_spmv_dia_f64": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P, _P, _P]),
    "loopsb_spmv_bcsr_f64": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(LayoutDesc), _P, _P, _P, _P, C.c_int32, _P]),
    "loopsb_select_schedule": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(C.c_int32)]),
    # format conversions on the device
    "loopsb_csr_to_coo": (C.c_int, [C.c_int32, C.c_int64, _P, _P, _P]),
    "loopsb_coo_to_csr": (C.c_int, [C.c_int32, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "loopsb_csr_to_csc": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "loopsb_csr_max_degree": (C.c_int, [C.c_int32, _P, C.POINTER(C.c_int32), _P]),
    "loopsb_csr_to_ell": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "loopsb_csr_to_bcsr_count": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, _P,
                                           C.POINTER(C.c_int64), _P]),
    "loopsb_csr_to_bcsr_fill": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, _P,
                                          C.c_int64, _P, _P, C.c_int32, _P]),
    "loopsb_csr_to_dia_count": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, _P, _P, C.POINTER(C.c_int32), _P]),
    "loopsb_csr_to_dia_fill": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, _P, _P, _P, C.c_int32, _P, _P, _P]),
    "loopsb_csr_split_columns_count": (C.c_int, [C.c_int32, C.c_int64, _P, _P, C.c_int32, C.c_int32, _P, C.c_int32,
                                                 _P, _P, _P]),
    "loopsb_csr_split_columns_fill": (C.c_int, [C.c_int32, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, _P,
                                                C.c_int32, _P, _P, _P, _P, _P]),
    # y += A x and the row-partitioned multi-GPU path
    "loopsb_spmv_acc_f32": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "loopsb_dist_unique_id": (C.c_int, [_P]),
    "loopsb_dist_create": (C.c_int, [C.POINTER(_P), _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                     _P, _P, _P, _P, C.c_int32, _P]),
    "loopsb_dist_spmv": (C.c_int, [_P, _P, _P, _P]),
    "loopsb_dist_info": (C.c_int, [_P, C.POINTER(DistInfo)]),
    "loopsb_dist_x_full": (C.c_int, [_P, C.POINTER(_P)]),
    "loopsb_dist_probe": (C.c_int, [_P, C.c_int32]),
    "loopsb_dist_probe_read": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), _P, C.c_int32]),
    "loopsb_dist_destroy": (C.c_int, [_P]),
}

_lib = None


def load() -> C.CDLL:
    """Load libloopsb200.so (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C loops_b200/csrc`. loops-b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, where: str) -> None:
    if status != OK:
        lib = load()
        detail = lib.loopsb_last_error().decode(errors="replace")
        raise LoopsbError(status, where, detail or lib.loopsb_status_string(status).decode())


def ptr(t) -> int | None:
    """Device/host address of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr() if t.numel() else None
    return t.ctypes.data if t.size else None


def stream_ptr(stream=None):
    """cudaStream_t of a torch stream (default: torch's current stream)."""
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)
