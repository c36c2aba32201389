"""``loops::algorithms::spmm::thread_mapped`` (reference
include/loops/algorithms/spmm/thread_mapped.cuh:55-80): C = A B with A in CSR and
B, C dense row-major (``matrix_t``, container/matrix.cuh -- here plain 2-D torch
tensors). Goes through ``loopsb_spmm_csr_f32``; synchronous unless ``sync=False``."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib
from ..container import csr_t


def thread_mapped(csr: csr_t, B: torch.Tensor, Cm: torch.Tensor, stream=None, sync=True):
    lib = _lib.load()
    stream = stream or torch.cuda.current_stream()
    for name, t, r in (("B", B, csr.cols), ("C", Cm, csr.rows)):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 2
                and t.is_contiguous() and t.shape[0] == r):
            raise ValueError(f"{name} must be a contiguous CUDA float32 [{r}, n] tensor")
    if B.shape[1] != Cm.shape[1]:
        raise ValueError("B and C must have the same number of columns")
    d = csr.layout().desc()
    _lib.check(lib.loopsb_spmm_csr_f32(C.byref(d), _lib.ptr(csr.values), _lib.ptr(csr.indices), _lib.ptr(B),
                                       _lib.ptr(Cm), csr.rows, csr.cols, int(B.shape[1]), _lib.stream_ptr(stream)),
               "loopsb_spmm_csr_f32")
    if sync:
        stream.synchronize()
