"""Format conversions on the device: the host-side
mirror of the ``loopsb_csr_to_*`` / ``loopsb_coo_to_csr`` entry points. The
reference converts with host loops (container/ell.hxx:113-145, bcsr.hxx:111-194,
dia.hxx:135-188) or thrust sorts of COO triples (coo.hxx:87-98, csr.hxx:86-94,
csc.hxx:86-108); these run as CUDA kernels over arrays already in HBM and return
containers whose arrays are bit-equal to the reference converters' output.
torch only allocates the outputs."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .container import bcsr_t, coo_t, csc_t, csr_t, dia_t, ell_t


def _need_cuda(csr):
    if not csr.values.is_cuda:
        raise ValueError("device conversions take a device-resident matrix (there is no CPU fallback)")


def csr_to_coo(csr: csr_t, stream=None) -> coo_t:
    """Row id of every atom; column ids and values are shared with the CSR."""
    _need_cuda(csr)
    rows_of = torch.empty(csr.nnzs, dtype=torch.int32, device=csr.values.device)
    _lib.check(_lib.load().loopsb_csr_to_coo(csr.rows, csr.nnzs, _lib.ptr(csr.offsets), _lib.ptr(rows_of),
                                             _lib.stream_ptr(stream)), "loopsb_csr_to_coo")
    return coo_t.from_tensors(csr.rows, csr.cols, rows_of, csr.indices, csr.values)


def coo_to_csr(coo: coo_t, stream=None) -> csr_t:
    """Sort by (row, col) and compress the rows; the input order is free."""
    _need_cuda(coo)
    dev = coo.values.device
    off = torch.empty(coo.rows + 1, dtype=torch.int32, device=dev)
    idx = torch.empty(coo.nnzs, dtype=torch.int32, device=dev)
    val = torch.empty(coo.nnzs, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().loopsb_coo_to_csr(coo.rows, coo.nnzs, _lib.ptr(coo.row_indices), _lib.ptr(coo.col_indices),
                                             _lib.ptr(coo.values), _lib.ptr(off), _lib.ptr(idx), _lib.ptr(val),
                                             _lib.stream_ptr(stream)), "loopsb_coo_to_csr")
    return csr_t.from_tensors(coo.rows, coo.cols, off, idx, val)


def csr_to_csc(csr: csr_t, stream=None) -> csc_t:
    """Structural transpose; entries ordered by (column, row)."""
    _need_cuda(csr)
    dev = csr.values.device
    off = torch.empty(csr.cols + 1, dtype=torch.int32, device=dev)
    rid = torch.empty(csr.nnzs, dtype=torch.int32, device=dev)
    val = torch.empty(csr.nnzs, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().loopsb_csr_to_csc(csr.rows, csr.cols, csr.nnzs, _lib.ptr(csr.offsets), _lib.ptr(csr.indices),
                                             _lib.ptr(csr.values), _lib.ptr(off), _lib.ptr(rid), _lib.ptr(val),
                                             _lib.stream_ptr(stream)), "loopsb_csr_to_csc")
    return csc_t.from_tensors(csr.rows, csr.cols, off, rid, val)


def csr_max_degree(csr: csr_t, stream=None) -> int:
    _need_cuda(csr)
    out = C.c_int32(0)
    _lib.check(_lib.load().loopsb_csr_max_degree(csr.rows, _lib.ptr(csr.offsets), C.byref(out),
                                                 _lib.stream_ptr(stream)), "loopsb_csr_max_degree")
    return int(out.value)


def csr_to_ell(csr: csr_t, stream=None) -> ell_t:
    """pitch = widest row; padding column -1 / value 0."""
    _need_cuda(csr)
    dev = csr.values.device
    pitch = csr_max_degree(csr, stream)
    e_idx = torch.empty(csr.rows * pitch, dtype=torch.int32, device=dev)
    e_val = torch.empty(csr.rows * pitch, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().loopsb_csr_to_ell(csr.rows, pitch, _lib.ptr(csr.offsets), _lib.ptr(csr.indices),
                                             _lib.ptr(csr.values), _lib.ptr(e_idx), _lib.ptr(e_val),
                                             _lib.stream_ptr(stream)), "loopsb_csr_to_ell")
    return ell_t.from_tensors(csr.rows, csr.cols, csr.nnzs, pitch, e_idx, e_val)


def csr_to_bcsr(csr: csr_t, R: int, C_: int, value_dtype=torch.float32, stream=None) -> bcsr_t:
    """R x C dense blocks, block columns ascending per block-row, zero padding;
    values in fp32 or bf16 (rounded to nearest even from the fp32 input)."""
    _need_cuda(csr)
    if value_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("block values are fp32 or bf16")
    dev = csr.values.device
    lib = _lib.load()
    nbr = (csr.rows + R - 1) // R
    b_off = torch.empty(nbr + 1, dtype=torch.int32, device=dev)
    atom_block = torch.empty(csr.nnzs, dtype=torch.int32, device=dev)
    nb = C.c_int64(0)
    _lib.check(lib.loopsb_csr_to_bcsr_count(R, C_, csr.rows, csr.cols, csr.nnzs, _lib.ptr(csr.offsets),
                                            _lib.ptr(csr.indices), _lib.ptr(b_off), _lib.ptr(atom_block),
                                            C.byref(nb), _lib.stream_ptr(stream)), "loopsb_csr_to_bcsr_count")
    nb = int(nb.value)
    b_col = torch.empty(nb, dtype=torch.int32, device=dev)
    b_val = torch.empty(nb * R * C_, dtype=value_dtype, device=dev)
    _lib.check(lib.loopsb_csr_to_bcsr_fill(R, C_, csr.rows, csr.cols, csr.nnzs, _lib.ptr(csr.offsets),
                                           _lib.ptr(csr.indices), _lib.ptr(csr.values), _lib.ptr(atom_block), nb,
                                           _lib.ptr(b_col), _lib.ptr(b_val), int(value_dtype == torch.bfloat16),
                                           _lib.stream_ptr(stream)), "loopsb_csr_to_bcsr_fill")
    return bcsr_t.from_tensors(R, C_, csr.rows, csr.cols, csr.nnzs, b_off, b_col, b_val)


def csr_to_dia(csr: csr_t, stream=None) -> dia_t:
    """Distinct (col - row) offsets ascending, values[d * rows + r], zero padding."""
    _need_cuda(csr)
    dev = csr.values.device
    lib = _lib.load()
    nd = C.c_int32(0)
    _lib.check(lib.loopsb_csr_to_dia_count(csr.rows, csr.cols, csr.nnzs, _lib.ptr(csr.offsets), _lib.ptr(csr.indices),
                                           C.byref(nd), _lib.stream_ptr(stream)), "loopsb_csr_to_dia_count")
    nd = int(nd.value)
    offs = torch.empty(nd, dtype=torch.int32, device=dev)
    vals = torch.empty(nd * csr.rows, dtype=torch.float32, device=dev)
    _lib.check(lib.loopsb_csr_to_dia_fill(csr.rows, csr.cols, csr.nnzs, _lib.ptr(csr.offsets), _lib.ptr(csr.indices),
                                          _lib.ptr(csr.values), nd, _lib.ptr(offs), _lib.ptr(vals),
                                          _lib.stream_ptr(stream)), "loopsb_csr_to_dia_fill")
    return dia_t.from_tensors(csr.rows, csr.cols, csr.nnzs, offs, vals)


def csr_split_columns(csr: csr_t, chunk_cols: int, block_of_chunk, stream=None) -> list[csr_t]:
    """Column blocks of a CSR matrix (``loopsb_csr_split_columns_*``): columns are cut into
    equal chunks of ``chunk_cols`` and ``block_of_chunk[c]`` names the block chunk c goes
    to. Blocks keep CSR order and GLOBAL column ids and share one allocation, each
    starting on a 16-byte boundary (what the multi-GPU plan holds per shard)."""
    import numpy as np
    _need_cuda(csr)
    dev = csr.values.device
    lib = _lib.load()
    nchunks, nblocks = len(block_of_chunk), int(max(block_of_chunk)) + 1
    boc = np.ascontiguousarray(block_of_chunk, np.int32)
    boff = torch.empty(nblocks * (csr.rows + 1), dtype=torch.int32, device=dev)
    bnnz = np.zeros(nblocks, np.int64)
    _lib.check(lib.loopsb_csr_split_columns_count(csr.rows, csr.nnzs, _lib.ptr(csr.offsets), _lib.ptr(csr.indices),
                                                  int(chunk_cols), nchunks, boc.ctypes.data, nblocks, _lib.ptr(boff),
                                                  bnnz.ctypes.data, _lib.stream_ptr(stream)),
               "loopsb_csr_split_columns_count")
    idx = torch.empty(csr.nnzs + 4 * nblocks, dtype=torch.int32, device=dev)
    val = torch.empty(csr.nnzs + 4 * nblocks, dtype=torch.float32, device=dev)
    _lib.check(lib.loopsb_csr_split_columns_fill(csr.rows, csr.nnzs, _lib.ptr(csr.offsets), _lib.ptr(csr.indices),
                                                 _lib.ptr(csr.values), int(chunk_cols), nchunks, boc.ctypes.data,
                                                 nblocks, _lib.ptr(boff), bnnz.ctypes.data, _lib.ptr(idx),
                                                 _lib.ptr(val), _lib.stream_ptr(stream)),
               "loopsb_csr_split_columns_fill")
    out, base = [], 0
    for b in range(nblocks):
        base = (base + 3) & ~3
        n = int(bnnz[b])
        out.append(csr_t.from_tensors(csr.rows, csr.cols, boff[b * (csr.rows + 1): (b + 1) * (csr.rows + 1)],
                                      idx[base: base + n], val[base: base + n]))
        base += n
    return out
