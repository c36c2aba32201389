"""Owning sparse containers on the GPU, mirroring the reference's
``csr_t / coo_t / ell_t / bcsr_t`` (reference include/loops/container/
{csr,coo,ell,bcsr}.hxx): same array names, dtypes (int32 ids, fp32 values) and
padding rules. Storage is torch tensors (device memory + streams are torch's
job here); the compute never touches torch.

Format conversions run on the host with numpy, like the reference's
(``ell_t(csr)`` ell.hxx:113-145, ``bcsr_t(csr)`` bcsr.hxx:111-194,
``coo_t(csr)`` coo.hxx:87-98); they are set-up code, not the SpMV hot path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .layout import (csr as csr_layout, coo as coo_layout, ell as ell_layout, bcsr as bcsr_layout,
                     csc as csc_layout, dia as dia_layout, flat_uniform_occupancy)


def _dev(a, dtype, device):
    t = torch.as_tensor(np.ascontiguousarray(a), dtype=dtype)
    return t.to(device) if device is not None else t


class _planned:
    """Caches one plan (the reference's preprocess_t) per schedule."""

    def __init__(self):
        self._plans = {}

    def plan(self, schedule: int, stream=None):
        from .plan import Plan
        p = self._plans.get(schedule)
        if p is None:
            p = Plan(self.layout(), schedule, stream)
            self._plans[schedule] = p
        return p

    def drop_plans(self):
        for p in self._plans.values():
            p.close()
        self._plans.clear()


class csr_t(_planned):
    """rows, cols, nnzs; offsets[rows+1], indices[nnzs], values[nnzs]."""

    def __init__(self, rows, cols, offsets, indices, values, device="cuda"):
        super().__init__()
        self.rows, self.cols = int(rows), int(cols)
        self.offsets = _dev(offsets, torch.int32, device)
        self.indices = _dev(indices, torch.int32, device)
        self.values = _dev(values, torch.float32, device)
        self.nnzs = int(self.indices.numel())
        assert self.offsets.numel() == self.rows + 1

    @classmethod
    def from_tensors(cls, rows, cols, offsets, indices, values):
        self = cls.__new__(cls)
        _planned.__init__(self)
        self.rows, self.cols = int(rows), int(cols)
        self.offsets, self.indices, self.values = offsets, indices, values
        self.nnzs = int(indices.numel())
        return self

    def layout(self):
        return csr_layout(self.offsets, self.rows, self.nnzs)

    def plan(self, schedule: int, stream=None, tiled=None):
        """merge_path_flat plans may carry a band-tiled copy of the matrix
        (``Plan.tile_csr``). ``tiled``: None = the LOOPSB_TILED environment
        variable ("auto" by default: ask the library's cost model; "1" = force;
        "0" = never), True = force, False = plain CSR kernel."""
        import os
        p = super().plan(schedule, stream)
        if schedule != _lib.SCHED_MERGE_PATH_FLAT or not self.values.is_cuda or self.nnzs == 0:
            return p
        # Staleness guard: the tiled copy is keyed by array addresses, torch counts in-place
        # writes. If values / indices were written since the copy was made, drop it now (the
        # plain kernel reads the live arrays) and re-tile only once a call sees the same
        # version twice in a row -- a one-off update pays one re-tile, a matrix whose values
        # change every iteration never does.
        ver = (self.indices._version, self.values._version, self.indices.data_ptr(), self.values.data_ptr())
        if getattr(p, "_tile_state", None) in ("forced", "auto") and getattr(p, "_tile_ver", ver) != ver:
            p.invalidate()
            p._tile_state = "stale"
            p._tile_ver = ver
            p._stale_reason = "values or column ids were modified in place after tiling"
            if tiled is None or tiled == "auto":
                return p
        elif getattr(p, "_tile_state", None) == "stale":
            if getattr(p, "_tile_ver", None) != ver:
                p._tile_ver = ver                      # still changing: stay on the plain kernel
                if tiled is not True:
                    return p
            p._tile_state = None                       # settled (or forced): tile again below
        if tiled is None:
            env = os.environ.get("LOOPSB_TILED", "auto")
            tiled = {"0": False, "1": True}.get(env, "auto")
        state = getattr(p, "_tile_state", None)
        if tiled is False:
            if state not in (None, "off"):
                p.untile()
            p._tile_state = "off"
        elif state != ("forced" if tiled is True else "auto") and not (state == "forced" and tiled == "auto"):
            p.tile_csr(self.indices, self.values, self.cols, force=(tiled is True), stream=stream)
            p._tile_state = "forced" if tiled is True else "auto"
            p._tile_ver = ver
        return p

    def values_changed(self):
        """Tell every cached plan that values / column ids were changed in place (what the
        version check in ``plan`` does by itself for torch in-place ops; needed only when the
        arrays were written behind torch's back, e.g. by a raw CUDA kernel)."""
        for p in self._plans.values():
            p.invalidate()
            if getattr(p, "_tile_state", None) in ("forced", "auto"):
                p._tile_state = "stale"
                p._tile_ver = None

    def host(self):
        return (self.offsets.cpu().numpy(), self.indices.cpu().numpy(), self.values.cpu().numpy())

    def flat_plan(self, K: int, stream=None):
        """thread_mapped plan over layout::flat_uniform_occupancy<K, csr> (cached per K)."""
        from .plan import Plan
        key = ("flat", int(K))
        p = self._plans.get(key)
        if p is None:
            p = Plan(flat_uniform_occupancy(K, self.layout()), _lib.SCHED_THREAD_MAPPED, stream)
            self._plans[key] = p
        return p


class coo_t(_planned):
    """row_indices, col_indices, values, each [nnzs]."""

    def __init__(self, rows, cols, row_indices, col_indices, values, device="cuda"):
        super().__init__()
        self.rows, self.cols = int(rows), int(cols)
        self.row_indices = _dev(row_indices, torch.int32, device)
        self.col_indices = _dev(col_indices, torch.int32, device)
        self.values = _dev(values, torch.float32, device)
        self.nnzs = int(self.values.numel())

    @classmethod
    def from_csr(cls, csr: csr_t):
        off, idx, val = csr.host()
        rows_of = np.repeat(np.arange(csr.rows, dtype=np.int32), np.diff(off))
        return cls(csr.rows, csr.cols, rows_of, idx, val, device=csr.values.device)

    def layout(self):
        return coo_layout(self.nnzs)


class ell_t(_planned):
    """Row-major rows*pitch slabs; padding = column -1, value 0."""

    SENTINEL = -1

    def __init__(self, rows, cols, nnzs, pitch, indices, values, device="cuda"):
        super().__init__()
        self.rows, self.cols, self.nnzs, self.pitch = int(rows), int(cols), int(nnzs), int(pitch)
        self.indices = _dev(indices, torch.int32, device)
        self.values = _dev(values, torch.float32, device)

    @classmethod
    def from_csr(cls, csr: csr_t):
        off, idx, val = csr.host()
        deg = np.diff(off)
        pitch = int(deg.max()) if csr.rows else 0
        e_idx = np.full((csr.rows, pitch), cls.SENTINEL, dtype=np.int32)
        e_val = np.zeros((csr.rows, pitch), dtype=np.float32)
        if csr.nnzs:
            r = np.repeat(np.arange(csr.rows), deg)
            slot = np.arange(csr.nnzs) - np.repeat(off[:-1], deg)
            e_idx[r, slot] = idx
            e_val[r, slot] = val
        return cls(csr.rows, csr.cols, csr.nnzs, pitch, e_idx.reshape(-1), e_val.reshape(-1),
                   device=csr.values.device)

    def layout(self):
        return ell_layout(self.rows, self.pitch)


class csc_t(_planned):
    """offsets[cols+1], indices[nnzs] = ROW ids, values[nnzs]; entries ordered by
    (column, row) (reference container/csc.hxx:88-102)."""

    def __init__(self, rows, cols, offsets, indices, values, device="cuda"):
        super().__init__()
        self.rows, self.cols = int(rows), int(cols)
        self.offsets = _dev(offsets, torch.int32, device)
        self.indices = _dev(indices, torch.int32, device)
        self.values = _dev(values, torch.float32, device)
        self.nnzs = int(self.indices.numel())

    @classmethod
    def from_csr(cls, csr: csr_t):
        off, idx, val = csr.host()
        r = np.repeat(np.arange(csr.rows, dtype=np.int32), np.diff(off))
        order = np.argsort(idx, kind="stable")          # (col, row): CSR order is row-major already
        c_off = np.zeros(csr.cols + 1, dtype=np.int64)
        np.add.at(c_off, idx.astype(np.int64) + 1, 1)
        return cls(csr.rows, csr.cols, np.cumsum(c_off).astype(np.int32), r[order], val[order],
                   device=csr.values.device)

    def layout(self):
        return csc_layout(self.offsets, self.cols, self.nnzs)


class dia_t:
    """Distinct (col - row) offsets ascending; values column-major
    values[d * stride + r], stride = rows, entries assigned (reference
    container/dia.hxx:56-62,135-188)."""

    def __init__(self, rows, cols, nnzs, diag_offsets, values, device="cuda"):
        self.rows, self.cols, self.nnzs = int(rows), int(cols), int(nnzs)
        self.stride = self.rows
        self.diag_offsets = _dev(diag_offsets, torch.int32, device)
        self.values = _dev(values, torch.float32, device)
        self.num_diagonals = int(self.diag_offsets.numel())

    @classmethod
    def from_csr(cls, csr: csr_t):
        off, idx, val = csr.host()
        r = np.repeat(np.arange(csr.rows, dtype=np.int64), np.diff(off))
        o = idx.astype(np.int64) - r
        offs = np.unique(o)
        d = np.searchsorted(offs, o)
        d_val = np.zeros(len(offs) * csr.rows, dtype=np.float32)
        d_val[d * csr.rows + r] = val                    # assignment, last wins
        return cls(csr.rows, csr.cols, csr.nnzs, offs.astype(np.int32), d_val, device=csr.values.device)

    def layout(self):
        return dia_layout(self.rows, self.num_diagonals)


class bcsr_t(_planned):
    """R x C dense blocks: block_offsets[nbr+1], block_col_indices[nb],
    values[nb*R*C] with values[b*R*C + i*C + j]; block columns ascending per
    block-row; entries assigned (not accumulated); padding zero."""

    def __init__(self, R, C, rows, cols, nnzs, block_offsets, block_col_indices, values,
                 device="cuda", value_dtype=torch.float32):
        super().__init__()
        self.R, self.C = int(R), int(C)
        self.rows, self.cols, self.nnzs = int(rows), int(cols), int(nnzs)
        self.num_block_rows = (self.rows + self.R - 1) // self.R
        self.num_block_cols = (self.cols + self.C - 1) // self.C
        self.block_offsets = _dev(block_offsets, torch.int32, device)
        self.block_col_indices = _dev(block_col_indices, torch.int32, device)
        self.values = _dev(values, torch.float32, device).to(value_dtype)
        self.num_blocks = int(self.block_col_indices.numel())

    @classmethod
    def from_csr(cls, csr: csr_t, R: int, C: int, value_dtype=torch.float32):
        off, idx, val = csr.host()
        rows, cols = csr.rows, csr.cols
        nbr, nbc = (rows + R - 1) // R, (cols + C - 1) // C
        r = np.repeat(np.arange(rows, dtype=np.int64), np.diff(off))
        br, bc = r // R, idx.astype(np.int64) // C
        key = br * max(nbc, 1) + bc
        uniq, inv = np.unique(key, return_inverse=True)   # sorted: (br, bc) ascending
        nb = int(uniq.size)
        b_off = np.zeros(nbr + 1, dtype=np.int64)
        np.add.at(b_off, (uniq // max(nbc, 1)) + 1, 1)
        b_off = np.cumsum(b_off)
        b_col = (uniq % max(nbc, 1)).astype(np.int32)
        b_val = np.zeros(nb * R * C, dtype=np.float32)
        b_val[inv * (R * C) + (r % R) * C + (idx % C)] = val   # assignment, last wins
        return cls(R, C, rows, cols, csr.nnzs, b_off.astype(np.int32), b_col, b_val,
                   device=csr.values.device, value_dtype=value_dtype)

    def layout(self):
        return bcsr_layout(self.block_offsets, self.num_block_rows, self.num_blocks)

    def padded_x(self, x: torch.Tensor) -> torch.Tensor:
        """x padded with zeros to num_block_cols*C (examples/spmv/bcsr_thread_mapped.cu:41)."""
        n = self.num_block_cols * self.C
        if x.numel() == n:
            return x
        out = torch.zeros(n, dtype=x.dtype, device=x.device)
        out[: x.numel()] = x
        return out


def _attach_from_tensors():
    """Zero-copy constructors for data that already lives on the device."""

    def coo_from_tensors(cls, rows, cols, row_indices, col_indices, values):
        self = cls.__new__(cls)
        _planned.__init__(self)
        self.rows, self.cols = int(rows), int(cols)
        self.row_indices, self.col_indices, self.values = row_indices, col_indices, values
        self.nnzs = int(values.numel())
        return self

    def ell_from_tensors(cls, rows, cols, nnzs, pitch, indices, values):
        self = cls.__new__(cls)
        _planned.__init__(self)
        self.rows, self.cols, self.nnzs, self.pitch = int(rows), int(cols), int(nnzs), int(pitch)
        self.indices, self.values = indices, values
        return self

    def bcsr_from_tensors(cls, R, C, rows, cols, nnzs, block_offsets, block_col_indices, values):
        self = cls.__new__(cls)
        _planned.__init__(self)
        self.R, self.C = int(R), int(C)
        self.rows, self.cols, self.nnzs = int(rows), int(cols), int(nnzs)
        self.num_block_rows = (self.rows + self.R - 1) // self.R
        self.num_block_cols = (self.cols + self.C - 1) // self.C
        self.block_offsets, self.block_col_indices, self.values = block_offsets, block_col_indices, values
        self.num_blocks = int(block_col_indices.numel())
        return self

    def csc_from_tensors(cls, rows, cols, offsets, indices, values):
        self = cls.__new__(cls)
        _planned.__init__(self)
        self.rows, self.cols = int(rows), int(cols)
        self.offsets, self.indices, self.values = offsets, indices, values
        self.nnzs = int(indices.numel())
        return self

    def dia_from_tensors(cls, rows, cols, nnzs, diag_offsets, values):
        self = cls.__new__(cls)
        self.rows, self.cols, self.nnzs = int(rows), int(cols), int(nnzs)
        self.stride = self.rows
        self.diag_offsets, self.values = diag_offsets, values
        self.num_diagonals = int(diag_offsets.numel())
        return self

    csc_t.from_tensors = classmethod(csc_from_tensors)
    dia_t.from_tensors = classmethod(dia_from_tensors)
    coo_t.from_tensors = classmethod(coo_from_tensors)
    ell_t.from_tensors = classmethod(ell_from_tensors)
    bcsr_t.from_tensors = classmethod(bcsr_from_tensors)


_attach_from_tensors()


def csr_to_coo_device(csr: "csr_t") -> "coo_t":
    """CSR -> COO on the device (loops_b200.convert.csr_to_coo)."""
    from .convert import csr_to_coo
    return csr_to_coo(csr)


def csr_to_ell_device(csr: "csr_t") -> "ell_t":
    """CSR -> ELL on the device (loops_b200.convert.csr_to_ell)."""
    from .convert import csr_to_ell
    return csr_to_ell(csr)
