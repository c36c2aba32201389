"""Plan = the reference's ``schedule::merge_path::preprocess_t`` generalised
(reference schedule/merge_path_flat.hxx:92-172): per-matrix coordinates,
workspace and launch geometry behind ``loopsb_plan_*``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class Plan:
    def __init__(self, layout, schedule: int, stream=None):
        lib = _lib.load()
        self._lib = lib
        self.layout = layout          # keeps the offsets tensor alive
        self.schedule = schedule
        self._desc = layout.desc()
        h = C.c_void_p()
        _lib.check(lib.loopsb_plan_create(C.byref(h), C.byref(self._desc), schedule,
                                          _lib.stream_ptr(stream)), "loopsb_plan_create")
        self.handle = h

    def info(self) -> _lib.PlanInfo:
        i = _lib.PlanInfo()
        _lib.check(self._lib.loopsb_plan_info(self.handle, C.byref(i)), "loopsb_plan_info")
        return i

    def merge_coords(self) -> np.ndarray:
        """[(M+1), 2] int32: S(b * TPB*IPT) for b = 0..M."""
        m = self.info().num_merge_tiles
        out = np.zeros((m + 1, 2), dtype=np.int32)
        _lib.check(self._lib.loopsb_plan_merge_coords_host(self.handle, out.ctypes.data, m + 1),
                   "loopsb_plan_merge_coords_host")
        return out

    def tile_csr(self, indices, values, cols: int, force: bool = False, stream=None) -> bool:
        """Give the plan a band-tiled copy of the CSR matrix (``loopsb_plan_tile_csr``).
        Returns False when the library declines (matrix does not fit the format
        or the cost model prefers the plain kernel); the plan stays usable."""
        rc = self._lib.loopsb_plan_tile_csr(self.handle, _lib.ptr(indices), _lib.ptr(values), int(cols),
                                            _lib.TILE_FORCE if force else 0, _lib.stream_ptr(stream))
        if rc == _lib.ERR_UNSUPPORTED:
            self.tile_declined = self._lib.loopsb_last_error().decode(errors="replace")
            return False
        _lib.check(rc, "loopsb_plan_tile_csr")
        self._tiled_keep = (indices, values)   # the copy is keyed by these pointers
        return True

    def invalidate(self):
        """The matrix values / column ids were changed in place: drop every plan-owned copy of
        them (``loopsb_plan_invalidate``); SpMV calls read the live arrays from now on."""
        _lib.check(self._lib.loopsb_plan_invalidate(self.handle), "loopsb_plan_invalidate")
        self._tiled_keep = None
        self._packed_key = None

    def tile_breakeven(self, cols: int) -> int:
        """SpMV calls after which tiling has paid for itself (-1 = never)."""
        n = C.c_int64(-1)
        _lib.check(self._lib.loopsb_plan_tile_breakeven(self.handle, int(cols), C.byref(n)), "loopsb_plan_tile_breakeven")
        return int(n.value)

    def untile(self):
        _lib.check(self._lib.loopsb_plan_untile(self.handle), "loopsb_plan_untile")
        self._tiled_keep = None

    def tiled_info(self):
        """dict of loopsb_tiled_info_t, or None when the plan holds no tiled copy."""
        i = _lib.TiledInfo()
        rc = self._lib.loopsb_plan_tiled_info(self.handle, C.byref(i))
        if rc == _lib.ERR_UNSUPPORTED:
            return None
        _lib.check(rc, "loopsb_plan_tiled_info")
        return i.as_dict()

    def probe_begin(self, capacity: int):
        """Bracket the dominant kernel of the next `capacity` SpMV calls with
        CUDA events (recorded on each call's stream)."""
        _lib.check(self._lib.loopsb_plan_probe_begin(self.handle, capacity), "loopsb_plan_probe_begin")

    def probe_collect(self, capacity: int) -> np.ndarray:
        ms = np.zeros(capacity, dtype=np.float32)
        n = C.c_int32()
        _lib.check(self._lib.loopsb_plan_probe_collect(self.handle, ms.ctypes.data, capacity, C.byref(n)),
                   "loopsb_plan_probe_collect")
        return ms[: n.value]

    def close(self):
        if getattr(self, "handle", None):
            self._lib.loopsb_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
