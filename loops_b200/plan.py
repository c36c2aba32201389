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
