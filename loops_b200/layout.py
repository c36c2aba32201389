"""Layout views (reference include/loops/container/layout.hxx): the six-method
tile/atom contract, as small Python objects that can describe themselves to the
C ABI (``loopsb_layout_t``). They are non-owning: a view over an offsets tensor
keeps a reference to it only so the device pointer stays valid."""
from __future__ import annotations

from . import _lib


class _view:
    kind = None
    offsets = None
    pitch = 0

    def desc(self) -> _lib.LayoutDesc:
        d = _lib.LayoutDesc()
        d.kind = self.kind
        d.num_tiles = self.num_tiles()
        d.num_atoms = self.num_atoms()
        d.pitch = self.pitch
        d.offsets = _lib.ptr(self.offsets)
        return d


class _offsets_view(_view):
    def __init__(self, offsets, num_tiles, num_atoms):
        self.offsets, self._t, self._a = offsets, int(num_tiles), int(num_atoms)

    def num_tiles(self): return self._t
    def num_atoms(self): return self._a
    def tile_begin(self, t): return int(self.offsets[t])
    def tile_end(self, t): return int(self.offsets[t + 1])
    def tile_size(self, t): return self.tile_end(t) - self.tile_begin(t)


class csr(_offsets_view):
    kind = _lib.LAYOUT_CSR


class csc(_offsets_view):
    kind = _lib.LAYOUT_CSC


class bcsr(_offsets_view):
    kind = _lib.LAYOUT_BCSR


class coo(_view):
    kind = _lib.LAYOUT_COO
    pitch = 1

    def __init__(self, nnz): self._n = int(nnz)
    def num_tiles(self): return self._n
    def num_atoms(self): return self._n
    def tile_begin(self, t): return t
    def tile_end(self, t): return t + 1
    def tile_size(self, t): return 1


class _pitch_view(_view):
    def __init__(self, num_tiles, pitch):
        self._t, self.pitch = int(num_tiles), int(pitch)

    def num_tiles(self): return self._t
    def num_atoms(self): return self._t * self.pitch
    def tile_begin(self, t): return t * self.pitch
    def tile_end(self, t): return (t + 1) * self.pitch
    def tile_size(self, t): return self.pitch


class ell(_pitch_view):
    kind = _lib.LAYOUT_ELL


class dia(_pitch_view):
    kind = _lib.LAYOUT_DIA


class flat_uniform_occupancy(_view):
    """Windows of K consecutive atoms over a base offsets layout (reference
    container/partitioning.hxx:71-141): tile_end(t) = min((t+1)K, A); the base
    layout (``base()``) keeps answering tile_of."""
    kind = _lib.LAYOUT_FLAT

    def __init__(self, K: int, base: _offsets_view):
        assert K > 0
        self.pitch, self._base = int(K), base
        self.offsets = base.offsets

    def base(self): return self._base
    def num_atoms(self): return self._base.num_atoms()
    def num_tiles(self): return (self.num_atoms() + self.pitch - 1) // self.pitch
    def tile_begin(self, t): return t * self.pitch
    def tile_end(self, t): return min((t + 1) * self.pitch, self.num_atoms())
    def tile_size(self, t): return self.tile_end(t) - self.tile_begin(t)
