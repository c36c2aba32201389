"""CPU: the C-ABI library loads and exports every symbol include/loopsb.h
declares; without a GPU the compute entry points fail loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "loopsb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(loopsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from loops_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/loopsb.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert lib.loopsb_version() == 100
    assert lib.loopsb_status_string(3).decode().startswith("unsupported")


def test_argument_validation_and_no_cpu_fallback():
    import torch
    from loops_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.loopsb_plan_create(C.byref(h), None, 0, None) == _lib.ERR_INVALID
    d = _lib.LayoutDesc()
    d.kind, d.num_tiles, d.num_atoms = _lib.LAYOUT_CSR, 4, 7
    d.offsets = None
    assert lib.loopsb_plan_create(C.byref(h), C.byref(d), 0, None) == _lib.ERR_INVALID
    assert b"offsets" in lib.loopsb_last_error()
    d.num_tiles = -1
    assert lib.loopsb_plan_create(C.byref(h), C.byref(d), 0, None) == _lib.ERR_INVALID
    if not torch.cuda.is_available():
        off = np.array([0, 2, 2, 5, 7], np.int32)
        d.num_tiles, d.offsets = 4, off.ctypes.data
        assert lib.loopsb_plan_create(C.byref(h), C.byref(d), 0, None) == _lib.ERR_CUDA
        assert b"no CPU fallback" in lib.loopsb_last_error()
        y = np.zeros(4, np.float32)
        rc = lib.loopsb_spmv_csr_host_f32(0, 4, 4, 7, off.ctypes.data, off.ctypes.data, y.ctypes.data,
                                          y.ctypes.data, y.ctypes.data, None)
        assert rc == _lib.ERR_CUDA
        sm = C.c_int32()
        assert lib.loopsb_device_info(C.byref(sm), None, None) == _lib.ERR_CUDA


def test_missing_library_is_loud(monkeypatch, tmp_path):
    from loops_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU / PyTorch fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    """The product path must not reach into oracle/ (only tests, smoke and the
    bench baseline may)."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "loops_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hxx")):
                if "oracle" in open(os.path.join(base, f), errors="ignore").read():
                    bad.append(f)
    for base, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            if "oracle" in open(os.path.join(base, f), errors="ignore").read():
                bad.append(f)
    assert not bad, bad
