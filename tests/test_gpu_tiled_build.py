"""GPU: the device-side builder of the band-tiled copy (loops_b200/csrc/
tiled_build.cuh) must produce the image of the host builder (bt::build_host, the
one the CPU format tests walk with the kernel emulation) BYTE FOR BYTE: steps,
stream bases, row-block cuts and the statistics."""
import os

import numpy as np
import pytest

from helpers import load_chesapeake, random_csr
from tiled_emul import build_image

pytestmark = pytest.mark.gpu


def _device_image(off, idx, val, rows, cols, geometry):
    from loops_b200 import _lib, csr_t
    lib = _lib.load()
    old = os.environ.get("LOOPSB_TILED_GEOM")
    if geometry is not None:
        os.environ["LOOPSB_TILED_GEOM"] = ",".join(str(v) for v in geometry)
    try:
        A = csr_t(rows, cols, off, idx, val)
        plan = A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled=True)
        info = plan.tiled_info()
        assert info is not None
    finally:
        if geometry is not None:
            if old is None:
                del os.environ["LOOPSB_TILED_GEOM"]
            else:
                os.environ["LOOPSB_TILED_GEOM"] = old
    ns = info["nb"] * info["q"] * info["warps"]
    words = (info["total_steps"] + info["es"]) * 256
    steps = np.zeros(words, np.uint32)
    base = np.zeros(ns + 1, np.int32)
    blk = np.zeros(info["nb"] + 1, np.int32)
    _lib.check(lib.loopsb_plan_tiled_download(plan.handle, steps.ctypes.data, words, base.ctypes.data, blk.ctypes.data),
               "loopsb_plan_tiled_download")
    return info, steps.reshape(-1, 256), base, blk, A, plan


def _compare(off, idx, val, rows, cols, geometry):
    from loops_b200 import _lib
    lib = _lib.load()
    info, steps, base, blk, A, plan = _device_image(off, idx, val, rows, cols, geometry)
    geo = tuple(info[k] for k in ("nb", "q", "warps", "cb", "xb", "es"))
    rc, img = build_image(lib, rows, cols, off, idx, val, geo)
    assert rc == 0
    h = img["g"]
    for k in ("rb", "cq", "nband", "total_steps", "real_entries", "pad_entries", "flagged_entries", "flagged_steps",
              "long_steps", "smem_bytes"):
        assert info[k] == h[k], (k, info[k], h[k])
    np.testing.assert_array_equal(blk, img["blk_begin"])
    np.testing.assert_array_equal(base, img["stream_base"])
    np.testing.assert_array_equal(steps, img["steps"])
    return info


GEOMS = [None, (3, 2, 4, 8, 2, 2), (5, 1, 8, 16, 3, 3), (2, 4, 16, 64, 2, 3), (7, 3, 12, 4, 2, 2)]


@pytest.mark.parametrize("geometry", GEOMS)
def test_chesapeake_image(geometry):
    c = load_chesapeake()
    _compare(c["off"], c["idx"], c["val"], 39, 39, geometry)


@pytest.mark.parametrize("geometry", GEOMS)
@pytest.mark.parametrize("shape", [(200, 150, 0.05), (64, 1000, 0.01), (500, 37, 0.3), (3000, 3000, 0.004)])
def test_random_images(geometry, shape):
    rows, cols, dens = shape
    off, idx, val = random_csr(rows, cols, dens, seed=rows + cols, empty_every=7)
    _compare(off, idx, val, rows, cols, geometry)


def test_dense_rows_and_empty_bands():
    off, idx, val = random_csr(64, 2000, 0.01, seed=5, heavy_row=(3, 2000))
    for geometry in [(2, 1, 4, 512, 2, 2), (2, 2, 8, 128, 2, 3), None]:
        info = _compare(off, idx, val, 64, 2000, geometry)
        assert info["long_steps"] > 0
    rows, cols = 96, 512
    rng = np.random.default_rng(11)
    o, i, v = [0], [], []
    for r in range(rows):
        c = np.unique(np.concatenate([rng.integers(0, 6, 2), rng.integers(500, 512, 2), rng.integers(250, 262, 1)])).astype(np.int32)
        i.append(c); v.append(rng.uniform(0.5, 1.5, len(c)).astype(np.float32)); o.append(o[-1] + len(c))
    o, i, v = np.array(o, np.int32), np.concatenate(i), np.concatenate(v)
    for geometry in [(2, 1, 4, 8, 2, 2), (2, 2, 4, 8, 3, 2), (1, 1, 4, 4, 2, 2)]:
        _compare(o, i, v, rows, cols, geometry)


def test_powerlaw_2_16_rows_default_geometry():
    from loops_b200 import generate as g
    rows = cols = 1 << 16
    off, idx, val = g.synth_csr(rows, cols, rows * 32)
    _compare(off.numpy(), idx.numpy(), val.numpy(), rows, cols, None)


def test_out_of_range_column_is_rejected_on_the_device():
    from loops_b200 import _lib, csr_t
    off = np.array([0, 2, 3], np.int32)
    idx = np.array([0, 9, 1], np.int32)
    val = np.ones(3, np.float32)
    A = csr_t(2, 4, off, idx, val)
    with pytest.raises(_lib.LoopsbError) as e:
        A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled=True)
    assert e.value.status == _lib.ERR_INVALID


def test_host_builder_switch_gives_the_same_image(monkeypatch):
    off, idx, val = random_csr(700, 900, 0.01, seed=2)
    a = _device_image(off, idx, val, 700, 900, (3, 2, 8, 64, 2, 2))
    monkeypatch.setenv("LOOPSB_TILED_HOST_BUILD", "1")
    b = _device_image(off, idx, val, 700, 900, (3, 2, 8, 64, 2, 2))
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[2], b[2])
    np.testing.assert_array_equal(a[3], b[3])


def test_unpacked_dirty_images_match(monkeypatch):
    monkeypatch.setenv("LOOPSB_TILED_PACK", "0")
    off, idx, val = random_csr(80, 640, 0.3, seed=9)
    info = _compare(off, idx, val, 80, 640, (1, 1, 4, 8, 2, 2))
    assert info["flagged_steps"] > 0
