"""GPU parity of the band-tiled merge_path_flat kernel (spmv_tiled.cuh) through
the public entry point ``spmv.merge_path_flat(csr, x, y, tiled=True)`` -> C ABI:
bit-exact against the oracle on exactly representable inputs, <= 1e-6 relative
otherwise, over geometries that exercise band switches, the x ring, ragged
tails, the flagged (row-collision) path and the q-way partial reduction."""
import os

import numpy as np
import pytest
import torch

from helpers import load_chesapeake, random_csr

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6


def _run(off, idx, val, x, rows, cols, geometry=None, repeat=1, want_info=False):
    from loops_b200 import _lib, csr_t
    from loops_b200.algorithms import spmv
    old = os.environ.get("LOOPSB_TILED_GEOM")
    if geometry is not None:
        os.environ["LOOPSB_TILED_GEOM"] = ",".join(str(v) for v in geometry)
    try:
        A = csr_t(rows, cols, off, idx, val)
        xd = torch.as_tensor(x).cuda()
        ys = []
        for _ in range(repeat):
            y = torch.full((rows,), float("nan"), dtype=torch.float32, device="cuda")
            spmv.merge_path_flat(A, xd, y, tiled=True)
            ys.append(y.cpu().numpy())
        info = A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled=True).tiled_info()
        assert info is not None, "the tiled plan was not built"
    finally:
        if geometry is not None:
            if old is None:
                del os.environ["LOOPSB_TILED_GEOM"]
            else:
                os.environ["LOOPSB_TILED_GEOM"] = old
    for y in ys[1:]:
        np.testing.assert_array_equal(y, ys[0])
    return (ys[0], info) if want_info else ys[0]


GEOMS = [None, (3, 2, 4, 8, 2, 2), (5, 1, 8, 16, 3, 3), (2, 4, 16, 64, 2, 3), (7, 3, 12, 4, 2, 2)]


@pytest.mark.parametrize("geometry", GEOMS)
def test_chesapeake(geometry):
    c = load_chesapeake()
    y = _run(c["off"], c["idx"], c["val"], c["x"], 39, 39, geometry, repeat=3)
    np.testing.assert_array_equal(y, c["y"])
    assert float(y.sum()) == 1794.0


@pytest.mark.parametrize("geometry", GEOMS)
def test_battery(oracle, battery, geometry):
    from test_gpu_spmv import _assert_close
    for b in battery:
        if b["off"][-1] == 0:
            continue
        y = _run(b["off"], b["idx"], b["val"], b["x"], b["rows"], b["cols"], geometry)
        _assert_close(oracle, b["off"], b["idx"], b["val"], b["x"], y, ("tiled", b["name"], geometry))


@pytest.mark.parametrize("geometry", GEOMS[1:])
@pytest.mark.parametrize("shape", [(200, 150, 0.05), (64, 1000, 0.01), (500, 37, 0.3), (3000, 3000, 0.004)])
def test_random_exact(oracle, geometry, shape):
    rows, cols, dens = shape
    off, idx, val = random_csr(rows, cols, dens, seed=rows + cols, empty_every=7, exact=True)
    x = oracle.x_recipe_int(cols)
    y = _run(off, idx, val, x, rows, cols, geometry, repeat=2)
    np.testing.assert_array_equal(y, oracle.spmv(off, idx, val, x))


def test_dense_rows_take_the_long_run_path(oracle):
    off, idx, val = random_csr(64, 2000, 0.01, seed=5, heavy_row=(3, 2000), exact=True)
    x = oracle.x_recipe_int(2000)
    for geometry in [(2, 1, 4, 512, 2, 2), (2, 2, 8, 128, 2, 3), None]:
        y, info = _run(off, idx, val, x, 64, 2000, geometry, repeat=2, want_info=True)
        assert info["long_steps"] > 0 and info["flagged_entries"] == 0
        np.testing.assert_array_equal(y, oracle.spmv(off, idx, val, x))


def test_general_floats_within_tolerance(oracle):
    from test_gpu_spmv import _assert_close
    off, idx, val = random_csr(4000, 5000, 0.01, seed=77, empty_every=13, heavy_row=(17, 3000))
    x = oracle.x_recipe_float(5000, -1.0, 1.0, 9)
    for geometry in [None, (4, 4, 16, 256, 2, 3), (3, 2, 8, 64, 3, 2)]:
        y = _run(off, idx, val, x, 4000, 5000, geometry)
        _assert_close(oracle, off, idx, val, x, y, ("tiled-float", geometry))


def test_nan_in_x_stays_in_its_rows(oracle):
    """Padding entries must not leak a non-finite x into rows that do not
    reference that column."""
    off, idx, val = random_csr(300, 400, 0.03, seed=3, exact=True)
    x = oracle.x_recipe_int(400)
    x[17] = np.nan
    x[399] = np.inf
    ref = oracle.spmv(off, idx, val, x)
    for geometry in [None, (3, 2, 4, 8, 2, 2)]:
        y = _run(off, idx, val, x, 300, 400, geometry)
        np.testing.assert_array_equal(np.isnan(y), np.isnan(ref))
        np.testing.assert_array_equal(y[np.isfinite(ref)], ref[np.isfinite(ref)])


def test_other_pointers_fall_back_to_the_csr_kernel(oracle):
    """The tiled copy is keyed by the (indices, values) pointers it was made
    from; a call with different arrays must not use it."""
    from loops_b200 import _lib, csr_t
    lib = _lib.load()
    off, idx, val = random_csr(500, 500, 0.02, seed=8, exact=True)
    x = oracle.x_recipe_int(500)
    A = csr_t(500, 500, off, idx, val)
    plan = A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled=True)
    assert plan.tiled_info() is not None
    val2 = torch.as_tensor(val * 2).cuda()
    xd = torch.as_tensor(x).cuda()
    y = torch.empty(500, dtype=torch.float32, device="cuda")
    _lib.check(lib.loopsb_spmv_f32(plan.handle, _lib.ptr(val2), _lib.ptr(A.indices), None, _lib.ptr(xd),
                                   _lib.ptr(y), 500, 500, _lib.stream_ptr()), "loopsb_spmv_f32")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv(off, idx, (val * 2).astype(np.float32), x))
    # misaligned x: same fallback
    xo = torch.zeros(504, dtype=torch.float32, device="cuda")
    xo[1:501] = xd
    _lib.check(lib.loopsb_spmv_f32(plan.handle, _lib.ptr(A.values), _lib.ptr(A.indices), None,
                                   xo.data_ptr() + 4, _lib.ptr(y), 500, 500, _lib.stream_ptr()), "loopsb_spmv_f32")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv(off, idx, val, x))


def test_cost_model_declines_small_matrices():
    from loops_b200 import _lib, csr_t
    off, idx, val = random_csr(100, 100, 0.1, seed=1, exact=True)
    A = csr_t(100, 100, off, idx, val)
    plan = A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled="auto")
    assert plan.tiled_info() is None and "not profitable" in plan.tile_declined


def test_powerlaw_2_16_rows_default_geometry(oracle):
    from loops_b200 import generate as g
    rows = cols = 1 << 16
    off, idx, val = g.synth_csr(rows, cols, rows * 32)
    off, idx, val = off.numpy(), idx.numpy(), val.numpy()
    x = g.x_recipe(cols).numpy()
    y, info = _run(off, idx, val, x, rows, cols, None, repeat=3, want_info=True)
    np.testing.assert_array_equal(y, oracle.spmv(off, idx, val, x))
    assert info["real_entries"] == rows * 32


def test_full_size_config2_tiled_equals_plain():
    """BASELINE config 2 at full size: the band-tiled kernel (what bench.py times)
    and the plain CSR merge-path kernel give the same bits on the exact workload,
    launch after launch; the cost model accepts the matrix without being forced."""
    from loops_b200 import _lib, csr_t, generate as g
    from loops_b200.algorithms import spmv
    rows = cols = 1 << 20
    off, idx, val = g.synth_csr(rows, cols, 1 << 25, device="cuda")
    x = g.x_recipe(cols, device="cuda")
    A = csr_t.from_tensors(rows, cols, off, idx, val)
    y_plain = torch.full((rows,), float("nan"), device="cuda")
    spmv.merge_path_flat(A, x, y_plain, tiled=False)
    assert A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled=False).tiled_info() is None
    y = torch.full((rows,), float("nan"), device="cuda")
    for _ in range(3):
        y.fill_(float("nan"))
        spmv.merge_path_flat(A, x, y, tiled="auto")
        assert torch.equal(y, y_plain)
    info = A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled="auto").tiled_info()
    assert info is not None and info["flagged_steps"] * 20 <= info["total_steps"]


def test_dirty_steps_take_the_general_path(oracle, monkeypatch):
    """Packer off: rows seen in two separate cell ranges of a step set the dirty bit
    and go through the match.any path; results stay exact."""
    monkeypatch.setenv("LOOPSB_TILED_PACK", "0")
    off, idx, val = random_csr(80, 640, 0.3, seed=9, exact=True)
    x = oracle.x_recipe_int(640)
    for geometry in [(1, 1, 4, 8, 2, 2), (2, 2, 8, 16, 3, 2)]:
        y, info = _run(off, idx, val, x, 80, 640, geometry, repeat=2, want_info=True)
        assert info["flagged_steps"] > 0
        np.testing.assert_array_equal(y, oracle.spmv(off, idx, val, x))


def test_in_place_value_updates_never_return_stale_results(oracle):
    """The band-tiled copy holds VALUES and is keyed by array addresses. An in-place update
    must not be answered from the old copy: the container watches torch's version counters
    (loopsb_plan_invalidate), and ``values_changed()`` covers writes torch cannot see."""
    from loops_b200 import _lib, csr_t
    from loops_b200.algorithms import spmv
    off, idx, val = random_csr(3000, 2500, 0.01, seed=21, exact=True, empty_every=7, heavy_row=(5, 2000))
    x = oracle.x_recipe_int(2500)
    A = csr_t(3000, 2500, off, idx, val)
    xd = torch.as_tensor(x).cuda()
    y = torch.full((3000,), float("nan"), device="cuda")
    spmv.merge_path_flat(A, xd, y, tiled=True)
    y1 = oracle.spmv(off, idx, val, x)
    np.testing.assert_array_equal(y.cpu().numpy(), y1)
    plan = A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled=True)
    assert plan.tiled_info() is not None
    # 1. torch-visible in-place write, default (auto) call: the copy is dropped, the plain kernel answers
    A.values.mul_(2.0)
    y.fill_(float("nan"))
    spmv.merge_path_flat(A, xd, y)
    np.testing.assert_array_equal(y.cpu().numpy(), 2.0 * y1)
    assert A.plan(_lib.SCHED_MERGE_PATH_FLAT).tiled_info() is None
    # 2. forced again: re-tiled from the live values
    y.fill_(float("nan"))
    spmv.merge_path_flat(A, xd, y, tiled=True)
    np.testing.assert_array_equal(y.cpu().numpy(), 2.0 * y1)
    assert A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled=True).tiled_info() is not None
    # 3. forced call right after another in-place write: re-tiled, not stale
    A.values.mul_(0.25)
    y.fill_(float("nan"))
    spmv.merge_path_flat(A, xd, y, tiled=True)
    np.testing.assert_array_equal(y.cpu().numpy(), 0.5 * y1)
    # 4. a write behind torch's back (``.data`` has its own version counter) + values_changed()
    A.values.data.mul_(4.0)
    A.values_changed()
    y.fill_(float("nan"))
    spmv.merge_path_flat(A, xd, y, tiled=True)
    np.testing.assert_array_equal(y.cpu().numpy(), 2.0 * y1)
    # 5. column ids changed in place as well (rotate every column id by one)
    idx2 = ((idx.astype(np.int64) + 1) % 2500).astype(np.int32)
    order = np.concatenate([off[r] + np.argsort(idx2[off[r]:off[r + 1]], kind="stable") for r in range(3000)]) \
        if len(idx2) else np.zeros(0, np.int64)
    A.indices.copy_(torch.as_tensor(idx2[order]).cuda())
    A.values.copy_(torch.as_tensor(val[order]).cuda())
    y.fill_(float("nan"))
    spmv.merge_path_flat(A, xd, y, tiled=True)
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv(off, idx2[order], val[order], x))


def test_two_plans_on_two_streams_interleaved(oracle):
    """Launches after the first on a plan use programmatic dependent launch instead of a cooperative
    launch -- only while no band-tiled launch of ANOTHER stream can still be running (two partially
    resident grids could hold each other's SMs); otherwise they fall back to the cooperative launch.
    Two plans driven from two streams without any synchronisation in between must neither hang nor
    disturb each other's results."""
    from loops_b200 import csr_t
    from loops_b200.algorithms import spmv
    mats = []
    for seed in (31, 32):
        off, idx, val = random_csr(6000, 5000, 0.004, seed=seed, exact=True, empty_every=11, heavy_row=(7, 3000))
        x = oracle.x_recipe_int(5000)
        A = csr_t(6000, 5000, off, idx, val)
        mats.append((A, torch.as_tensor(x).cuda(), oracle.spmv(off, idx, val, x)))
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    ys = [[torch.full((6000,), float("nan"), device="cuda") for _ in range(8)] for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(8):
        for k in (0, 1):
            A, xd, _ = mats[k]
            with torch.cuda.stream(streams[k]):
                spmv.merge_path_flat(A, xd, ys[k][rep], stream=streams[k], sync=False, tiled=True)
    torch.cuda.synchronize()
    for k in (0, 1):
        for rep in range(8):
            np.testing.assert_array_equal(ys[k][rep].cpu().numpy(), mats[k][2])
    # and back on one stream: a chain of dependent launches (x of launch n+1 is y of launch n)
    A, xd, ref = mats[0]
    sq = csr_t(*_square_power_matrix())
    v = torch.ones(sq.rows, device="cuda")
    w = torch.empty_like(v)
    want = np.ones(sq.rows, np.float32)
    o, i_, vv = sq.host()
    for _ in range(6):
        spmv.merge_path_flat(sq, v, w, sync=False, tiled=True)
        v, w = w, v
        want = oracle.spmv(o, i_, vv, want)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(v.cpu().numpy(), want)


def _square_power_matrix():
    """3000 x 3000, entries 1/2 or 1/4 on a few bands: repeated products stay exactly representable."""
    n = 3000
    off = np.zeros(n + 1, np.int32)
    idx, val = [], []
    for r in range(n):
        cols = sorted({r, (r * 7 + 3) % n, (r + 1) % n, (r * 13 + 11) % n})
        idx += cols
        val += [0.25 if (c + r) % 2 else 0.5 for c in cols]
        off[r + 1] = len(idx)
    return n, n, off, np.asarray(idx, np.int32), np.asarray(val, np.float32)
