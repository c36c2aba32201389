"""CPU tests of the band-tiled plan builder (loopsb_tiled_image_build_host):
the image is walked by a host emulation of the kernel (tests/tiled_emul.py)
that checks every format invariant the sm_100a kernel relies on -- each nonzero
present exactly once, rows confined to their warp, the x-ring protocol free of
deadlock, collision flags exact -- and the resulting y is compared bit for bit
with the oracle (reference util/reference.hxx:61-76 restated) on exactly
representable inputs. No device is involved."""
import numpy as np
import pytest

from helpers import load_chesapeake, random_csr
from tiled_emul import build_image, emulate


@pytest.fixture(scope="module")
def lib():
    from loops_b200 import _lib
    return _lib.load()


def _check(lib, oracle, rows, cols, off, idx, val, x, geometry, expect_flags=None):
    rc, img = build_image(lib, rows, cols, off, idx, val, geometry)
    assert rc == 0, lib.loopsb_last_error()
    g = img["g"]
    nnz = int(off[-1])
    assert g["real_entries"] == nnz
    assert g["total_steps"] * 128 == nnz + g["pad_entries"]
    y, tr = emulate(img, x, rows, cols)
    # every (row, col, value) exactly once
    rows_of = np.repeat(np.arange(rows), np.diff(off))
    want = np.stack([rows_of, idx.astype(np.int64), val.view(np.uint32).astype(np.int64)], axis=1)
    order = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
    np.testing.assert_array_equal(order(tr), order(want))
    ref = oracle.spmv(off, idx, val, x)
    np.testing.assert_array_equal(y, ref)
    if expect_flags is not None:
        assert (g["long_steps"] > 0) == expect_flags
        assert g["flagged_entries"] == 0        # the packer keeps every row in one range of cells
    return img


GEOMS = [
    (3, 2, 4, 8, 2, 2),      # tiny bands: many band switches, x ring depth 2
    (5, 1, 2, 16, 3, 3),     # single column part, ring depth 3
    (2, 4, 8, 64, 2, 3),     # wide parts
    (7, 3, 16, 4, 2, 2),     # narrowest legal band
]


@pytest.mark.parametrize("geometry", GEOMS)
def test_chesapeake(lib, oracle, geometry):
    c = load_chesapeake()
    _check(lib, oracle, 39, 39, c["off"], c["idx"], c["val"], c["x"], geometry)


@pytest.mark.parametrize("geometry", GEOMS)
@pytest.mark.parametrize("shape", [(200, 150, 0.05), (64, 1000, 0.01), (500, 37, 0.3)])
def test_random_exact(lib, oracle, geometry, shape):
    rows, cols, dens = shape
    off, idx, val = random_csr(rows, cols, dens, seed=rows + cols, empty_every=7, exact=True)
    x = oracle.x_recipe_int(cols)
    _check(lib, oracle, rows, cols, off, idx, val, x, geometry)


def test_dense_rows_set_flags(lib, oracle):
    """Rows much longer than a lane's 4 entries inside one band: their runs cover
    three or more lanes, the step must carry the long-run bit (segmented scan)."""
    off, idx, val = random_csr(40, 300, 0.02, seed=5, heavy_row=(3, 300), exact=True)
    x = oracle.x_recipe_int(300)
    _check(lib, oracle, 40, 300, off, idx, val, x, (2, 1, 2, 128, 2, 2), expect_flags=True)
    _check(lib, oracle, 40, 300, off, idx, val, x, (2, 2, 4, 32, 2, 3), expect_flags=True)


def test_empty_bands_and_window_padding(lib, oracle):
    """Columns only at the far ends of each part: most bands are empty for most
    warps, so consecutive non-empty bands are further apart than the x ring is
    deep and the builder has to pad to a step boundary."""
    rows, cols = 96, 512
    rng = np.random.default_rng(11)
    off, idx, val = [0], [], []
    for r in range(rows):
        c = np.unique(np.concatenate([rng.integers(0, 6, 2), rng.integers(500, 512, 2),
                                      rng.integers(250, 262, 1)])).astype(np.int32)
        idx.append(c)
        val.append((rng.integers(1, 17, len(c)) / 8.0).astype(np.float32))
        off.append(off[-1] + len(c))
    off, idx, val = np.array(off, np.int32), np.concatenate(idx), np.concatenate(val)
    x = oracle.x_recipe_int(cols)
    for geometry in [(2, 1, 2, 8, 2, 2), (2, 2, 2, 8, 3, 2), (1, 1, 1, 4, 2, 2)]:
        img = _check(lib, oracle, rows, cols, off, idx, val, x, geometry)
        assert img["g"]["pad_entries"] > 0


def test_ragged_and_degenerate_shapes(lib, oracle):
    # more row blocks than rows, more column parts than columns, one row, one column
    for rows, cols, dens in [(3, 5, 0.9), (1, 64, 0.5), (64, 1, 1.0), (2, 2, 1.0)]:
        off, idx, val = random_csr(rows, cols, dens, seed=rows * 31 + cols, exact=True)
        if off[-1] == 0:
            continue
        x = oracle.x_recipe_int(cols)
        _check(lib, oracle, rows, cols, off, idx, val, x, (8, 4, 4, 4, 2, 2))


def test_unsorted_and_duplicate_columns(lib, oracle):
    """CSR rows need not be sorted or duplicate-free for the tiled copy."""
    off = np.array([0, 5, 5, 9], np.int32)
    idx = np.array([7, 2, 7, 0, 2, 1, 1, 1, 3], np.int32)
    val = (np.arange(1, 10) / 8.0).astype(np.float32)
    x = oracle.x_recipe_int(8)
    _check(lib, oracle, 3, 8, off, idx, val, x, (2, 2, 2, 4, 2, 2))


def test_geometry_limits_are_reported(lib):
    off = np.array([0, 1], np.int32)
    idx = np.zeros(1, np.int32)
    val = np.ones(1, np.float32)
    for bad in [(1, 1, 1, 6, 2, 2),          # band width not a multiple of 4
                (1, 1, 1, 40000, 2, 2),      # ring wider than 16 bits of position
                (1, 1, 40, 8, 2, 2),         # too many consumer warps
                (0, 1, 1, 8, 2, 2)]:
        rc, _ = build_image(lib, 1, 1, off, idx, val, bad)
        assert rc == 3, bad                   # LOOPSB_ERR_UNSUPPORTED
    rc, _ = build_image(lib, 40000, 1, np.zeros(40001, np.int32), idx[:0], val[:0], (1, 1, 1, 8, 2, 2))
    assert rc == 3                            # 40000 rows in one block exceed 15 row bits


def test_out_of_range_column_is_rejected(lib):
    off = np.array([0, 1], np.int32)
    rc, _ = build_image(lib, 1, 4, off, np.array([9], np.int32), np.ones(1, np.float32), (1, 1, 1, 4, 2, 2))
    assert rc == 1                            # LOOPSB_ERR_INVALID


def test_powerlaw_sample_of_the_bench_workload(lib, oracle):
    """A 2^12-row cut of the bench generator: padding (step tails, whole
    prefetch groups) stays a small fraction when streams are tens of steps long."""
    from loops_b200 import generate as g
    rows = cols = 1 << 12
    off, idx, val = g.synth_csr(rows, cols, rows * 32)
    off, idx, val = off.numpy(), idx.numpy(), val.numpy()
    x = g.x_recipe(cols).numpy()
    img = _check(lib, oracle, rows, cols, off, idx, val, x, (2, 2, 8, 512, 2, 2))
    assert img["g"]["pad_entries"] < 0.1 * rows * 32


def test_unpacked_steps_can_be_dirty(lib, oracle, monkeypatch):
    """With the packer off, a step that holds the tail of one band and the head of
    the next can see a row twice in separate cell ranges: the dirty bit (general
    y-update path) must be set, and the emulated result must still be right."""
    monkeypatch.setenv("LOOPSB_TILED_PACK", "0")
    off, idx, val = random_csr(20, 64, 0.5, seed=9, exact=True)
    x = oracle.x_recipe_int(64)
    img = _check(lib, oracle, 20, 64, off, idx, val, x, (1, 1, 1, 8, 2, 2))
    assert img["g"]["flagged_entries"] > 0 and img["g"]["flagged_steps"] > 0


@pytest.mark.parametrize("env", [{}, {"LOOPSB_TILED_ROUND": "1"}, {"LOOPSB_TILED_SPLIT": "0"},
                                 {"LOOPSB_TILED_ROUND": "1", "LOOPSB_TILED_SPLIT": "0"}])
def test_random_matrices_and_geometries_under_every_builder_variant(lib, oracle, monkeypatch, env):
    """Seeded random shapes, densities, heavy rows and geometries, through the round-2 builder (streams end at
    their last real step; warp boundaries at row midpoints) and the round-1 rules it can be switched back to:
    every image must satisfy all format invariants of the emulation and reproduce the oracle's y bit for bit."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(2024)
    done = 0
    for trial in range(40):
        rows, cols = int(rng.integers(1, 400)), int(rng.integers(1, 700))
        dens = float(rng.choice([0.005, 0.02, 0.1, 0.4]))
        heavy = (int(rng.integers(0, rows)), int(rng.integers(1, cols + 1))) if rng.random() < 0.5 else None
        off, idx, val = random_csr(rows, cols, dens, seed=1000 + trial, empty_every=int(rng.choice([0, 3, 7])),
                                   heavy_row=heavy, exact=True)
        if off[-1] == 0:
            continue
        cb = int(rng.choice([4, 8, 16, 64, 256]))
        geometry = (int(rng.integers(1, 7)), int(rng.integers(1, 5)), int(rng.choice([1, 2, 4, 8, 12, 16])), cb,
                    int(rng.integers(2, 5)), int(rng.integers(2, 5)))
        x = oracle.x_recipe_int(cols)
        rc, img = build_image(lib, rows, cols, off, idx, val, geometry)
        if rc == 3:        # the control word cannot hold this many starting bands per step: reported, not built
            continue
        assert rc == 0, (geometry, lib.loopsb_last_error())
        y, _ = emulate(img, x, rows, cols)
        np.testing.assert_array_equal(y, oracle.spmv(off, idx, val, x), err_msg=str((trial, rows, cols, geometry)))
        nst = np.diff(img["stream_base"])
        if env.get("LOOPSB_TILED_ROUND") == "1":
            assert np.all(nst % geometry[5] == 0)          # round-1 format: whole prefetch groups
        done += 1
    assert done >= 25


def test_midpoint_split_balances_the_warps_of_the_bench_workload(lib, monkeypatch):
    """Round-2 rule: a row goes to the warp whose share of the nonzeros holds its midpoint. On a 2^17-row cut of
    the bench generator with the production proportions (24 warps, ~1100 rows per warp, rows of up to 1024
    entries, 7168-column bands) the slowest warp of a CTA -- what its finishing time follows -- gets shorter
    and the spread of stream lengths shrinks, against the round-1 rule (boundary at the row's start)."""
    from loops_b200 import generate as g
    rows = cols = 1 << 17
    off, idx, val = g.synth_csr(rows, cols, rows * 32)
    off, idx, val = off.numpy(), idx.numpy(), val.numpy()
    geometry = (5, 4, 24, 7168, 3, 3)
    stats = {}
    for split in ("0", "1"):
        monkeypatch.setenv("LOOPSB_TILED_SPLIT", split)
        rc, img = build_image(lib, rows, cols, off, idx, val, geometry)
        assert rc == 0
        nst = np.diff(img["stream_base"]).reshape(-1, geometry[2])
        stats[split] = (float(nst.max(1).mean()), int(nst.max() - nst.min()))
    assert stats["1"][0] < stats["0"][0] and stats["1"][1] < stats["0"][1], stats
