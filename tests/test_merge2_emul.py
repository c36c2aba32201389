"""CPU checks of the second-generation merge-path kernel's ALGORITHM through its
host emulation (tests/merge2_emul.py): flags / chunks / fold / carries produce the
reference's y on the battery, on ragged and empty-row matrices, on rows that span
many tiles, and on ELL slabs. The GPU tests then compare the kernel itself with
this emulation bit for bit."""
import numpy as np
import pytest

from helpers import random_csr
from merge2_emul import merge_coords, spmv_merge2


def _exact_x(n, seed=3):
    return np.random.default_rng(seed).integers(1, 11, size=n).astype(np.float32)


def test_coords_match_oracle(oracle, battery):
    for m in battery:
        lay = oracle.csr_layout(m["off"])
        ref = oracle.emit(lay, 0)["coords"]
        T, A = len(m["off"]) - 1, int(m["off"][-1])
        mine = merge_coords(np.asarray(m["off"][1:], np.int64), T, A)
        assert [tuple(r) for r in ref.tolist()] == [tuple(int(v) for v in c) for c in mine], m["name"]


def test_battery_exact(oracle, battery):
    for m in battery:
        val = (np.round(m["val"] * 8) / 8).astype(np.float32)
        x = _exact_x(m["cols"])
        ref = oracle.spmv(m["off"], m["idx"], val, x)
        got = spmv_merge2(m["off"], m["idx"], val, x)
        assert np.array_equal(got, ref), m["name"]


def test_battery_float_close(oracle, battery):
    for m in battery:
        ref = oracle.spmv_f64(m["off"], m["idx"], m["val"], m["x"])
        l1 = oracle.row_l1(m["off"], m["idx"], m["val"], m["x"])
        got = spmv_merge2(m["off"], m["idx"], m["val"], m["x"])
        assert np.all(np.abs(got - ref) <= 1e-6 * np.maximum(l1, 1e-30)), m["name"]


@pytest.mark.parametrize("case", [
    dict(rows=300, cols=257, density=0.05, empty_every=3),
    dict(rows=64, cols=5000, density=0.6),                       # rows longer than a tile's chunk span
    dict(rows=5, cols=9000, density=0.9),                        # rows spanning several tiles, spill chunks
    dict(rows=4000, cols=40, density=0.01, empty_every=2),       # tiles that are mostly row ends
    dict(rows=1, cols=3000, density=1.0),
    dict(rows=2500, cols=64, density=0.0),                       # nothing stored
    dict(rows=700, cols=700, density=0.02, heavy_row=(350, 699)),
])
def test_shapes_exact(oracle, case):
    off, idx, val = random_csr(seed=11, exact=True, **case)
    x = _exact_x(case["cols"])
    ref = oracle.spmv(off, idx, val, x)
    got = spmv_merge2(off, idx, val, x)
    assert np.array_equal(got, ref)


def test_threads_256_same_result(oracle):
    off, idx, val = random_csr(900, 1200, 0.03, seed=5, exact=True, empty_every=7)
    x = _exact_x(1200)
    ref = oracle.spmv(off, idx, val, x)
    assert np.array_equal(spmv_merge2(off, idx, val, x, threads=256), ref)


def test_ell_slab(oracle, battery):
    for m in battery[:5]:
        val = (np.round(m["ell_val"] * 8) / 8).astype(np.float32)
        x = _exact_x(m["cols"])
        ref = oracle.spmv_ell(m["rows"], m["ell_pitch"], m["ell_idx"], val, x)
        got = spmv_merge2(None, m["ell_idx"], val, x, ell_pitch=m["ell_pitch"])
        assert np.array_equal(got, ref), m["name"]
