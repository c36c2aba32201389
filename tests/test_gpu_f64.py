"""GPU parity: fp64 CSR SpMV through the C ABI against the oracle's
restatement of reference::spmv<double> (util/reference.hxx:61-76; pinned to the
reference header itself by tests/test_oracle_vs_ref.py::test_validator_in_double).
Bars: thread_mapped is the same sequential arithmetic -> bit-equal; the merge-path
kernel (merge_path_flat / work_oriented / group_mapped) sums a row in the same
order inside a tile and joins tiles with fp64 atomics -> within 1e-13 of the row's
L1 mass, and bit-equal on exactly representable inputs."""
import numpy as np
import pytest
import torch

from helpers import load_chesapeake, random_csr

pytestmark = pytest.mark.gpu
SCHEDULES = ["merge_path_flat", "work_oriented", "thread_mapped", "group_mapped"]
CASES = {
    "ragged": dict(rows=257, cols=301, density=0.03, seed=3, empty_every=5),
    "heavy_row": dict(rows=300, cols=4096, density=0.002, seed=4, heavy_row=(7, 3000)),
    "tall": dict(rows=5000, cols=64, density=0.05, seed=6, empty_every=3),
    "single_row": dict(rows=1, cols=777, density=0.9, seed=7),
    "many_tiles": dict(rows=4000, cols=3000, density=0.01, seed=9),
}


def _run(schedule, off, idx, val64, x64, rows, cols):
    from loops_b200.algorithms import spmv
    y = torch.full((rows,), float("nan"), dtype=torch.float64, device="cuda")
    spmv.spmv_f64(schedule, torch.as_tensor(off).cuda(), torch.as_tensor(idx).cuda(), torch.as_tensor(val64).cuda(),
                  torch.as_tensor(x64).cuda(), y, rows, cols)
    return y.cpu().numpy()


@pytest.mark.parametrize("schedule", SCHEDULES)
@pytest.mark.parametrize("name", sorted(CASES))
def test_f64_random_values(oracle, schedule, name):
    kw = dict(CASES[name]); rows, cols = kw.pop("rows"), kw.pop("cols")
    off, idx, _ = random_csr(rows, cols, **kw)
    rng = np.random.default_rng(17)
    val64 = rng.uniform(-1.5, 1.5, len(idx)); x64 = rng.uniform(-2.0, 2.0, cols)
    want = oracle.spmv_d(off, idx, val64, x64)
    got = _run(schedule, off, idx, val64, x64, rows, cols)
    assert np.all(np.isfinite(got))
    if schedule == "thread_mapped":
        np.testing.assert_array_equal(got, want)
    else:
        l1 = np.zeros(rows); np.add.at(l1, np.repeat(np.arange(rows), np.diff(off)), np.abs(val64 * x64[idx]))
        assert np.max(np.abs(got - want) / np.maximum(l1, 1e-300)) <= 1e-13


@pytest.mark.parametrize("schedule", SCHEDULES)
def test_f64_chesapeake_known_answers(schedule):
    c = load_chesapeake()
    got = _run(schedule, c["off"], c["idx"], c["val"].astype(np.float64), c["x"].astype(np.float64), 39, 39)
    np.testing.assert_array_equal(got, c["y"].astype(np.float64))
    assert float(got.sum()) == 1794.0


@pytest.mark.parametrize("schedule", SCHEDULES)
def test_f64_exact_inputs_are_bit_equal(oracle, schedule):
    rows, cols = 3000, 2048
    off, idx, val = random_csr(rows, cols, 0.02, 31, empty_every=11, heavy_row=(5, 1500), exact=True)
    x = oracle.x_recipe_int(cols)
    want = oracle.spmv_d(off, idx, val.astype(np.float64), x.astype(np.float64))
    np.testing.assert_array_equal(_run(schedule, off, idx, val.astype(np.float64), x.astype(np.float64), rows, cols), want)


def test_f64_empty_and_argument_checks():
    from loops_b200 import _lib
    from loops_b200.algorithms import spmv
    off = torch.zeros(6, dtype=torch.int32, device="cuda")
    e_i = torch.zeros(0, dtype=torch.int32, device="cuda"); e_d = torch.zeros(0, dtype=torch.float64, device="cuda")
    y = torch.full((5,), float("nan"), dtype=torch.float64, device="cuda")
    spmv.spmv_f64("merge_path_flat", off, e_i, e_d, torch.ones(4, dtype=torch.float64, device="cuda"), y, 5, 4)
    assert torch.equal(y, torch.zeros_like(y))
    with pytest.raises(ValueError):
        spmv.spmv_f64("merge_path_flat", off, e_i, e_d.float(), torch.ones(4, dtype=torch.float64, device="cuda"), y, 5, 4)
