"""GPU parity of the second-generation merge-path kernel (spmv_merge2.cuh, the
default for merge_path_flat / work_oriented / ell_merge_path on the CSR arrays):
bit-for-bit against its host emulation (tests/merge2_emul.py -- same adds in the
same order, so this holds for ANY float input), exact against the oracle on
exactly representable inputs, and the first-generation kernel still serves
pointers that are not 16-byte aligned."""
import numpy as np
import pytest
import torch

from helpers import random_csr
from merge2_emul import spmv_merge2

pytestmark = pytest.mark.gpu


def _gpu(off, idx, val, x, rows, cols, schedule="merge_path_flat"):
    from loops_b200 import csr_t
    from loops_b200.algorithms import spmv
    A = csr_t(rows, cols, off, idx, val)
    xd = torch.as_tensor(x).cuda()
    y = torch.full((rows,), float("nan"), dtype=torch.float32, device="cuda")
    if schedule == "merge_path_flat":
        spmv.merge_path_flat(A, xd, y, tiled=False)
    else:
        spmv.BY_NAME[schedule](A, xd, y)
    return y.cpu().numpy()


def test_battery_bit_equal_to_emulation(battery):
    for b in battery:
        y = _gpu(b["off"], b["idx"], b["val"], b["x"], b["rows"], b["cols"])
        emu = spmv_merge2(b["off"], b["idx"], b["val"], b["x"])
        np.testing.assert_array_equal(y, emu, err_msg=b["name"])


@pytest.mark.parametrize("case", [
    dict(rows=300, cols=257, density=0.05, empty_every=3),
    dict(rows=64, cols=5000, density=0.6),
    dict(rows=5, cols=9000, density=0.9),                        # rows spanning tiles; 257th chunk
    dict(rows=4000, cols=40, density=0.01, empty_every=2),       # tiles of mostly row ends
    dict(rows=1, cols=3000, density=1.0),
    dict(rows=700, cols=700, density=0.02, heavy_row=(350, 699)),
    dict(rows=20000, cols=3000, density=0.004),                  # > 2 tiles per CTA (pipeline steady state)
])
def test_shapes_float_bit_equal_to_emulation(case):
    off, idx, val = random_csr(seed=21, **case)
    x = np.random.default_rng(4).uniform(-1, 1, case["cols"]).astype(np.float32)
    y = _gpu(off, idx, val, x, case["rows"], case["cols"])
    np.testing.assert_array_equal(y, spmv_merge2(off, idx, val, x))


def test_many_tiles_exact(oracle):
    """Enough tiles that every CTA runs its software pipeline for several rounds."""
    from loops_b200 import generate as g
    rows = cols = 1 << 17
    off, idx, val = g.synth_csr(rows, cols, rows * 24)
    x = g.x_recipe(cols)
    off, idx, val, x = off.numpy(), idx.numpy(), val.numpy(), x.numpy()
    ref = oracle.spmv(off, idx, val, x)
    for sched in ("merge_path_flat", "work_oriented"):
        np.testing.assert_array_equal(_gpu(off, idx, val, x, rows, cols, sched), ref, err_msg=sched)


def test_misaligned_arrays_take_the_first_generation_kernel(oracle):
    """indices/values views that start 4 bytes into an allocation (the bulk
    copies must not gather through the words in front of the arrays)."""
    from loops_b200 import csr_t
    from loops_b200.algorithms import spmv
    off, idx, val = random_csr(3000, 2000, 0.01, seed=9, exact=True)
    x = np.random.default_rng(1).integers(1, 11, 2000).astype(np.float32)
    ref = oracle.spmv(off, idx, val, x)
    big_i = torch.full((len(idx) + 8,), 2 ** 30, dtype=torch.int32, device="cuda")   # poison around the view
    big_v = torch.full((len(val) + 8,), float("nan"), dtype=torch.float32, device="cuda")
    for lead in (1, 2, 3):
        big_i[lead:lead + len(idx)] = torch.as_tensor(idx).cuda()
        big_v[lead:lead + len(val)] = torch.as_tensor(val).cuda()
        A = csr_t.from_tensors(3000, 2000, torch.as_tensor(off).cuda(), big_i[lead:lead + len(idx)],
                               big_v[lead:lead + len(val)])
        y = torch.full((3000,), float("nan"), device="cuda")
        spmv.merge_path_flat(A, torch.as_tensor(x).cuda(), y, tiled=False)
        np.testing.assert_array_equal(y.cpu().numpy(), ref, err_msg=f"lead {lead}")
        big_i.fill_(2 ** 30)
        big_v.fill_(float("nan"))


def test_ell_merge_path_bit_equal_to_emulation(battery):
    from loops_b200 import csr_t, ell_t
    from loops_b200.algorithms import spmv
    for b in battery[:6]:
        A = csr_t(b["rows"], b["cols"], b["off"], b["idx"], b["val"])
        E = ell_t.from_csr(A)
        xd = torch.as_tensor(b["x"]).cuda()
        y = torch.full((b["rows"],), float("nan"), device="cuda")
        spmv.ell_merge_path(E, xd, y)
        emu = spmv_merge2(None, E.indices.cpu().numpy(), E.values.cpu().numpy(), b["x"], ell_pitch=E.pitch)
        np.testing.assert_array_equal(y.cpu().numpy(), emu, err_msg=b["name"])
