"""CPU: the Python mirrors of csc_t(csr) / dia_t(csr) (reference
container/csc.hxx:88-102, dia.hxx:135-188) against the oracle's restatements,
which tests/test_oracle_vs_ref.py pins to the reference's own headers; and the
flat_uniform_occupancy layout view against the reference's contract
(container/partitioning.hxx:71-141, unittests/test_layout_flat_partitioner.cu:27-59)."""
import numpy as np
import pytest

from helpers import random_csr

CASES = [(64, 64, 0.1, 1, 0, None), (200, 150, 0.05, 2, 5, None), (97, 300, 0.02, 3, 0, (7, 250)),
         (1, 40, 0.5, 4, 0, None), (300, 17, 0.3, 5, 3, None)]


@pytest.mark.parametrize("rows,cols,dens,seed,empty,heavy", CASES)
def test_csc_and_dia_from_csr(oracle, rows, cols, dens, seed, empty, heavy):
    from loops_b200 import csc_t, csr_t, dia_t
    off, idx, val = random_csr(rows, cols, dens, seed, empty, heavy)
    A = csr_t(rows, cols, off, idx, val, device="cpu")
    csc = csc_t.from_csr(A)
    c_off, c_row, c_val = oracle.csc(rows, cols, off, idx, val)
    np.testing.assert_array_equal(csc.offsets.numpy(), c_off)
    np.testing.assert_array_equal(csc.indices.numpy(), c_row)
    np.testing.assert_array_equal(csc.values.numpy(), c_val)
    assert csc.layout().num_tiles() == cols and csc.layout().num_atoms() == len(idx)
    dia = dia_t.from_csr(A)
    d_off, d_val = oracle.dia(rows, off, idx, val)
    np.testing.assert_array_equal(dia.diag_offsets.numpy(), d_off)
    np.testing.assert_array_equal(dia.values.numpy(), d_val)
    assert dia.stride == rows and dia.num_diagonals == len(d_off)
    assert dia.layout().num_tiles() == rows and dia.layout().pitch == len(d_off)


def test_flat_uniform_occupancy_view():
    """offsets {0,2,2,5,7}, K = 3: tiles [0,3) [3,6) [6,7); base().tile_of unchanged."""
    import torch
    from loops_b200 import layout
    off = torch.tensor([0, 2, 2, 5, 7], dtype=torch.int32)
    base = layout.csr(off, 4, 7)
    lay = layout.flat_uniform_occupancy(3, base)
    assert lay.num_tiles() == 3 and lay.num_atoms() == 7
    assert [lay.tile_begin(t) for t in range(3)] == [0, 3, 6]
    assert [lay.tile_end(t) for t in range(3)] == [3, 6, 7]
    assert [lay.tile_size(t) for t in range(3)] == [3, 3, 1]
    assert lay.base() is base
    d = lay.desc()
    assert d.pitch == 3 and d.num_tiles == 3 and d.num_atoms == 7 and d.offsets == off.data_ptr()


def test_spmm_restatement_is_columnwise_spmv(oracle):
    """algorithms/spmm/thread_mapped.cuh:28-53 restated: column j of A B is the
    validator's SpMV with x = B[:, j] (same sequential order, so the same bits)."""
    off, idx, val = random_csr(120, 90, 0.08, seed=12, empty_every=11)
    B = np.random.default_rng(3).uniform(-2, 2, (90, 7)).astype(np.float32)
    Cm = oracle.spmm(off, idx, val, B)
    for j in range(7):
        np.testing.assert_array_equal(Cm[:, j], oracle.spmv(off, idx, val, np.ascontiguousarray(B[:, j])))
