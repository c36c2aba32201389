"""GPU parity: device format conversions through the C ABI against
the oracle's restatement of the reference's host converters (which the CPU
suite pins to the reference headers themselves, tests/test_oracle_vs_ref.py).
Everything here is integer / copy work: the bar is bit-exact arrays.

Reference: container/coo.hxx:87-98 + detail/convert.hxx:36-60 (csr->coo),
csr.hxx:86-94 (coo->csr), csc.hxx:86-108, ell.hxx:113-145, bcsr.hxx:111-194,
dia.hxx:135-188; round trips as in unittests/test_format_round_trip.cu:131-168."""
import numpy as np
import pytest
import torch

from helpers import load_chesapeake, random_csr

pytestmark = pytest.mark.gpu

CASES = {
    "ragged_empty_rows": dict(rows=257, cols=301, density=0.03, seed=3, empty_every=5),
    "heavy_row": dict(rows=300, cols=4096, density=0.002, seed=4, heavy_row=(7, 3000)),
    "wide": dict(rows=64, cols=5000, density=0.01, seed=5),
    "tall": dict(rows=5000, cols=64, density=0.05, seed=6, empty_every=3),
    "single_row": dict(rows=1, cols=77, density=0.5, seed=7),
    "single_col": dict(rows=90, cols=1, density=0.5, seed=8),
}


def _case(name):
    if name == "chesapeake":
        c = load_chesapeake()
        return 39, 39, c["off"], c["idx"], c["val"]
    kw = dict(CASES[name])
    rows, cols = kw.pop("rows"), kw.pop("cols")
    off, idx, val = random_csr(rows, cols, **kw)
    return rows, cols, off, idx, val


ALL = ["chesapeake"] + sorted(CASES)


def _dev(rows, cols, off, idx, val):
    from loops_b200 import csr_t
    return csr_t(rows, cols, off, idx, val)


def _np(t):
    return t.cpu().numpy()


@pytest.mark.parametrize("name", ALL)
def test_csr_to_coo_rows(oracle, name):
    from loops_b200 import convert
    rows, cols, off, idx, val = _case(name)
    coo = convert.csr_to_coo(_dev(rows, cols, off, idx, val))
    assert np.array_equal(_np(coo.row_indices), oracle.coo_rows(off))
    assert coo.nnzs == int(off[-1]) and np.array_equal(_np(coo.col_indices), idx)


@pytest.mark.parametrize("name", ALL)
def test_csr_to_csc(oracle, name):
    from loops_b200 import convert
    rows, cols, off, idx, val = _case(name)
    csc = convert.csr_to_csc(_dev(rows, cols, off, idx, val))
    c_off, c_row, c_val = oracle.csc(rows, cols, off, idx, val)
    assert np.array_equal(_np(csc.offsets), c_off)
    assert np.array_equal(_np(csc.indices), c_row)
    assert np.array_equal(_np(csc.values), c_val)


@pytest.mark.parametrize("name", ALL)
def test_csr_to_ell(oracle, name):
    from loops_b200 import convert
    rows, cols, off, idx, val = _case(name)
    ell = convert.csr_to_ell(_dev(rows, cols, off, idx, val))
    pitch, e_idx, e_val = oracle.ell(off, idx, val)
    assert ell.pitch == pitch == int(np.diff(off).max())
    assert np.array_equal(_np(ell.indices), e_idx)
    assert np.array_equal(_np(ell.values), e_val)


@pytest.mark.parametrize("name", ALL)
@pytest.mark.parametrize("shape", [(4, 4), (2, 3), (8, 1), (1, 1)])
def test_csr_to_bcsr_f32(oracle, name, shape):
    from loops_b200 import convert
    R, Cc = shape
    rows, cols, off, idx, val = _case(name)
    b = convert.csr_to_bcsr(_dev(rows, cols, off, idx, val), R, Cc)
    b_off, b_col, b_val = oracle.bcsr(R, Cc, rows, cols, off, idx, val)
    assert b.num_blocks == len(b_col)
    assert np.array_equal(_np(b.block_offsets), b_off)
    assert np.array_equal(_np(b.block_col_indices), b_col)
    assert np.array_equal(_np(b.values), b_val)


def test_csr_to_bcsr_bf16_rounds_like_the_oracle(oracle):
    from loops_b200 import convert
    rows, cols, off, idx, val = _case("ragged_empty_rows")
    b = convert.csr_to_bcsr(_dev(rows, cols, off, idx, val), 4, 4, value_dtype=torch.bfloat16)
    _, _, b_val = oracle.bcsr(4, 4, rows, cols, off, idx, val)
    want = np.array([oracle.bf16_round(float(v)) for v in b_val], np.float32)
    assert np.array_equal(_np(b.values.float()), want)


@pytest.mark.parametrize("name", ["chesapeake", "ragged_empty_rows", "tall", "single_row", "single_col"])
def test_csr_to_dia(oracle, name):
    from loops_b200 import convert
    rows, cols, off, idx, val = _case(name)
    d = convert.csr_to_dia(_dev(rows, cols, off, idx, val))
    d_off, d_val = oracle.dia(rows, off, idx, val)
    assert d.num_diagonals == len(d_off)
    assert np.array_equal(_np(d.diag_offsets), d_off)
    assert np.array_equal(_np(d.values), d_val)


@pytest.mark.parametrize("name", ALL)
def test_coo_to_csr_from_shuffled_triples(name):
    """csr -> coo -> shuffle -> csr is the identity (reference csr.hxx:86-94
    sorts by (row, col); unittests/test_format_round_trip.cu:131-168)."""
    from loops_b200 import convert, coo_t
    rows, cols, off, idx, val = _case(name)
    A = _dev(rows, cols, off, idx, val)
    coo = convert.csr_to_coo(A)
    g = torch.Generator(device="cpu").manual_seed(11)
    perm = torch.randperm(coo.nnzs, generator=g).cuda()
    shuffled = coo_t.from_tensors(rows, cols, coo.row_indices[perm].contiguous(),
                                  coo.col_indices[perm].contiguous(), coo.values[perm].contiguous())
    back = convert.coo_to_csr(shuffled)
    assert np.array_equal(_np(back.offsets), off)
    assert np.array_equal(_np(back.indices), idx)
    assert np.array_equal(_np(back.values), val)


def test_transpose_twice_is_identity():
    """csr -> csc is the structural transpose: applied to the transpose it gives
    the matrix back (csc.hxx:96-108)."""
    from loops_b200 import convert, csr_t
    rows, cols, off, idx, val = _case("heavy_row")
    A = _dev(rows, cols, off, idx, val)
    T = convert.csr_to_csc(A)                                  # arrays of A^T in CSR form
    At = csr_t.from_tensors(cols, rows, T.offsets, T.indices, T.values)
    back = convert.csr_to_csc(At)
    assert np.array_equal(_np(back.offsets), off)
    assert np.array_equal(_np(back.indices), idx)
    assert np.array_equal(_np(back.values), val)


def test_empty_matrix_conversions():
    from loops_b200 import convert, csr_t
    off = np.zeros(6, np.int32)
    A = csr_t(5, 4, off, np.zeros(0, np.int32), np.zeros(0, np.float32))
    assert convert.csr_to_coo(A).nnzs == 0
    assert np.array_equal(_np(convert.csr_to_csc(A).offsets), np.zeros(5, np.int32))
    assert convert.csr_to_ell(A).pitch == 0
    b = convert.csr_to_bcsr(A, 2, 2)
    assert b.num_blocks == 0 and np.array_equal(_np(b.block_offsets), np.zeros(4, np.int32))
    assert convert.csr_to_dia(A).num_diagonals == 0


def test_converted_containers_feed_the_spmv_kernels(oracle):
    """The device-built COO / ELL / CSC / BCSR / DIA containers run through their
    SpMV entry points and agree with the oracle SpMV (exact inputs: bit-equal)."""
    from loops_b200 import convert
    from loops_b200.algorithms import spmv
    rows, cols = 300, 280
    off, idx, val = random_csr(rows, cols, 0.04, 21, empty_every=7, exact=True)
    x = oracle.x_recipe_int(cols)
    want = oracle.spmv(off, idx, val, x)
    A = _dev(rows, cols, off, idx, val)
    xd = torch.as_tensor(x).cuda()

    def run(fn, M, xin=xd):
        y = torch.zeros(rows, dtype=torch.float32, device="cuda")
        fn(M, xin, y)
        return y.cpu().numpy()

    assert np.array_equal(run(spmv.coo_thread_mapped, convert.csr_to_coo(A)), want)
    assert np.array_equal(run(spmv.ell_thread_mapped, convert.csr_to_ell(A)), want)
    assert np.array_equal(run(spmv.csc_thread_mapped, convert.csr_to_csc(A)), want)
    assert np.array_equal(run(spmv.dia_thread_mapped, convert.csr_to_dia(A)), want)
    B = convert.csr_to_bcsr(A, 4, 4)
    assert np.array_equal(run(spmv.bcsr_thread_mapped, B, B.padded_x(xd)), want)


def test_full_size_config2_round_trips():
    """BASELINE config 2 (2^20 rows / 2^25 nnz): size-independent properties --
    coo -> csr of the expanded triples gives the matrix back, the transpose of the
    transpose is the matrix, ELL keeps every entry in its row-major slot."""
    from loops_b200 import convert, csr_t, generate as g
    rows = cols = 1 << 20
    off, idx, val = g.synth_csr(rows, cols, 1 << 25, device="cuda")
    A = csr_t.from_tensors(rows, cols, off, idx, val)
    coo = convert.csr_to_coo(A)
    deg = (off[1:] - off[:-1]).long()
    assert torch.equal(torch.bincount(coo.row_indices.long(), minlength=rows), deg)
    assert bool((coo.row_indices[1:] >= coo.row_indices[:-1]).all())
    back = convert.coo_to_csr(coo)
    assert torch.equal(back.offsets, off) and torch.equal(back.indices, idx) and torch.equal(back.values, val)
    T = convert.csr_to_csc(A)
    At = csr_t.from_tensors(cols, rows, T.offsets, T.indices, T.values)
    tt = convert.csr_to_csc(At)
    assert torch.equal(tt.offsets, off) and torch.equal(tt.indices, idx) and torch.equal(tt.values, val)
    assert convert.csr_max_degree(A) == int(deg.max().item())
    B = convert.csr_to_bcsr(A, 4, 4)
    assert int(B.block_offsets[-1].item()) == B.num_blocks
    # every block holds at least one entry and the payload sums to the matrix's
    assert abs(float(B.values.double().sum().item()) - float(val.double().sum().item())) < 1e-3
    per_block = B.values.view(-1, 16).ne(0).sum(1)
    assert int(per_block.min().item()) >= 1 and int(per_block.sum().item()) == 1 << 25


def test_market_file_to_device_csr(tmp_path):
    """Matrix Market file -> device CSR with the sort and row compression on the
    device equals the host path (symmetric mirrors, a duplicate entry kept in file
    order, empty rows; reference container/market.hxx:100-289 + csr.hxx:86-94)."""
    from loops_b200 import market
    path = tmp_path / "m.mtx"
    path.write_text("%%MatrixMarket matrix coordinate real symmetric\n% comment\n6 6 7\n"
                    "3 1 2.5\n1 1 1.0\n5 2 -3.0\n6 5 4.0\n3 1 7.0\n2 2 0.5\n6 1 9.0\n")
    R, Cc, r, c, v = market.load_coo(str(path))
    off, idx, val = market.coo_to_csr(R, Cc, r, c, v)
    A = market.load_csr(str(path), device="cuda")
    assert (A.rows, A.cols, A.nnzs) == (6, 6, len(idx))
    assert np.array_equal(_np(A.offsets), off)
    assert np.array_equal(_np(A.indices), idx)
    assert np.array_equal(_np(A.values), val)
