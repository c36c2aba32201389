"""GPU parity: every SpMV entry point, through the C ABI, against the oracle.

Bars: bit-exact y where the arithmetic is exactly representable or sequential
(integer x with k/8 values; thread_mapped on any input); otherwise
|y - y64| <= 1e-6 * L1(row) against the f64-accumulating reference, AND
element-wise |y - y64| <= 1e-6 * |y64| on every row whose |y64| is at least half
its L1 mass (north_star tolerance; rows dominated by cancellation have no
meaningful element-wise relative error in fp32) --
plus the reference's own acceptance rules (count_errors == 0 and a Wilkinson
verdict of NOT_A_BUG, util/reference.hxx:116-131,278-337)."""
import numpy as np
import pytest
import torch

from helpers import load_chesapeake, random_csr

pytestmark = pytest.mark.gpu

REL_TOL = 1e-6        # north_star: y within 1e-6 relative for fp32
CSR_KERNELS = ["merge_path_flat", "thread_mapped", "group_mapped", "work_oriented"]


def _run_csr(name, off, idx, val, x, rows, cols, poison=True):
    from loops_b200 import csr_t
    from loops_b200.algorithms import spmv
    A = csr_t(rows, cols, off, idx, val)
    xd = torch.as_tensor(x).cuda()
    y = torch.full((rows,), float("nan") if poison else 0.0, dtype=torch.float32, device="cuda")
    spmv.BY_NAME[name](A, xd, y)
    return y.cpu().numpy()


def _assert_close(oracle, off, idx, val, x, y, label):
    """rel err <= 1e-6 vs the f64 reference scaled by the row's L1 mass (the
    quantity fp32 summation error is proportional to), and the reference's
    own two validators."""
    y64 = oracle.spmv_f64(off, idx, val, x).astype(np.float64)
    l1 = np.maximum(oracle.row_l1(off, idx, val, x).astype(np.float64), 1e-30)
    rel = np.abs(y.astype(np.float64) - y64) / l1
    assert np.all(np.isfinite(y)), label
    assert rel.max() <= REL_TOL, (label, float(rel.max()))
    # element-wise |y - y64| <= 1e-6 |y64| wherever the row is not dominated by cancellation
    # (|y64| >= half its L1 mass); rows below that bar are held to the L1-scaled bound only,
    # which is what fp32 summation error is proportional to ("Tolerance")
    well = np.abs(y64) >= 0.5 * l1
    if well.any():
        elem = np.abs(y.astype(np.float64) - y64)[well] / np.abs(y64[well])
        assert elem.max() <= REL_TOL, (label, "element-wise", float(elem.max()))
    assert oracle.count_errors(y, oracle.spmv(off, idx, val, x)) == 0, label
    rep = oracle.rigorous(off, idx, val, x, y)
    assert rep.gpu_overruns == 0, (label, "POTENTIAL_BUG")


@pytest.mark.parametrize("kernel", CSR_KERNELS)
def test_chesapeake_config1(oracle, kernel):
    """BASELINE config 1: Dimensions 39 x 39 (340), Errors: 0 -- and exact."""
    c = load_chesapeake()
    y = _run_csr(kernel, c["off"], c["idx"], c["val"], c["x"], 39, 39)
    np.testing.assert_array_equal(y, c["y"])
    assert float(y.sum()) == 1794.0
    assert oracle.count_errors(y, c["y"]) == 0


@pytest.mark.parametrize("kernel", CSR_KERNELS)
def test_battery_csr(oracle, battery, kernel):
    """unittests/test_spmv_csr.cu:32-73 battery; tolerance tightened from the
    reference's atol 1e-3 / rtol 1e-4 to 1e-6 relative."""
    for b in battery:
        y = _run_csr(kernel, b["off"], b["idx"], b["val"], b["x"], b["rows"], b["cols"])
        _assert_close(oracle, b["off"], b["idx"], b["val"], b["x"], y, (kernel, b["name"]))
        if kernel == "thread_mapped":      # sequential, un-fused: bit-exact on any input
            np.testing.assert_array_equal(y, b["y"])


def test_battery_coo_ell(oracle, battery):
    """unittests/test_spmv_coo.cu:24, test_spmv_ell.cu:21-38."""
    from loops_b200 import csr_t, coo_t, ell_t
    from loops_b200.algorithms import spmv
    for b in battery:
        A = csr_t(b["rows"], b["cols"], b["off"], b["idx"], b["val"])
        x = torch.as_tensor(b["x"]).cuda()
        coo, ell = coo_t.from_csr(A), ell_t.from_csr(A)
        np.testing.assert_array_equal(coo.row_indices.cpu().numpy(), b["coo_rows"])
        np.testing.assert_array_equal(ell.indices.cpu().numpy(), b["ell_idx"])
        np.testing.assert_array_equal(ell.values.cpu().numpy(), b["ell_val"])
        for fn, M in ((spmv.coo_thread_mapped, coo), (spmv.ell_thread_mapped, ell), (spmv.ell_merge_path, ell)):
            y = torch.full((b["rows"],), float("nan"), device="cuda")
            fn(M, x, y)
            _assert_close(oracle, b["off"], b["idx"], b["val"], b["x"], y.cpu().numpy(), (fn.__name__, b["name"]))
        y = torch.empty(b["rows"], device="cuda")
        spmv.ell_thread_mapped(ell, x, y)
        np.testing.assert_array_equal(y.cpu().numpy(), b["y"])      # sequential -> exact


def test_battery_csc_dia_flat_original(oracle, battery):
    """The remaining in-tree entry points (unittests/test_spmv_csc.cu, test_spmv_dia.cu,
    test_spmv_partitioned.cu; algorithms/spmv/original.cuh): containers equal the
    oracle's conversions (pinned to the reference headers on CPU), y within the
    north_star tolerance; the sequential kernels (dia, original) bit-exact."""
    from loops_b200 import csc_t, csr_t, dia_t
    from loops_b200.algorithms import spmv
    for b in battery:
        rows, cols = b["rows"], b["cols"]
        A = csr_t(rows, cols, b["off"], b["idx"], b["val"])
        x = torch.as_tensor(b["x"]).cuda()
        label = lambda n: (n, b["name"])
        csc = csc_t.from_csr(A)
        c_off, c_row, c_val = oracle.csc(rows, cols, b["off"], b["idx"], b["val"])
        np.testing.assert_array_equal(csc.offsets.cpu().numpy(), c_off)
        np.testing.assert_array_equal(csc.indices.cpu().numpy(), c_row)
        np.testing.assert_array_equal(csc.values.cpu().numpy(), c_val)
        y = torch.full((rows,), float("nan"), device="cuda")
        spmv.csc_thread_mapped(csc, x, y)
        _assert_close(oracle, b["off"], b["idx"], b["val"], b["x"], y.cpu().numpy(), label("csc_thread_mapped"))
        for K in (1, 8, 13):
            y = torch.full((rows,), float("nan"), device="cuda")
            spmv.flat_partitioned(A, x, y, K=K)
            _assert_close(oracle, b["off"], b["idx"], b["val"], b["x"], y.cpu().numpy(), label(f"flat_partitioned<{K}>"))
        y = torch.full((rows,), float("nan"), device="cuda")
        spmv.original(A, x, y)
        np.testing.assert_array_equal(y.cpu().numpy(), b["y"])
        d_off, d_val = oracle.dia(rows, b["off"], b["idx"], b["val"])
        if len(d_off) * rows <= (1 << 24):          # DIA of a scattered matrix is huge; keep it sane
            dia = dia_t.from_csr(A)
            np.testing.assert_array_equal(dia.diag_offsets.cpu().numpy(), d_off)
            np.testing.assert_array_equal(dia.values.cpu().numpy(), d_val)
            y = torch.full((rows,), float("nan"), device="cuda")
            spmv.dia_thread_mapped(dia, x, y)
            np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv_dia(rows, cols, d_off, d_val, b["x"]))
            np.testing.assert_array_equal(y.cpu().numpy(), b["y"])      # ascending diagonals == ascending columns


def test_csc_flat_exact_inputs_and_banded_dia(oracle):
    from loops_b200 import csc_t, csr_t, dia_t
    from loops_b200.algorithms import spmv
    off, idx, val = random_csr(3000, 2500, 0.004, seed=21, empty_every=9, heavy_row=(5, 2000), exact=True)
    x = oracle.x_recipe_int(2500)
    want = oracle.spmv(off, idx, val, x)
    A = csr_t(3000, 2500, off, idx, val)
    xd = torch.as_tensor(x).cuda()
    for fn, M, kw in ((spmv.csc_thread_mapped, csc_t.from_csr(A), {}), (spmv.flat_partitioned, A, {"K": 8}),
                      (spmv.flat_partitioned, A, {"K": 128}), (spmv.original, A, {})):
        y = torch.full((3000,), float("nan"), device="cuda")
        fn(M, xd, y, **kw)
        np.testing.assert_array_equal(y.cpu().numpy(), want, err_msg=fn.__name__)
    # a banded matrix (7 diagonals, some entries missing) in DIA, general floats
    n = 5000
    rng = np.random.default_rng(4)
    rows_l, cols_l = [], []
    for d in (-40, -3, -1, 0, 1, 2, 57):
        r = np.arange(max(0, -d), min(n, n - d))
        keep = rng.random(len(r)) < 0.9
        rows_l.append(r[keep]); cols_l.append(r[keep] + d)
    r, c = np.concatenate(rows_l), np.concatenate(cols_l)
    order = np.lexsort((c, r))
    r, c = r[order], c[order]
    off = np.zeros(n + 1, np.int32); np.add.at(off, r + 1, 1); off = np.cumsum(off).astype(np.int32)
    idx = c.astype(np.int32)
    val = rng.uniform(0.5, 1.5, len(idx)).astype(np.float32)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    A = csr_t(n, n, off, idx, val)
    dia = dia_t.from_csr(A)
    assert dia.num_diagonals == 7
    y = torch.full((n,), float("nan"), device="cuda")
    spmv.dia_thread_mapped(dia, torch.as_tensor(x).cuda(), y)
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv(off, idx, val, x))   # same order, un-fused


@pytest.mark.parametrize("n", [1, 5, 32, 70])
def test_spmm_thread_mapped(oracle, battery, n):
    """algorithms/spmm/thread_mapped.cuh: C = A B; per (row, col) the same sequential
    un-fused sum as the restatement -> bit-exact on any input."""
    from loops_b200 import csr_t
    from loops_b200.algorithms import spmm
    rng = np.random.default_rng(n)
    for b in battery[:6]:
        A = csr_t(b["rows"], b["cols"], b["off"], b["idx"], b["val"])
        B = rng.uniform(-1, 1, (b["cols"], n)).astype(np.float32)
        Cm = torch.full((b["rows"], n), float("nan"), device="cuda")
        spmm.thread_mapped(A, torch.as_tensor(B).cuda(), Cm)
        np.testing.assert_array_equal(Cm.cpu().numpy(), oracle.spmm(b["off"], b["idx"], b["val"], B), err_msg=b["name"])
    # column 0 of an SpMM with B[:, 0] = x is the SpMV
    b = battery[0]
    A = csr_t(b["rows"], b["cols"], b["off"], b["idx"], b["val"])
    B = np.zeros((b["cols"], n), np.float32); B[:, 0] = b["x"]
    Cm = torch.empty((b["rows"], n), device="cuda")
    spmm.thread_mapped(A, torch.as_tensor(B).cuda(), Cm)
    np.testing.assert_array_equal(Cm[:, 0].cpu().numpy(), b["y"])


@pytest.mark.parametrize("R", [2, 3, 4])
def test_battery_bcsr_f32(oracle, battery, R):
    """unittests/test_spmv_bcsr.cu:24-42 (2x2, 3x3) + 4x4, padded x."""
    from loops_b200 import csr_t, bcsr_t
    from loops_b200.algorithms import spmv
    for b in battery:
        A = csr_t(b["rows"], b["cols"], b["off"], b["idx"], b["val"])
        B = bcsr_t.from_csr(A, R, R)
        g_off, g_col, g_val = b[f"bcsr{R}"]
        np.testing.assert_array_equal(B.block_offsets.cpu().numpy(), g_off)
        np.testing.assert_array_equal(B.block_col_indices.cpu().numpy(), g_col)
        np.testing.assert_array_equal(B.values.cpu().numpy(), g_val)
        xp = B.padded_x(torch.as_tensor(b["x"]).cuda())
        y = torch.full((b["rows"],), float("nan"), device="cuda")
        spmv.bcsr_thread_mapped(B, xp, y)
        ref = oracle.spmv_bcsr(R, R, b["rows"], g_off, g_col, g_val, xp.cpu().numpy())
        np.testing.assert_array_equal(y.cpu().numpy(), ref)          # same order, un-fused


EDGE = {
    "single_row": (1, 300, 0.5, 1, 0, None),
    "single_col": (257, 1, 1.0, 2, 0, None),
    "all_empty_but_one": (500, 64, 0.0, 3, 0, (250, 64)),
    "empty_rows_every_2": (1000, 200, 0.05, 4, 2, None),
    "one_huge_row": (300, 20000, 0.0005, 5, 0, (150, 20000)),     # spans several CTA tiles
    "rows_gt_tile": (9000, 64, 0.01, 6, 3, None),                 # > 4096 row ends in a tile
    "dense_small": (64, 64, 1.0, 7, 0, None),
    "tall_thin": (20000, 8, 0.3, 8, 0, None),
}


@pytest.mark.parametrize("kernel", CSR_KERNELS)
@pytest.mark.parametrize("case", sorted(EDGE))
def test_edge_cases(oracle, kernel, case):
    rows, cols, dens, seed, empty, heavy = EDGE[case]
    off, idx, val = random_csr(rows, cols, dens, seed, empty, heavy)
    x = np.random.default_rng(seed).uniform(-1, 1, cols).astype(np.float32)
    y = _run_csr(kernel, off, idx, val, x, rows, cols)
    _assert_close(oracle, off, idx, val, x, y, (kernel, case))


@pytest.mark.parametrize("kernel", CSR_KERNELS)
def test_degenerate_shapes(kernel):
    """0 x 0, rows without any stored entry: not exercised by the reference's
    battery but must not crash; y = 0."""
    y = _run_csr(kernel, np.zeros(1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32),
                 np.zeros(0, np.float32), 0, 0)
    assert y.size == 0
    y = _run_csr(kernel, np.zeros(6, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32),
                 np.ones(5, np.float32), 5, 5)
    np.testing.assert_array_equal(y, np.zeros(5, np.float32))


@pytest.mark.parametrize("kernel", CSR_KERNELS)
def test_exact_inputs_bit_identical(oracle, kernel):
    """Exactly representable inputs (values k/8, integer x): every summation
    order gives the same bits, so y must EQUAL the reference's."""
    from loops_b200 import generate as g
    rows = cols = 1 << 14
    off, idx, val = g.synth_csr(rows, cols, rows * 24)
    x = g.x_recipe(cols)
    off, idx, val, x = off.numpy(), idx.numpy(), val.numpy(), x.numpy()
    y = _run_csr(kernel, off, idx, val, x, rows, cols)
    np.testing.assert_array_equal(y, oracle.spmv(off, idx, val, x))


def test_y_is_overwritten_and_deterministic(oracle):
    """No pre-zero precondition (the reference's atomic kernels need one) and,
    for merge_path_flat, run-to-run bit-stable results on general floats."""
    off, idx, val = random_csr(3000, 3000, 0.01, 11, 7, (5, 3000))
    x = np.random.default_rng(1).normal(size=3000).astype(np.float32)
    ys = [_run_csr("merge_path_flat", off, idx, val, x, 3000, 3000, poison=p) for p in (True, False, True)]
    np.testing.assert_array_equal(ys[0], ys[1])
    np.testing.assert_array_equal(ys[0], ys[2])
    for k in CSR_KERNELS:
        y = _run_csr(k, off, idx, val, x, 3000, 3000, poison=True)
        _assert_close(oracle, off, idx, val, x, y, k)


def test_host_buffer_entry_point(oracle, battery):
    """loopsb_spmv_csr_host_f32: the reference example's whole flow."""
    import ctypes as C
    from loops_b200 import _lib
    lib = _lib.load()
    b = battery[8]
    for sched in range(4):
        y = np.full(b["rows"], np.nan, np.float32)
        ms = C.c_float()
        rc = lib.loopsb_spmv_csr_host_f32(sched, b["rows"], b["cols"], len(b["idx"]), b["off"].ctypes.data,
                                          b["idx"].ctypes.data, b["val"].ctypes.data, b["x"].ctypes.data,
                                          y.ctypes.data, C.byref(ms))
        _lib.check(rc, "loopsb_spmv_csr_host_f32")
        _assert_close(oracle, b["off"], b["idx"], b["val"], b["x"], y, ("host", sched))
        assert ms.value > 0


def test_unsupported_cells_fail_loudly():
    from loops_b200 import _lib, csr_t
    from loops_b200.container import csc_t
    off, idx, val = random_csr(10, 10, 0.3, 1)
    csc = csc_t.from_csr(csr_t(10, 10, off, idx, val))
    with pytest.raises(_lib.LoopsbError) as e:
        csc.plan(_lib.SCHED_GROUP_MAPPED)          # csc has only its thread_mapped entry point
    assert e.value.status == _lib.ERR_UNSUPPORTED


def test_all_twelve_config3_cells(oracle, battery):
    """BASELINE configs[2]: {thread_mapped, group_mapped, work_oriented, merge_path_flat} x
    {csr, coo, ell}. Seven cells have a reference kernel; the other five (coo x group / work /
    merge, ell x group / work) are schedule::setup<> over that layout with the format's
    per-atom body. Every cell: exact on exact inputs, and within the north-star
    tolerance on the battery's floats; COO cells also on row-unsorted triples."""
    from loops_b200 import csr_t, coo_t, ell_t
    from loops_b200.algorithms import spmv
    assert len(spmv.CELLS) == 12
    off, idx, val = random_csr(900, 700, 0.02, seed=17, exact=True, empty_every=5, heavy_row=(400, 650))
    x = np.random.default_rng(6).integers(1, 11, 700).astype(np.float32)
    ref = oracle.spmv(off, idx, val, x)
    A = csr_t(900, 700, off, idx, val)
    perm = np.random.default_rng(1).permutation(len(idx))
    rows_of = np.repeat(np.arange(900, dtype=np.int32), np.diff(off))
    containers = {"csr": [A], "coo": [coo_t.from_csr(A), coo_t(900, 700, rows_of[perm], idx[perm], val[perm])],
                  "ell": [ell_t.from_csr(A)]}
    xd = torch.as_tensor(x).cuda()
    for (layout, sched), fn in spmv.CELLS.items():
        for M in containers[layout]:
            y = torch.full((900,), float("nan"), device="cuda")
            fn(M, xd, y)
            np.testing.assert_array_equal(y.cpu().numpy(), ref, err_msg=f"{layout} x {sched}")
    for b in battery:
        A = csr_t(b["rows"], b["cols"], b["off"], b["idx"], b["val"])
        cs = {"csr": A, "coo": coo_t.from_csr(A), "ell": ell_t.from_csr(A)}
        xb = torch.as_tensor(b["x"]).cuda()
        for (layout, sched), fn in spmv.CELLS.items():
            y = torch.full((b["rows"],), float("nan"), device="cuda")
            fn(cs[layout], xb, y)
            _assert_close(oracle, b["off"], b["idx"], b["val"], b["x"], y.cpu().numpy(), (layout, sched, b["name"]))


def test_full_size_config2_properties(oracle):
    """BASELINE config 2 (2^20 rows, 2^25 nnz) at full size, through
    size-independent properties: (i) exact inputs -> all four schedules agree
    bit for bit with each other and with a checksum computed independently in
    float64; (ii) linearity A(2x) == 2 A(x) exactly; (iii) row-sum identity
    A * 1 == per-row sum of values."""
    from loops_b200 import csr_t, generate as g
    from loops_b200.algorithms import spmv
    rows = cols = 1 << 20
    nnz = 1 << 25
    off, idx, val = g.synth_csr(rows, cols, nnz, device="cuda")
    A = csr_t.from_tensors(rows, cols, off, idx, val)
    x = g.x_recipe(cols, device="cuda")
    ys = {}
    for k in CSR_KERNELS:
        y = torch.full((rows,), float("nan"), device="cuda")
        spmv.BY_NAME[k](A, x, y)
        ys[k] = y
    for k in CSR_KERNELS[1:]:
        assert torch.equal(ys[k], ys["merge_path_flat"]), k
    # independent float64 checksum with torch ops (exact: all terms are k/8 * int)
    prod = val.double() * x[idx.long()].double()
    rowid = torch.repeat_interleave(torch.arange(rows, device="cuda"), (off[1:] - off[:-1]).long())
    y64 = torch.zeros(rows, dtype=torch.float64, device="cuda").index_add_(0, rowid, prod)
    assert torch.equal(ys["merge_path_flat"].double(), y64)
    y2 = torch.empty(rows, device="cuda")
    spmv.merge_path_flat(A, 2 * x, y2)
    assert torch.equal(y2, 2 * ys["merge_path_flat"])
    ones = torch.ones(cols, device="cuda")
    y1 = torch.empty(rows, device="cuda")
    spmv.merge_path_flat(A, ones, y1)
    rs = torch.zeros(rows, dtype=torch.float64, device="cuda").index_add_(0, rowid, val.double())
    assert torch.equal(y1.double(), rs)
    # a 64K-row prefix against the CPU oracle itself
    n = 1 << 16
    o = off[: n + 1].cpu().numpy()
    yo = oracle.spmv(o, idx[: o[-1]].cpu().numpy(), val[: o[-1]].cpu().numpy(), x.cpu().numpy())
    np.testing.assert_array_equal(ys["merge_path_flat"][:n].cpu().numpy(), yo)


def test_automatic_picks_a_schedule_and_matches_the_oracle(oracle):
    """spmv.automatic: tiny matrices and light regular rows go to
    thread_mapped, everything else to merge_path_flat; y equals the oracle either way."""
    from loops_b200 import csr_t
    from loops_b200.algorithms import spmv
    c = load_chesapeake()
    A = csr_t(39, 39, c["off"], c["idx"], c["val"])
    assert spmv.select_schedule(A) == "thread_mapped"
    y = torch.full((39,), float("nan"), device="cuda")
    spmv.automatic(A, torch.as_tensor(c["x"]).cuda(), y)
    np.testing.assert_array_equal(y.cpu().numpy(), c["y"])
    rows, cols = 3000, 2500
    off, idx, val = random_csr(rows, cols, 0.01, 5, empty_every=9, heavy_row=(3, 2000), exact=True)
    B = csr_t(rows, cols, off, idx, val)
    assert spmv.select_schedule(B) == "merge_path_flat"       # 75 K nonzeros, a 2000-wide row
    x = oracle.x_recipe_int(cols)
    y = torch.full((rows,), float("nan"), device="cuda")
    spmv.automatic(B, torch.as_tensor(x).cuda(), y)
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv(off, idx, val, x))


def test_coo_unsorted_rows(oracle, tmp_path):
    """COO triples are not row-sorted in general -- the loader
    returns file order and mirrors symmetric entries in place. Rows [5,3,5]: the two
    5s are separate runs; the reference adds every nonzero atomically and accepts
    any order (coo_thread_mapped.cuh:37-51)."""
    from loops_b200 import coo_t
    from loops_b200 import market
    from loops_b200.algorithms import spmv
    # the advisor's minimal case
    r = np.array([5, 3, 5], np.int32); c = np.array([0, 1, 2], np.int32)
    v = np.array([1.0, 10.0, 100.0], np.float32)
    y = torch.full((8,), float("nan"), device="cuda")
    spmv.coo_thread_mapped(coo_t(8, 3, r, c, v), torch.ones(3, device="cuda"), y)
    np.testing.assert_array_equal(y.cpu().numpy(), np.array([0, 0, 0, 10, 0, 101, 0, 0], np.float32))
    # a permuted COO of a random matrix, exact inputs
    off, idx, val = random_csr(700, 500, 0.03, seed=31, exact=True)
    x = np.random.default_rng(2).integers(1, 11, 500).astype(np.float32)
    rows_of = np.repeat(np.arange(700, dtype=np.int32), np.diff(off))
    for seed in (0, 1):
        perm = np.random.default_rng(seed).permutation(len(idx))
        if seed == 1:      # short runs of equal rows separated by other rows
            perm = np.argsort(rows_of % 7, kind="stable")
        y = torch.full((700,), float("nan"), device="cuda")
        spmv.coo_thread_mapped(coo_t(700, 500, rows_of[perm], idx[perm], val[perm]), torch.as_tensor(x).cuda(), y)
        np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv(off, idx, val, x))
    # a symmetric .mtx loaded straight to COO (mirrors sit right after their sources)
    p = tmp_path / "sym.mtx"
    ent = [(i + 1, j + 1, (i * 7 + j) % 16 / 8 + 0.125) for i in range(60) for j in range(0, i + 1, 3)]
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n" + "60 60 %d\n" % len(ent) +
                 "".join("%d %d %.3f\n" % e for e in ent))
    rows, cols, rr, cc, vv = market.load_coo(str(p))
    xs = np.random.default_rng(5).integers(1, 11, cols).astype(np.float32)
    y = torch.full((rows,), float("nan"), device="cuda")
    spmv.coo_thread_mapped(coo_t(rows, cols, rr, cc, vv), torch.as_tensor(xs).cuda(), y)
    o2, i2, v2 = market.coo_to_csr(rows, cols, rr, cc, vv)
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv(o2, i2, v2, xs))
