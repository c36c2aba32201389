"""CPU: the C restatement (oracle/) against the reference's own fixtures and
known answers. These pin the oracle before anything trusts it."""
import numpy as np
import pytest

from helpers import Oracle, SCHED, load_chesapeake, ORC_FLAT


def test_chesapeake_known_answers(oracle):
    """SURVEY 8c / site/content/experimentation.md:19-37: 39 x 39 (340),
    offsets 0 11 22 29 33 37, x 1 10 6 2 10, y 50 52 53 26 18, sum 1794."""
    c = load_chesapeake()
    assert tuple(c["dims"]) == (39, 39) and len(c["idx"]) == 340
    assert list(c["off"][:6]) == [0, 11, 22, 29, 33, 37]
    x = oracle.x_recipe_int(39, 1, 10, 42)
    assert list(x[:5]) == [1, 10, 6, 2, 10]
    np.testing.assert_array_equal(x, c["x"])
    y = oracle.spmv(c["off"], c["idx"], c["val"], x)
    assert list(y[:5]) == [50, 52, 53, 26, 18] and float(y.sum()) == 1794.0
    np.testing.assert_array_equal(y, c["y"])
    np.testing.assert_array_equal(oracle.spmv_f64(c["off"], c["idx"], c["val"], x), c["y64"])
    np.testing.assert_array_equal(oracle.row_l1(c["off"], c["idx"], c["val"], x), c["l1"])
    assert oracle.count_errors(y, c["y"]) == 0


def test_x_recipe_golden(oracle):
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "xrecipe.npz"))
    for seed in (42, 7):
        np.testing.assert_array_equal(oracle.x_recipe_int(4096, 1, 10, seed), z[f"int_1_10_seed{seed}"])
        np.testing.assert_array_equal(oracle.x_recipe_float(4096, 0.0, 1.0, seed), z[f"float_0_1_seed{seed}"])
    h = np.array([oracle.L.orc_hash(i) for i in range(1024)], np.uint32)
    np.testing.assert_array_equal(h, z["hash_0_1023"])


def test_battery_spmv_bit_exact(oracle, battery):
    """reference_spmv of unittests/test_helpers.hxx:268-278 on its own battery."""
    assert len(battery) == 9
    for b in battery:
        y = oracle.spmv(b["off"], b["idx"], b["val"], b["x"])
        np.testing.assert_array_equal(y, b["y"], err_msg=b["name"])


def test_battery_conversions(oracle, battery):
    """CSR->COO/ELL/BCSR restatements equal the reference's converters
    (coo.hxx:87-98, ell.hxx:113-145, bcsr.hxx:111-194)."""
    for b in battery:
        np.testing.assert_array_equal(oracle.coo_rows(b["off"]), b["coo_rows"])
        pitch, e_idx, e_val = oracle.ell(b["off"], b["idx"], b["val"])
        assert pitch == b["ell_pitch"]
        np.testing.assert_array_equal(e_idx, b["ell_idx"])
        np.testing.assert_array_equal(e_val, b["ell_val"])
        for R in (2, 3, 4):
            b_off, b_col, b_val = oracle.bcsr(R, R, b["rows"], b["cols"], b["off"], b["idx"], b["val"])
            g_off, g_col, g_val = b[f"bcsr{R}"]
            np.testing.assert_array_equal(b_off, g_off)
            np.testing.assert_array_equal(b_col, g_col)
            np.testing.assert_array_equal(b_val, g_val)


def test_battery_format_spmv_agree(oracle, battery):
    """Every format's sequential SpMV reproduces the CSR answer on the battery
    (dense-equivalence idea of unittests/test_format_round_trip.cu:131-168)."""
    for b in battery:
        y_ell = oracle.spmv_ell(b["rows"], b["ell_pitch"], b["ell_idx"], b["ell_val"], b["x"])
        np.testing.assert_array_equal(y_ell, b["y"])
        y_coo = oracle.spmv_coo(b["rows"], b["coo_rows"], b["idx"], b["val"], b["x"])
        np.testing.assert_array_equal(y_coo, b["y"])
        for R in (2, 3, 4):
            b_off, b_col, b_val = b[f"bcsr{R}"]
            nbc = (b["cols"] + R - 1) // R
            xp = np.zeros(nbc * R, np.float32)
            xp[: b["cols"]] = b["x"]
            y_b = oracle.spmv_bcsr(R, R, b["rows"], b_off, b_col, b_val, xp)
            np.testing.assert_allclose(y_b, b["y"], rtol=1e-6, atol=1e-6)


def test_layout_csr_fixture(oracle):
    """unittests/test_layout_csr.cu:26-51: offsets {0,2,2,5,7}."""
    off = np.array([0, 2, 2, 5, 7], np.int32)
    lay = Oracle.csr_layout(off)
    L = oracle.L
    import ctypes as C
    assert L.orc_num_tiles(C.byref(lay)) == 4 and L.orc_num_atoms(C.byref(lay)) == 7
    assert [L.orc_tile_size(C.byref(lay), t) for t in range(4)] == [2, 0, 3, 2]
    assert [L.orc_tile_of(C.byref(lay), a) for a in (0, 1, 2, 4, 5, 6)] == [0, 0, 2, 2, 3, 3]
    assert L.orc_tile_begin(C.byref(lay), 0) == 0 and L.orc_tile_end(C.byref(lay), 3) == 7


def test_layout_arithmetic_kinds(oracle):
    """coo (test_layout_coo.cu:19-58), ell (test_layout_ell.cu:19-67) and the
    partitioner (test_layout_flat_partitioner.cu:27-59) contracts."""
    import ctypes as C
    L = oracle.L
    coo = Oracle.coo_layout(9)
    assert L.orc_num_tiles(C.byref(coo)) == 9 and L.orc_tile_end(C.byref(coo), 4) == 5
    assert L.orc_tile_of(C.byref(coo), 7) == 7 and L.orc_tile_size(C.byref(coo), 3) == 1
    ell = Oracle.ell_layout(5, 3)
    assert L.orc_num_atoms(C.byref(ell)) == 15 and L.orc_tile_begin(C.byref(ell), 2) == 6
    assert L.orc_tile_end(C.byref(ell), 4) == 15 and L.orc_tile_of(C.byref(ell), 8) == 2
    flat = Oracle.layout(ORC_FLAT, None, 0, 10, 4)     # 10 atoms in windows of 4
    assert L.orc_num_tiles(C.byref(flat)) == 3
    assert [L.orc_tile_end(C.byref(flat), t) for t in range(3)] == [4, 8, 10]
    assert L.orc_tile_size(C.byref(flat), 2) == 2 and L.orc_tile_of(C.byref(flat), 9) == 2


def _check_invariants(oracle, lay):
    """unittests/test_layout_contract.hxx:30-88."""
    import ctypes as C
    L = oracle.L
    T, A = L.orc_num_tiles(C.byref(lay)), L.orc_num_atoms(C.byref(lay))
    if T == 0:
        return
    assert L.orc_tile_begin(C.byref(lay), 0) == 0
    assert L.orc_tile_end(C.byref(lay), T - 1) == A
    prev = 0
    for t in range(T):
        e = L.orc_tile_end(C.byref(lay), t)
        assert e >= prev and L.orc_tile_size(C.byref(lay), t) == e - L.orc_tile_begin(C.byref(lay), t)
        prev = e
    for a in range(A):
        t = L.orc_tile_of(C.byref(lay), a)
        assert L.orc_tile_begin(C.byref(lay), t) <= a < L.orc_tile_end(C.byref(lay), t)


def test_layout_invariants_battery(oracle, battery):
    for b in battery:
        _check_invariants(oracle, Oracle.csr_layout(b["off"]))
        _check_invariants(oracle, Oracle.coo_layout(len(b["idx"])))
        _check_invariants(oracle, Oracle.ell_layout(b["rows"], b["ell_pitch"]))


@pytest.mark.parametrize("sched,grid,tpb,ipt", [
    ("thread_mapped", 2, 32, 1), ("thread_mapped", 16, 256, 1),   # test_schedule_coverage.cu:58-111
    ("group_mapped", 0, 128, 1), ("work_oriented", 1, 128, 1), ("work_oriented", 3, 128, 1),
    ("merge_path_flat", 0, 128, 8), ("merge_path_flat", 0, 128, 7), ("merge_path_flat", 0, 128, 5)])
def test_every_atom_visited_exactly_once(oracle, battery, sched, grid, tpb, ipt):
    """unittests/test_schedule_coverage.cu:54-112 generalised to all four
    schedules and three layout kinds, plus ownership of every atom's tile."""
    for b in battery:
        lays = [Oracle.csr_layout(b["off"]), Oracle.coo_layout(len(b["idx"])),
                Oracle.ell_layout(b["rows"], b["ell_pitch"])]
        for lay in lays:
            s = oracle.emit(lay, SCHED[sched], grid, tpb, ipt)
            A = oracle.num_atoms(lay)
            assert np.all(s["visits"][:A] == 1), (b["name"], sched)
            import ctypes as C
            owners = np.array([oracle.L.orc_tile_of(C.byref(lay), a) for a in range(A)], np.int32)
            np.testing.assert_array_equal(s["tile"][:A], owners)


def test_schedule_coverage_fixture(oracle):
    """test_schedule_coverage.cu:58-85: 4 tiles / 9 atoms, grid 2 x 32."""
    off = np.array([0, 3, 3, 7, 9], np.int32)
    s = oracle.emit(Oracle.csr_layout(off), SCHED["thread_mapped"], 2, 32, 1)
    assert list(s["visits"]) == [1] * 9
    assert list(s["visitor"]) == [0, 0, 0, 2, 2, 2, 2, 3, 3]
    assert list(s["step"]) == [0, 1, 2, 0, 1, 2, 3, 0, 1]


def test_merge_path_hand_example(oracle):
    """Hand-checked merge path: offsets {0,2,2,5,7}, W = 11 -> one merge tile;
    path alternates atoms and row ends: a0 a1 | E0 E1 | a2 a3 a4 | E2 | a5 a6 | E3."""
    off = np.array([0, 2, 2, 5, 7], np.int32)
    s = oracle.emit(Oracle.csr_layout(off), SCHED["merge_path_flat"], 0, 128, 8)
    assert s["coords"].tolist() == [[0, 0], [4, 7]]
    # thread 0 owns path items 0..7: a0 a1 E0 E1 a2 a3 a4 E2
    assert s["dense_emit"][:8].tolist() == [1, 1, 0, 0, 1, 1, 1, 0]
    assert s["dense_tile"][:8].tolist() == [0, 0, 0, 1, 2, 2, 2, 2]
    assert s["dense_atom"][:8].tolist() == [0, 1, 2, 2, 2, 3, 4, 5]
    # thread 1 starts at diagonal 8 = (3, 5): a5 a6 E3, then runs off the end
    assert s["thread_start"][2:4].tolist() == [3, 5]
    assert s["dense_emit"][8:11].tolist() == [1, 1, 0]
    assert s["visitor"].tolist() == [0, 0, 0, 0, 0, 1, 1]


def test_golden_streams_from_reference_gpu(oracle, battery):
    """Index streams recorded from the reference's own schedule::setup<>
    templates on a B200 (tests/golden/streams.npz, made by make_golden.py gpu)."""
    import os
    from helpers import GOLDEN
    path = os.path.join(GOLDEN, "streams.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/streams.npz not captured yet")
    z = np.load(path)
    keys = sorted({k.rsplit("|", 1)[0] for k in z.files if "|" in k})
    assert keys
    for key in keys:
        i, lname, sname, grid, tpb, ipt = key.split("|")
        b = battery[int(i)]
        lay = {"csr": Oracle.csr_layout(b["off"]), "coo": Oracle.coo_layout(len(b["idx"])),
               "ell": Oracle.ell_layout(b["rows"], b["ell_pitch"])}[lname]
        s = oracle.emit(lay, SCHED[sname], int(grid), int(tpb), int(ipt))
        for field in ("visitor", "step", "tile", "visits", "commit", "map", "dense_tile", "dense_atom",
                      "dense_emit", "thread_start"):
            gk = f"{key}|{field}"
            if gk in z.files:
                np.testing.assert_array_equal(s[field], z[gk], err_msg=gk)
