"""CPU: the C restatement against the unmodified reference headers compiled
host-only (oracle/_ref/libloopsref_host.so) on seeded random inputs."""
import ctypes as C

import numpy as np
import pytest

from helpers import P, random_csr

CASES = [(64, 64, 0.1, 1, 0, None), (200, 150, 0.05, 2, 5, None), (97, 300, 0.02, 3, 0, (7, 250)),
         (1, 40, 0.5, 4, 0, None), (300, 17, 0.3, 5, 3, None)]


@pytest.mark.parametrize("rows,cols,dens,seed,empty,heavy", CASES)
def test_validator_functions(oracle, ref_host, rows, cols, dens, seed, empty, heavy):
    off, idx, val = random_csr(rows, cols, dens, seed, empty, heavy)
    nnz = len(idx)
    x = np.random.default_rng(seed + 100).uniform(-2, 2, cols).astype(np.float32)
    for name, fn in (("ref_spmv_f32", oracle.spmv), ("ref_spmv_f64", oracle.spmv_f64), ("ref_row_l1", oracle.row_l1)):
        y = np.zeros(rows, np.float32)
        getattr(ref_host, name)(rows, cols, nnz, P(off), P(idx), P(val), P(x), P(y))
        np.testing.assert_array_equal(fn(off, idx, val, x), y, err_msg=name)


@pytest.mark.parametrize("rows,cols,dens,seed,empty,heavy", CASES)
def test_validator_in_double(oracle, ref_host, rows, cols, dens, seed, empty, heavy):
    """reference::spmv<double> (what the reference's .f64 example builds validate against)."""
    off, idx, val = random_csr(rows, cols, dens, seed, empty, heavy)
    rng = np.random.default_rng(seed + 200)
    val64 = rng.uniform(-1.5, 1.5, len(idx))
    x64 = rng.uniform(-2, 2, cols)
    y = np.zeros(rows, np.float64)
    ref_host.ref_spmv_d(rows, cols, len(idx), P(off), P(idx), P(val64), P(x64), P(y))
    np.testing.assert_array_equal(oracle.spmv_d(off, idx, val64, x64), y)


@pytest.mark.parametrize("rows,cols,dens,seed,empty,heavy", CASES)
def test_conversions(oracle, ref_host, rows, cols, dens, seed, empty, heavy):
    off, idx, val = random_csr(rows, cols, dens, seed, empty, heavy)
    nnz = len(idx)
    pitch, e_idx, e_val = oracle.ell(off, idx, val)
    assert pitch == ref_host.ref_ell_pitch(rows, cols, nnz, P(off), P(idx), P(val))
    r_idx, r_val = np.zeros_like(e_idx), np.zeros_like(e_val)
    ref_host.ref_csr_to_ell(rows, cols, nnz, P(off), P(idx), P(val), P(r_idx), P(r_val))
    np.testing.assert_array_equal(e_idx, r_idx)
    np.testing.assert_array_equal(e_val, r_val)
    rr = np.zeros(nnz, np.int32)
    ref_host.ref_csr_to_coo_rows(rows, cols, nnz, P(off), P(idx), P(val), P(rr))
    np.testing.assert_array_equal(oracle.coo_rows(off), rr)
    for R in (2, 3, 4):
        b_off, b_col, b_val = oracle.bcsr(R, R, rows, cols, off, idx, val)
        nb = C.c_int()
        g_off, g_col, g_val = np.zeros_like(b_off), np.zeros_like(b_col), np.zeros_like(b_val)
        rc = ref_host.ref_csr_to_bcsr(R, rows, cols, nnz, P(off), P(idx), P(val), C.byref(nb),
                                      P(g_off), P(g_col), P(g_val), len(b_col))
        assert rc == 0 and nb.value == len(b_col)
        np.testing.assert_array_equal(b_off, g_off)
        np.testing.assert_array_equal(b_col, g_col)
        np.testing.assert_array_equal(b_val, g_val)


@pytest.mark.parametrize("rows,cols,dens,seed,empty,heavy", CASES)
def test_csc_dia_conversions_and_their_spmv(oracle, ref_host, rows, cols, dens, seed, empty, heavy):
    """csc_t(csr) / dia_t(csr) (container/csc.hxx:88-102, dia.hxx:135-188) and the
    sequential restatements of the csc / dia / flat_partitioned SpMV kernels: on
    exact inputs every order of the adds gives the validator's y."""
    off, idx, val = random_csr(rows, cols, dens, seed, empty, heavy, exact=True)
    nnz = len(idx)
    c_off, c_row, c_val = oracle.csc(rows, cols, off, idx, val)
    g_off, g_row, g_val = np.zeros_like(c_off), np.zeros_like(c_row), np.zeros_like(c_val)
    ref_host.ref_csr_to_csc(rows, cols, nnz, P(off), P(idx), P(val), P(g_off), P(g_row), P(g_val))
    np.testing.assert_array_equal(c_off, g_off)
    np.testing.assert_array_equal(c_row, g_row)
    np.testing.assert_array_equal(c_val, g_val)
    d_off, d_val = oracle.dia(rows, off, idx, val)
    nd = ref_host.ref_csr_to_dia(rows, cols, nnz, P(off), P(idx), P(val), None, None, 0)
    assert nd == len(d_off)
    r_off, r_val = np.zeros_like(d_off), np.zeros_like(d_val)
    assert ref_host.ref_csr_to_dia(rows, cols, nnz, P(off), P(idx), P(val), P(r_off), P(r_val), nd) == nd
    np.testing.assert_array_equal(d_off, r_off)
    np.testing.assert_array_equal(d_val, r_val)
    x = oracle.x_recipe_int(cols)
    want = oracle.spmv(off, idx, val, x)
    np.testing.assert_array_equal(oracle.spmv_csc(rows, c_off, c_row, c_val, x), want)
    np.testing.assert_array_equal(oracle.spmv_dia(rows, cols, d_off, d_val, x), want)
    for K in (1, 8, 13):
        np.testing.assert_array_equal(oracle.spmv_flat_partitioned(K, off, idx, val, x), want)


def test_x_recipe(oracle, ref_host):
    for seed in (42, 1, 123456789):
        for (lo, hi) in ((1, 10), (0, 1), (-5, 5)):
            r = np.zeros(5000, np.float32)
            ref_host.ref_x_recipe_int(5000, lo, hi, seed, P(r))
            np.testing.assert_array_equal(oracle.x_recipe_int(5000, lo, hi, seed), r)
        r = np.zeros(5000, np.float32)
        ref_host.ref_x_recipe_float(5000, C.c_float(1.0), C.c_float(10.0), seed, P(r))
        np.testing.assert_array_equal(oracle.x_recipe_float(5000, 1.0, 10.0, seed), r)


def test_tolerance_predicate(oracle, ref_host):
    rng = np.random.default_rng(0)
    a = rng.normal(size=2000).astype(np.float32) * 10
    b = a + rng.normal(size=2000).astype(np.float32) * 0.02
    for u, v in zip(a, b):
        assert oracle.L.orc_tolerance_ne(C.c_float(u), C.c_float(v)) == \
            ref_host.ref_default_tolerance_ne(C.c_float(u), C.c_float(v))
