"""CPU: the Matrix Market loader -- the Python mirror
(loops_b200/market.py) and the C++ header (include/loops/container/market.hxx,
through tests/cpp/market_load) against
  * the reference's own loader run on the same files (oracle/_ref, when the
    reference is mounted: container/market.hxx:100-289 + csr_t(coo)),
  * the chesapeake known answers of BASELINE config 1 (39 x 39, 340 nnz,
    offsets 0 11 22 29 33 37 ..., y[0..4] = 50 52 53 26 18, sum 1794),
  * the rejection rules of unittests/test_market_loader.cu:95-295."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import P, load_chesapeake

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "market_load")


def write_mtx(path, rows, cols, entries, field="real", symmetry="general", comments=True):
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} {symmetry}\n")
        if comments:
            f.write("% a comment line\n%\n\n")
        f.write(f"  {rows} {cols} {len(entries)}\n")
        for k, e in enumerate(entries):
            if field == "pattern":
                f.write(f"{e[0]} {e[1]}\n")
            elif field == "integer":
                f.write(f"{e[0]}\t{e[1]} {int(e[2])}\n")
            else:
                f.write(f"{e[0]} {e[1]}   {e[2]!r}\n" if k % 2 else f" {e[0]} {e[1]} {e[2]:.9e}\r\n")


def cpp_load(path):
    p = subprocess.run([EXE, path], capture_output=True, text=True, timeout=60)
    if p.returncode != 0:
        return p.returncode, p.stderr
    lines = p.stdout.split("\n")
    rows, cols, nnz = (int(t) for t in lines[0].split())
    coo = np.array([[float(t) for t in ln.split()] for ln in lines[1:1 + nnz]], dtype=np.float64).reshape(nnz, 3)
    off = np.array(lines[1 + nnz].split(), dtype=np.int32)
    idx = np.array(lines[2 + nnz].split(), dtype=np.int32)
    val = np.array(lines[3 + nnz].split(), dtype=np.float32)
    return 0, (rows, cols, coo, off, idx, val)


def random_entries(rng, rows, cols, n, symmetric):
    seen, out = set(), []
    while len(out) < n:
        r, c = int(rng.integers(1, rows + 1)), int(rng.integers(1, cols + 1))
        if symmetric and c > r:
            r, c = c, r
        if (r, c) in seen:
            continue
        seen.add((r, c))
        out.append((r, c, float(np.float32(rng.uniform(-3, 3)))))
    return out


CASES = [("real", "general"), ("real", "symmetric"), ("integer", "general"), ("pattern", "symmetric"),
         ("pattern", "general"), ("integer", "symmetric")]


@pytest.mark.parametrize("field,symmetry", CASES)
def test_loader_matches_reference(tmp_path, field, symmetry):
    from loops_b200 import market
    rng = np.random.default_rng(hash((field, symmetry)) % 1000)
    sym = symmetry == "symmetric"
    rows = cols = 57 if sym else 0
    if not sym:
        rows, cols = 41, 73
    entries = random_entries(rng, rows, cols, 300, sym)
    if field == "integer":
        entries = [(r, c, float(int(v * 3))) for r, c, v in entries]
    path = str(tmp_path / "m.mtx")
    write_mtx(path, rows, cols, entries, field, symmetry)
    R, Cc, r, c, v = market.load_coo(path)
    off, idx, val = market.coo_to_csr(R, Cc, r, c, v)
    assert (R, Cc) == (rows, cols)
    # C++ header: same COO order, same CSR
    if os.path.exists(EXE):
        rc, out = cpp_load(path)
        assert rc == 0, out
        assert out[0] == rows and out[1] == cols
        np.testing.assert_array_equal(out[2][:, 0].astype(np.int32), r)
        np.testing.assert_array_equal(out[2][:, 1].astype(np.int32), c)
        np.testing.assert_array_equal(out[2][:, 2].astype(np.float32), v)
        np.testing.assert_array_equal(out[3], off)
        np.testing.assert_array_equal(out[4], idx)
        np.testing.assert_array_equal(out[5], val)
    # the reference's loader on the same file
    so = os.path.join(ROOT, "oracle", "_ref", "libloopsref_host.so")
    if os.path.exists(so):
        H = C.CDLL(so)
        a, b, n = C.c_int(), C.c_int(), C.c_int()
        assert H.ref_load_mtx(path.encode(), C.byref(a), C.byref(b), C.byref(n)) == 0
        assert (a.value, b.value, n.value) == (rows, cols, len(idx))
        g_off, g_idx, g_val = np.zeros(rows + 1, np.int32), np.zeros(n.value, np.int32), np.zeros(n.value, np.float32)
        H.ref_loaded_csr(P(g_off), P(g_idx), P(g_val))
        np.testing.assert_array_equal(off, g_off)
        np.testing.assert_array_equal(idx, g_idx)
        np.testing.assert_array_equal(val, g_val)


def test_chesapeake_known_answers(tmp_path, oracle):
    """BASELINE config 1 from a Matrix Market file: the lower triangle of the golden
    CSR written as `pattern symmetric` (the shape of datasets/chesapeake/chesapeake.mtx)."""
    from loops_b200 import market
    c = load_chesapeake()
    rows_of = np.repeat(np.arange(39), np.diff(c["off"]))
    lower = [(int(r) + 1, int(col) + 1, 1.0) for r, col in zip(rows_of, c["idx"]) if col <= r]
    assert len(lower) == 170
    path = str(tmp_path / "chesapeake.mtx")
    write_mtx(path, 39, 39, lower, "pattern", "symmetric")
    R, Cc, r, col, v = market.load_coo(path)
    off, idx, val = market.coo_to_csr(R, Cc, r, col, v)
    assert (R, Cc, len(idx)) == (39, 39, 340)
    np.testing.assert_array_equal(off, c["off"])
    np.testing.assert_array_equal(idx, c["idx"])
    np.testing.assert_array_equal(val, c["val"])
    assert list(off[:6]) == [0, 11, 22, 29, 33, 37]
    y = oracle.spmv(off, idx, val, c["x"])
    assert list(y[:5]) == [50, 52, 53, 26, 18] and float(y.sum()) == 1794.0
    if os.path.exists(EXE):
        rc, out = cpp_load(path)
        assert rc == 0
        np.testing.assert_array_equal(out[3], c["off"])
        np.testing.assert_array_equal(out[4], c["idx"])


BAD = {
    "array": "%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n",
    "complex": "%%MatrixMarket matrix coordinate complex general\n2 2 1\n1 1 1.0 0.0\n",
    "hermitian": "%%MatrixMarket matrix coordinate real hermitian\n2 2 1\n1 1 1.0\n",
    "skew": "%%MatrixMarket matrix coordinate real skew-symmetric\n2 2 1\n2 1 1.0\n",
    "zero-index": "%%MatrixMarket matrix coordinate real general\n2 2 1\n0 1 1.0\n",
    "truncated": "%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1.0\n",
    "no-banner": "2 2 1\n1 1 1.0\n",
    "missing-value": "%%MatrixMarket matrix coordinate real general\n2 2 1\n1 1\n",
    "empty": "",
}


@pytest.mark.parametrize("name", sorted(BAD))
def test_rejections(tmp_path, name):
    from loops_b200 import market
    path = str(tmp_path / "bad.mtx")
    with open(path, "w") as f:
        f.write(BAD[name])
    with pytest.raises(market.MatrixMarketError):
        market.load_coo(path)
    if os.path.exists(EXE):
        rc, msg = cpp_load(path)
        assert rc == 2 and "matrix-market" in msg, (name, rc, msg)


@pytest.mark.gpu
def test_load_csr_runs_on_the_device(tmp_path, oracle):
    import torch
    from loops_b200 import market
    from loops_b200.algorithms import spmv
    c = load_chesapeake()
    rows_of = np.repeat(np.arange(39), np.diff(c["off"]))
    lower = [(int(r) + 1, int(col) + 1, 1.0) for r, col in zip(rows_of, c["idx"]) if col <= r]
    path = str(tmp_path / "chesapeake.mtx")
    write_mtx(path, 39, 39, lower, "pattern", "symmetric", comments=False)
    A = market.load_csr(path)
    y = torch.full((39,), float("nan"), device="cuda")
    spmv.merge_path_flat(A, torch.as_tensor(c["x"]).cuda(), y)
    np.testing.assert_array_equal(y.cpu().numpy(), c["y"])
