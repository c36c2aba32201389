"""Matrix Market (coordinate) loader, the Python mirror of
``include/loops/container/market.hxx`` (reference container/market.hxx:100-289,
detail/mtx_parser.hxx:150-211): ``matrix coordinate {real,integer,pattern}
{general,symmetric}``; pattern entries get value 1; symmetric files mirror every
off-diagonal entry right after its source; complex / hermitian / skew-symmetric /
array files are rejected. Host-side set-up code (numpy), not the SpMV hot path."""
from __future__ import annotations

import numpy as np


class MatrixMarketError(ValueError):
    pass


def load_coo(path: str):
    """-> (rows, cols, row_indices i32, col_indices i32, values f32) in file order."""
    with open(path, "rb") as f:
        text = f.read().decode("ascii", errors="replace")
    if not text:
        raise MatrixMarketError(f"matrix-market: empty file {path}")
    lines = text.split("\n")
    banner = lines[0].split()
    if len(banner) < 5 or banner[0].lower() != "%%matrixmarket":
        raise MatrixMarketError(f"matrix-market: missing %%MatrixMarket banner in {path}")
    obj, fmt, field, sym = (t.lower() for t in banner[1:5])
    if obj != "matrix":
        raise MatrixMarketError(f"matrix-market: object must be 'matrix' in {path}")
    if fmt != "coordinate":
        raise MatrixMarketError(f"matrix-market: only the coordinate (sparse) format is supported in {path}")
    if field == "complex":
        raise MatrixMarketError(f"matrix-market: complex values not supported in {path}")
    if sym in ("hermitian", "skew-symmetric"):
        raise MatrixMarketError(f"matrix-market: hermitian / skew-symmetric not supported in {path}")
    if sym not in ("general", "symmetric"):
        raise MatrixMarketError(f"matrix-market: missing or unrecognized symmetry tag in {path}")
    if field not in ("real", "integer", "pattern"):
        raise MatrixMarketError(f"matrix-market: missing or unrecognized field tag in {path}")
    body = [ln for ln in lines[1:] if ln.strip() and not ln.lstrip().startswith("%")]
    if not body:
        raise MatrixMarketError("matrix-market: expected dimension line (first integer)")
    dims = body[0].split()
    if len(dims) < 3 or not all(d.isdigit() for d in dims[:3]):
        raise MatrixMarketError("matrix-market: expected dimension line")
    rows, cols, nnz = (int(d) for d in dims[:3])
    if len(body) - 1 < nnz:
        raise MatrixMarketError("matrix-market: expected row index in body")
    pattern = field == "pattern"
    r = np.empty(nnz, np.int64)
    c = np.empty(nnz, np.int64)
    v = np.ones(nnz, np.float64)
    for i in range(nnz):
        tok = body[1 + i].split()
        if len(tok) < (2 if pattern else 3) or not tok[0].isdigit() or not tok[1].isdigit():
            raise MatrixMarketError("matrix-market: expected row index / column index / value in body")
        r[i], c[i] = int(tok[0]), int(tok[1])
        if not pattern:
            try:
                v[i] = float(tok[2])
            except ValueError:
                raise MatrixMarketError("matrix-market: expected value in body") from None
    if nnz and (r.min() == 0 or c.min() == 0):
        raise MatrixMarketError("matrix-market: zero-indexed entry (Matrix Market is 1-indexed)")
    if nnz and (r.max() > rows or c.max() > cols):
        raise MatrixMarketError(f"matrix-market: entry outside the declared dimensions in {path}")
    r -= 1
    c -= 1
    if sym == "symmetric":
        off = r != c
        reps = 1 + off.astype(np.int64)                       # the mirror sits right after its source
        src = np.repeat(np.arange(nnz), reps)
        is_mirror = np.concatenate([[False], src[1:] == src[:-1]]) if len(src) else np.zeros(0, bool)
        r2, c2 = r[src].copy(), c[src].copy()
        r2[is_mirror], c2[is_mirror] = c[src][is_mirror], r[src][is_mirror]
        r, c, v = r2, c2, v[src]
    return rows, cols, r.astype(np.int32), c.astype(np.int32), v.astype(np.float32)


def coo_to_csr(rows, cols, r, c, v):
    """csr_t(coo_t) (reference container/csr.hxx:86-94): stable sort by (row, col),
    duplicates preserved."""
    order = np.lexsort((c, r))
    off = np.zeros(rows + 1, np.int64)
    np.add.at(off, r.astype(np.int64) + 1, 1)
    return np.cumsum(off).astype(np.int32), c[order].astype(np.int32), v[order].astype(np.float32)


def load_csr(path: str, device="cuda"):
    """Matrix Market file -> ``csr_t`` on ``device`` (examples/spmv/helpers.hxx flow)."""
    from .container import coo_t, csr_t
    rows, cols, r, c, v = load_coo(path)
    if str(device).startswith("cuda"):
        # the triples go up in file order; sorting and row compression run on the
        # device (loopsb_coo_to_csr: stable radix sort by (row, col), duplicates kept)
        from .convert import coo_to_csr as coo_to_csr_device
        return coo_to_csr_device(coo_t(rows, cols, r, c, v, device=device))
    off, idx, val = coo_to_csr(rows, cols, r, c, v)
    return csr_t(rows, cols, off, idx, val, device=device)
