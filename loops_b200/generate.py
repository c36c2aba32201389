"""Deterministic synthetic inputs:

* ``powerlaw_degrees``  truncated discrete power law, rescaled so the degrees
  sum to ``nnz`` exactly, seeded random row order (no degree sorting);
* ``synth_csr``         CSR with unique, ascending columns per row drawn by a
  counter-based hash (stratified over the column range), values ``k/8`` with
  ``k in 1..16`` (exact in fp32 and bf16);
* ``x_recipe``          the reference's input vector recipe, bit for bit:
  ``generate::random::uniform_distribution(x, 1, 10, 42u)`` (reference
  util/generate.hxx:33-41,54-79 as called at examples/spmv/merge_path.cu:33-34):
  per index a minstd_rand seeded with ``hash(i) * seed``, first draw mapped by
  thrust's uniform_int_distribution -> integers 1..10 stored as fp32.

Everything is written with torch ops on int64 so the same code runs on CPU
(tests) and on the GPU (bench, where 2^25..2^29 nonzeros are generated in HBM).
"""
from __future__ import annotations

import numpy as np
import torch

MATRIX_SEED = 0x5EED0001
X_SEED = 42

_M64 = (1 << 64)


def _s64(c: int) -> int:
    c &= _M64 - 1
    return c - _M64 if c >= (1 << 63) else c


def _lsr(z: torch.Tensor, s: int) -> torch.Tensor:
    return (z >> s) & ((1 << (64 - s)) - 1)


def mix64(z: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (two's-complement wrap-around)."""
    z = z + _s64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def mix64_np(z: np.ndarray) -> np.ndarray:
    """numpy/uint64 twin of mix64 (used by tests to cross-check)."""
    z = z.astype(np.uint64)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def powerlaw_degrees(rows: int, nnz: int, gamma: float = 2.1, d_max: int = 1024,
                     spread: float = 1024.0, seed: int = MATRIX_SEED) -> np.ndarray:
    """int64[rows], each in [1, d_max], sum == nnz. P(d) ~ d^-gamma."""
    if rows == 0:
        return np.zeros(0, dtype=np.int64)
    d_max = int(min(d_max, max(1, nnz)))
    if not (rows <= nnz <= rows * d_max):
        raise ValueError("need rows <= nnz <= rows*d_max")
    h = mix64_np(np.arange(rows, dtype=np.uint64) ^ np.uint64(seed * 0x9E3779B1 & (_M64 - 1)))
    u = ((h >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)       # (0,1)
    a = gamma - 1.0
    raw = (1.0 - u * (1.0 - spread ** (-a))) ** (-1.0 / a)                      # Pareto on [1, spread]
    lo, hi = 1e-6, float(d_max)
    for _ in range(80):                                                          # bisection on the scale
        mid = 0.5 * (lo + hi)
        s = np.clip(np.rint(mid * raw), 1, d_max).sum()
        if s < nnz:
            lo = mid
        else:
            hi = mid
    deg = np.clip(np.rint(lo * raw), 1, d_max).astype(np.int64)
    order = np.argsort(h, kind="stable")                                        # hashed row order
    gap = int(nnz - deg.sum())
    pos = 0
    while gap != 0:                                                              # settle the residual
        step = 1 if gap > 0 else -1
        ok = order[(deg[order] + step >= 1) & (deg[order] + step <= d_max)]
        take = ok[pos: pos + abs(gap)]
        if take.size == 0:
            pos = 0
            continue
        deg[take] += step
        gap -= step * take.size
    assert deg.sum() == nnz and deg.min() >= 1 and deg.max() <= d_max
    return deg


def synth_csr(rows: int, cols: int, nnz: int, seed: int = MATRIX_SEED, device="cpu",
              degrees: np.ndarray | None = None, row_begin: int = 0, row_end: int | None = None,
              gamma: float = 2.1, d_max: int = 1024):
    """Rows [row_begin, row_end) of the synthetic matrix as local CSR tensors
    (offsets rebased to 0, GLOBAL column ids): (offsets i32, indices i32,
    values f32). The full matrix is the concatenation over row ranges."""
    if degrees is None:
        degrees = powerlaw_degrees(rows, nnz, gamma=gamma, d_max=min(d_max, cols), seed=seed)
    row_end = rows if row_end is None else row_end
    goff = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(degrees, out=goff[1:])
    first = int(goff[row_begin])
    local_n = int(goff[row_end] - first)
    deg_t = torch.as_tensor(degrees[row_begin:row_end], dtype=torch.int64, device=device)
    off_t = torch.zeros(row_end - row_begin + 1, dtype=torch.int64, device=device)
    torch.cumsum(deg_t, 0, out=off_t[1:])
    nloc = row_end - row_begin
    row_of = torch.repeat_interleave(torch.arange(nloc, device=device, dtype=torch.int64), deg_t,
                                     output_size=local_n)
    e = torch.arange(local_n, device=device, dtype=torch.int64)
    k = e - off_t[row_of]                      # slot inside the row
    d = deg_t[row_of]
    lo = (k * cols) // d                        # stratum [lo, hi) of the column range
    hi = ((k + 1) * cols) // d
    hsh = mix64((e + first) ^ _s64(seed * 0xD6E8FEB86659FD93))
    col = lo + _lsr(hsh, 1) % (hi - lo)
    val = ((_lsr(hsh, 40) % 16) + 1).to(torch.float32) / 8.0
    return off_t.to(torch.int32), col.to(torch.int32), val


def hash32(a: torch.Tensor) -> torch.Tensor:
    """generate::random::hash (reference util/generate.hxx:33-41) on uint32
    values carried in int64 tensors."""
    m = 0xFFFFFFFF
    a = ((a + 0x7ED55D16) + (a << 12)) & m
    a = ((a ^ 0xC761C23C) ^ (a >> 19)) & m
    a = ((a + 0x165667B1) + (a << 5)) & m
    a = ((a + 0xD3A2646C) ^ (a << 9)) & m
    a = ((a + 0xFD7046C5) + (a << 3)) & m
    a = ((a ^ 0xB55A4F09) ^ (a >> 16)) & m
    return a


def x_recipe(n: int, lo: int = 1, hi: int = 10, seed: int = X_SEED, device="cpu") -> torch.Tensor:
    """fp32[n] holding the integers the reference puts in x."""
    i = torch.arange(n, dtype=torch.int64, device=device) & 0xFFFFFFFF
    s = (hash32(i) * seed) & 0xFFFFFFFF
    m = 2147483647
    st = s % m
    st = torch.where(st == 0, torch.ones_like(st), st)     # engine never sits at 0
    draw = (48271 * st) % m                                # minstd_rand, first output
    u = (draw - 1).to(torch.float64) / float(m - 1)        # [0, 1)
    r = u * float(hi + 1 - lo) + float(lo)
    return r.to(torch.int64).to(torch.float32)
