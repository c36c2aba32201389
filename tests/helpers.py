"""Test-side helpers: ctypes wrapper of the oracle, fixture loaders, matrix
factories. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may touch oracle/."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
P = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)

SCHED = {"merge_path_flat": 0, "work_oriented": 1, "thread_mapped": 2, "group_mapped": 3}
ORC_OFFSETS, ORC_COO, ORC_PITCH, ORC_FLAT = 0, 1, 2, 3


class OrcLayout(C.Structure):
    _fields_ = [("kind", C.c_int32), ("offsets", C.c_void_p), ("num_tiles", C.c_int32),
                ("num_atoms", C.c_int32), ("pitch", C.c_int32)]


class RigorousReport(C.Structure):
    _fields_ = [("total_rows", C.c_int64), ("naive_mismatches", C.c_int64),
                ("f32_baseline_overruns", C.c_int64), ("gpu_overruns", C.c_int64),
                ("max_gpu_abs_error", C.c_double), ("max_gpu_rel_error", C.c_double),
                ("wilkinson_k", C.c_double)]


class Oracle:
    def __init__(self, path):
        self.L = C.CDLL(path)
        self.L.orc_emit_merge_path.restype = C.c_int64
        self.L.orc_hash.restype = C.c_uint32
        self.L.orc_bf16_round.restype = C.c_float
        self.L.orc_bf16_round.argtypes = [C.c_float]
        self.L.orc_count_errors.restype = C.c_int64

    # -- layouts ---------------------------------------------------------
    @staticmethod
    def layout(kind, offsets=None, num_tiles=0, num_atoms=0, pitch=0):
        l = OrcLayout()
        l.kind, l.num_tiles, l.num_atoms, l.pitch = kind, num_tiles, num_atoms, pitch
        l._keep = offsets
        l.offsets = offsets.ctypes.data if offsets is not None else None
        return l

    @classmethod
    def csr_layout(cls, off):
        off = np.ascontiguousarray(off, np.int32)
        return cls.layout(ORC_OFFSETS, off, len(off) - 1, int(off[-1]) if len(off) else 0)

    @classmethod
    def coo_layout(cls, nnz):
        return cls.layout(ORC_COO, None, nnz, nnz, 1)

    @classmethod
    def ell_layout(cls, rows, pitch):
        return cls.layout(ORC_PITCH, None, rows, rows * pitch, pitch)

    def num_atoms(self, lay):
        return self.L.orc_num_atoms(C.byref(lay))

    def num_tiles(self, lay):
        return self.L.orc_num_tiles(C.byref(lay))

    # -- emit ------------------------------------------------------------
    def emit(self, lay, schedule, grid_blocks=0, tpb=128, ipt=8):
        A, T = self.num_atoms(lay), self.num_tiles(lay)
        mk = lambda n, f=-1: np.full(max(int(n), 1), f, np.int32)
        visitor, step, tile, visits = mk(A), mk(A), mk(A), mk(A, 0)
        res = {}
        if schedule == SCHED["thread_mapped"]:
            self.L.orc_emit_thread_mapped(C.byref(lay), C.c_int64(grid_blocks * tpb), P(visitor), P(step),
                                          P(tile), P(visits))
        elif schedule == SCHED["group_mapped"]:
            self.L.orc_emit_group_mapped(C.byref(lay), 128, P(visitor), P(step), P(tile), P(visits))
        elif schedule == SCHED["work_oriented"]:
            commit, m = mk(A), mk(grid_blocks * 128 * 4)
            self.L.orc_emit_work_oriented(C.byref(lay), C.c_int64(grid_blocks * 128), P(visitor), P(step),
                                          P(tile), P(visits), P(commit), P(m))
            res["commit"], res["map"] = commit[:A], m[: grid_blocks * 128 * 4]
        elif schedule == SCHED["merge_path_flat"]:
            I = tpb * ipt
            M = (T + A + I - 1) // I
            dense = M * I
            dt, da, de = mk(dense), mk(dense), mk(dense)
            coords, ts = mk(2 * (M + 1)), mk(M * tpb * 2)
            m2 = self.L.orc_emit_merge_path(C.byref(lay), tpb, ipt, P(visitor), P(step), P(tile), P(visits),
                                            P(dt), P(da), P(de), P(coords), P(ts))
            assert m2 == M
            res.update(dense_tile=dt[:dense], dense_atom=da[:dense], dense_emit=de[:dense],
                       coords=coords[: 2 * (M + 1)].reshape(-1, 2), thread_start=ts[: M * tpb * 2])
        else:
            raise ValueError(schedule)
        res.update(visitor=visitor[:A], step=step[:A], tile=tile[:A], visits=visits[:A])
        return res

    # -- validator ---------------------------------------------------------
    def spmv(self, off, idx, val, x):
        rows = len(off) - 1
        y = np.zeros(rows, np.float32)
        self.L.orc_spmv_f32(rows, P(off), P(idx), P(val), P(x), P(y))
        return y

    def spmv_f64(self, off, idx, val, x):
        rows = len(off) - 1
        y = np.zeros(rows, np.float32)
        self.L.orc_spmv_f64(rows, P(off), P(idx), P(val), P(x), P(y))
        return y

    def spmv_d(self, off, idx, val64, x64):
        """reference::spmv<double>: fp64 values, x, accumulator and y."""
        rows = len(off) - 1
        y = np.zeros(rows, np.float64)
        self.L.orc_spmv_d(rows, P(off), P(idx), P(np.ascontiguousarray(val64, np.float64)),
                          P(np.ascontiguousarray(x64, np.float64)), P(y))
        return y

    def row_l1(self, off, idx, val, x):
        rows = len(off) - 1
        y = np.zeros(rows, np.float32)
        self.L.orc_row_l1(rows, P(off), P(idx), P(val), P(x), P(y))
        return y

    def count_errors(self, y, ref):
        return int(self.L.orc_count_errors(P(y), P(ref), C.c_int64(len(ref))))

    def rigorous(self, off, idx, val, x, y_gpu, k=8.0, atol=1e-3):
        rep = RigorousReport()
        self.L.orc_rigorous_validate(len(off) - 1, P(off), P(idx), P(val), P(x), P(y_gpu),
                                     C.c_double(k), C.c_double(atol), C.byref(rep))
        return rep

    def x_recipe_int(self, n, lo=1, hi=10, seed=42):
        out = np.zeros(n, np.float32)
        self.L.orc_x_recipe_int(n, lo, hi, C.c_uint32(seed), P(out))
        return out

    def x_recipe_float(self, n, lo, hi, seed):
        out = np.zeros(n, np.float32)
        self.L.orc_x_recipe_float(n, C.c_float(lo), C.c_float(hi), C.c_uint32(seed), P(out))
        return out

    # -- conversions -------------------------------------------------------
    def coo_rows(self, off):
        rows = len(off) - 1
        out = np.zeros(int(off[-1]), np.int32)
        self.L.orc_csr_to_coo_rows(rows, P(off), P(out))
        return out

    def ell(self, off, idx, val):
        rows = len(off) - 1
        pitch = self.L.orc_ell_pitch(rows, P(off))
        e_idx = np.zeros(rows * pitch, np.int32)
        e_val = np.zeros(rows * pitch, np.float32)
        self.L.orc_csr_to_ell(rows, P(off), P(idx), P(val), pitch, P(e_idx), P(e_val))
        return pitch, e_idx, e_val

    def bcsr(self, R, Cc, rows, cols, off, idx, val):
        nb = C.c_int32()
        self.L.orc_csr_to_bcsr(R, Cc, rows, cols, P(off), P(idx), P(val), C.byref(nb), None, None, None)
        nbr = (rows + R - 1) // R
        b_off = np.zeros(nbr + 1, np.int32)
        b_col = np.zeros(nb.value, np.int32)
        b_val = np.zeros(nb.value * R * Cc, np.float32)
        self.L.orc_csr_to_bcsr(R, Cc, rows, cols, P(off), P(idx), P(val), C.byref(nb), P(b_off), P(b_col), P(b_val))
        return b_off, b_col, b_val

    def csc(self, rows, cols, off, idx, val):
        nnz = int(off[-1])
        c_off = np.zeros(cols + 1, np.int32)
        c_row = np.zeros(nnz, np.int32)
        c_val = np.zeros(nnz, np.float32)
        self.L.orc_csr_to_csc(rows, cols, P(off), P(idx), P(val), P(c_off), P(c_row), P(c_val))
        return c_off, c_row, c_val

    def spmv_csc(self, rows, c_off, c_row, c_val, x):
        y = np.zeros(rows, np.float32)
        self.L.orc_spmv_csc(len(c_off) - 1, P(c_off), P(c_row), P(c_val), P(x), P(y))
        return y

    def dia(self, rows, off, idx, val):
        """(diag_offsets i32[nd], values f32[nd * rows]) -- stride == rows."""
        nd = self.L.orc_dia_offsets(rows, P(off), P(idx), None, 0)
        d_off = np.zeros(max(nd, 1), np.int32)[:nd].copy() if nd else np.zeros(0, np.int32)
        if nd:
            self.L.orc_dia_offsets(rows, P(off), P(idx), P(d_off), nd)
        d_val = np.zeros(nd * rows, np.float32)
        if nd:
            self.L.orc_csr_to_dia(rows, P(off), P(idx), P(val), nd, P(d_off), P(d_val))
        return d_off, d_val

    def spmv_dia(self, rows, cols, d_off, d_val, x):
        y = np.zeros(rows, np.float32)
        self.L.orc_spmv_dia(rows, cols, C.c_int64(rows), len(d_off), P(d_off), P(d_val), P(x), P(y))
        return y

    def spmv_flat_partitioned(self, K, off, idx, val, x):
        rows = len(off) - 1
        y = np.zeros(rows, np.float32)
        self.L.orc_spmv_flat_partitioned(rows, K, P(off), P(idx), P(val), P(x), P(y))
        return y

    def spmm(self, off, idx, val, B):
        rows, n = len(off) - 1, B.shape[1]
        B = np.ascontiguousarray(B, np.float32)
        out = np.zeros((rows, n), np.float32)
        self.L.orc_spmm(rows, n, P(off), P(idx), P(val), P(B), P(out))
        return out

    def spmv_coo(self, rows, row, col, val, x):
        y = np.zeros(rows, np.float32)
        self.L.orc_spmv_coo(C.c_int64(len(val)), P(row), P(col), P(val), P(x), P(y))
        return y

    def spmv_ell(self, rows, pitch, e_idx, e_val, x):
        y = np.zeros(rows, np.float32)
        self.L.orc_spmv_ell(rows, pitch, P(e_idx), P(e_val), P(x), P(y))
        return y

    def spmv_bcsr(self, R, Cc, rows, b_off, b_col, b_val, x_padded):
        y = np.zeros(rows, np.float32)
        self.L.orc_spmv_bcsr(R, Cc, rows, len(b_off) - 1, P(b_off), P(b_col), P(b_val), P(x_padded), P(y))
        return y

    def bf16_round(self, a):
        return np.array([self.L.orc_bf16_round(float(v)) for v in np.asarray(a).ravel()],
                        np.float32).reshape(np.shape(a))


def load_battery():
    """The reference's 9-matrix battery (tests/golden/battery.npz)."""
    z = np.load(os.path.join(GOLDEN, "battery.npz"))
    out = []
    for i, name in enumerate(z["names"]):
        d = {k: np.ascontiguousarray(z[f"{i}_{k}"]) for k in
             ("off", "idx", "val", "x", "y", "ell_idx", "ell_val", "coo_rows")}
        d["name"] = str(name)
        d["rows"], d["cols"] = (int(v) for v in z[f"{i}_dims"])
        d["ell_pitch"] = int(z[f"{i}_ell_pitch"][0])
        for R in (2, 3, 4):
            d[f"bcsr{R}"] = tuple(np.ascontiguousarray(z[f"{i}_bcsr{R}_{k}"]) for k in ("off", "col", "val"))
        out.append(d)
    return out


def load_chesapeake():
    z = np.load(os.path.join(GOLDEN, "chesapeake.npz"))
    return {k: np.ascontiguousarray(z[k]) for k in z.files}


def random_csr(rows, cols, density, seed, empty_every=0, heavy_row=None, exact=False):
    """Small CSR factory for extra cases (numpy RNG; not a reference fixture)."""
    rng = np.random.default_rng(seed)
    off = [0]
    idx, val = [], []
    for r in range(rows):
        if empty_every and r % empty_every == 0:
            off.append(off[-1]); continue
        if heavy_row is not None and r == heavy_row[0]:
            n = min(cols, heavy_row[1])
        else:
            n = rng.binomial(cols, density)
        c = np.sort(rng.choice(cols, size=n, replace=False)).astype(np.int32)
        idx.append(c)
        if exact:
            val.append((rng.integers(1, 17, size=n) / 8.0).astype(np.float32))
        else:
            val.append(rng.uniform(0.5, 1.5, size=n).astype(np.float32))
        off.append(off[-1] + n)
    idx = np.concatenate(idx) if idx else np.zeros(0, np.int32)
    val = np.concatenate(val) if val else np.zeros(0, np.float32)
    return (np.array(off, np.int32), np.ascontiguousarray(idx, np.int32), np.ascontiguousarray(val, np.float32))
