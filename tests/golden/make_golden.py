"""Regenerate the golden fixtures under tests/golden/ from the REFERENCE itself.

Stage "host" (runs in the build container, needs /root/reference mounted and
`make -C oracle ref`):
    python tests/golden/make_golden.py host
  battery.npz     the 9 matrices of the reference's standard_battery()
                  (unittests/test_spmv_battery.hxx:52-65) with its input vector
                  (seed 23) and its reference_spmv output, plus the reference's
                  own CSR->ELL and CSR->BCSR{2,3,4} conversions of each
  chesapeake.npz  datasets/chesapeake/chesapeake.mtx through the reference's
                  loader and csr_t(coo) conversion, x = uniform_distribution
                  (x,1,10,42u), y = reference::spmv, spmv_f64, row L1
  xrecipe.npz     first 4096 entries of the x recipe for seeds 42 and 7

Stage "gpu" (runs on the B200 box through gpurun; needs the reference-built
oracle/_ref/libloopsref_gpu.so that travels with the snapshot):
    python tests/golden/make_golden.py gpu gpurun_out/golden
  streams.npz     index streams emitted by the reference's own
                  schedule::setup<> templates for the battery matrices
The files written by the gpu stage are copied from gpurun_out/golden/ into
tests/golden/ and committed.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
P = lambda a: a.ctypes.data_as(C.c_void_p)


def battery_from_reference():
    L = C.CDLL(os.path.join(REFDIR, "libloopsref_battery.so"))
    L.ref_battery_name.restype = C.c_char_p
    out = []
    for i in range(L.ref_battery_count()):
        r, c, n = C.c_int(), C.c_int(), C.c_int()
        L.ref_battery_dims(i, C.byref(r), C.byref(c), C.byref(n))
        off = np.zeros(r.value + 1, np.int32)
        idx = np.zeros(n.value, np.int32)
        val = np.zeros(n.value, np.float32)
        x = np.zeros(c.value, np.float32)
        y = np.zeros(r.value, np.float32)
        L.ref_battery_get(i, P(off), P(idx), P(val), P(x), P(y))
        out.append(dict(name=L.ref_battery_name(i).decode(), rows=r.value, cols=c.value,
                        off=off, idx=idx, val=val, x=x, y=y))
    return out


def host_stage(dst):
    H = C.CDLL(os.path.join(REFDIR, "libloopsref_host.so"))
    arrays = {}
    bat = battery_from_reference()
    arrays["names"] = np.array([b["name"] for b in bat])
    for i, b in enumerate(bat):
        for k in ("off", "idx", "val", "x", "y"):
            arrays[f"{i}_{k}"] = b[k]
        arrays[f"{i}_dims"] = np.array([b["rows"], b["cols"]], np.int32)
        rows, cols, nnz = b["rows"], b["cols"], len(b["idx"])
        pitch = H.ref_ell_pitch(rows, cols, nnz, P(b["off"]), P(b["idx"]), P(b["val"]))
        e_idx = np.zeros(rows * pitch, np.int32)
        e_val = np.zeros(rows * pitch, np.float32)
        H.ref_csr_to_ell(rows, cols, nnz, P(b["off"]), P(b["idx"]), P(b["val"]), P(e_idx), P(e_val))
        arrays[f"{i}_ell_pitch"] = np.array([pitch], np.int32)
        arrays[f"{i}_ell_idx"], arrays[f"{i}_ell_val"] = e_idx, e_val
        coo_rows = np.zeros(nnz, np.int32)
        H.ref_csr_to_coo_rows(rows, cols, nnz, P(b["off"]), P(b["idx"]), P(b["val"]), P(coo_rows))
        arrays[f"{i}_coo_rows"] = coo_rows
        for R in (2, 3, 4):
            nb = C.c_int()
            H.ref_csr_to_bcsr(R, rows, cols, nnz, P(b["off"]), P(b["idx"]), P(b["val"]), C.byref(nb),
                              None, None, None, 0)
            nbr = (rows + R - 1) // R
            b_off = np.zeros(nbr + 1, np.int32)
            b_col = np.zeros(nb.value, np.int32)
            b_val = np.zeros(nb.value * R * R, np.float32)
            H.ref_csr_to_bcsr(R, rows, cols, nnz, P(b["off"]), P(b["idx"]), P(b["val"]), C.byref(nb),
                              P(b_off), P(b_col), P(b_val), nb.value)
            arrays[f"{i}_bcsr{R}_off"], arrays[f"{i}_bcsr{R}_col"], arrays[f"{i}_bcsr{R}_val"] = b_off, b_col, b_val
    np.savez_compressed(os.path.join(dst, "battery.npz"), **arrays)

    r, c, n = C.c_int(), C.c_int(), C.c_int()
    rc = H.ref_load_mtx(b"/root/reference/datasets/chesapeake/chesapeake.mtx", C.byref(r), C.byref(c), C.byref(n))
    assert rc == 0
    off = np.zeros(r.value + 1, np.int32); idx = np.zeros(n.value, np.int32); val = np.zeros(n.value, np.float32)
    H.ref_loaded_csr(P(off), P(idx), P(val))
    x = np.zeros(c.value, np.float32)
    H.ref_x_recipe_int(c.value, 1, 10, 42, P(x))
    y = np.zeros(r.value, np.float32); y64 = np.zeros(r.value, np.float32); l1 = np.zeros(r.value, np.float32)
    H.ref_spmv_f32(r.value, c.value, n.value, P(off), P(idx), P(val), P(x), P(y))
    H.ref_spmv_f64(r.value, c.value, n.value, P(off), P(idx), P(val), P(x), P(y64))
    H.ref_row_l1(r.value, c.value, n.value, P(off), P(idx), P(val), P(x), P(l1))
    np.savez_compressed(os.path.join(dst, "chesapeake.npz"), off=off, idx=idx, val=val, x=x, y=y, y64=y64, l1=l1,
                        dims=np.array([r.value, c.value], np.int32))

    xs = {}
    for seed in (42, 7):
        v = np.zeros(4096, np.float32)
        H.ref_x_recipe_int(4096, 1, 10, seed, P(v))
        xs[f"int_1_10_seed{seed}"] = v
        f = np.zeros(4096, np.float32)
        H.ref_x_recipe_float(4096, C.c_float(0.0), C.c_float(1.0), seed, P(f))
        xs[f"float_0_1_seed{seed}"] = f
    H.ref_hash.restype = C.c_uint
    xs["hash_0_1023"] = np.array([H.ref_hash(i) for i in range(1024)], np.uint32)
    np.savez_compressed(os.path.join(dst, "xrecipe.npz"), **xs)
    print("host golden written to", dst)


# (kind, schedule) cells recorded on the GPU. schedule ids = reference enum.
SCHEDULES = {"merge_path_flat": 0, "work_oriented": 1, "thread_mapped": 2, "group_mapped": 3}


def ref_gpu_emit(G, kind, off, T, A, pitch, sched, grid, tpb, ipt):
    """Call oracle/_ref/libloopsref_gpu.so:ref_gpu_emit; returns dict of arrays."""
    nA = A if kind != 2 else T * pitch
    mk = lambda n, fill=-1: np.full(max(int(n), 1), fill, np.int32)
    visitor, step, tile, visits = mk(nA), mk(nA), mk(nA), mk(nA, 0)
    I = tpb * ipt
    M = (T + nA + I - 1) // I if sched == 0 else 0
    dense = M * I
    d_tile, d_atom, d_emit = mk(dense), mk(dense), mk(dense)
    extra_a = mk(nA)
    extra_b = mk(max(grid * 128 * 4, M * tpb * 2))
    offp = P(off) if off is not None else None
    rc = G.ref_gpu_emit(kind, offp, T, A, pitch, sched, grid, tpb, ipt, P(visitor), P(step), P(tile),
                        P(visits), P(extra_a), P(extra_b), P(d_tile), P(d_atom), P(d_emit),
                        C.c_longlong(dense))
    assert rc == 0, f"ref_gpu_emit rc={rc}"
    res = dict(visitor=visitor[:nA], step=step[:nA], tile=tile[:nA], visits=visits[:nA])
    if sched == 1:
        res["commit"] = extra_a[:nA]
        res["map"] = extra_b[: grid * 128 * 4]
    if sched == 0:
        res.update(dense_tile=d_tile[:dense], dense_atom=d_atom[:dense], dense_emit=d_emit[:dense],
                   thread_start=extra_b[: M * tpb * 2])
    return res


def gpu_stage(dst):
    os.makedirs(dst, exist_ok=True)
    G = C.CDLL(os.path.join(REFDIR, "libloopsref_gpu.so"))
    bat = np.load(os.path.join(ROOT, "tests", "golden", "battery.npz"))
    arrays = {}
    nb = len(bat["names"])
    for i in range(nb):
        off = np.ascontiguousarray(bat[f"{i}_off"])
        T = len(off) - 1
        A = int(off[-1])
        pitch = int(bat[f"{i}_ell_pitch"][0])
        cells = [
            ("csr", 0, off, T, A, 0, "thread_mapped", (T + 127) // 128, 128, 1),
            ("csr", 0, off, T, A, 0, "thread_mapped", 2, 32, 1),
            ("csr", 0, off, T, A, 0, "group_mapped", 0, 128, 1),
            ("csr", 0, off, T, A, 0, "work_oriented", 1, 128, 1),
            ("csr", 0, off, T, A, 0, "work_oriented", 3, 128, 1),
            ("csr", 0, off, T, A, 0, "merge_path_flat", 0, 128, 8),
            ("csr", 0, off, T, A, 0, "merge_path_flat", 0, 128, 7),
            ("coo", 1, None, A, A, 1, "thread_mapped", (A + 127) // 128, 128, 1),
            ("coo", 1, None, A, A, 1, "merge_path_flat", 0, 128, 8),
            ("coo", 1, None, A, A, 1, "work_oriented", 2, 128, 1),
            ("coo", 1, None, A, A, 1, "group_mapped", 0, 128, 1),
            ("ell", 2, None, T, T * pitch, pitch, "thread_mapped", (T + 127) // 128, 128, 1),
            ("ell", 2, None, T, T * pitch, pitch, "merge_path_flat", 0, 128, 5),
            ("ell", 2, None, T, T * pitch, pitch, "work_oriented", 2, 128, 1),
            ("ell", 2, None, T, T * pitch, pitch, "group_mapped", 0, 128, 1),
        ]
        for (lname, kind, o, t, a, p, sname, grid, tpb, ipt) in cells:
            if grid == 0 and sname == "group_mapped":
                grid = (t + 127) // 128
            if grid == 0 and sname == "merge_path_flat":
                grid = 1
            if grid == 0:
                continue
            res = ref_gpu_emit(G, kind, o, t, a if kind != 2 else 0, p, SCHEDULES[sname], grid, tpb, ipt)
            key = f"{i}|{lname}|{sname}|{grid}|{tpb}|{ipt}"
            for k, v in res.items():
                arrays[f"{key}|{k}"] = v
    arrays["wo_grid_reference"] = np.array([G.ref_gpu_work_oriented_grid()], np.int32)
    np.savez_compressed(os.path.join(dst, "streams.npz"), **arrays)
    print("gpu golden written to", dst, "with", len(arrays), "arrays")


if __name__ == "__main__":
    stage = sys.argv[1] if len(sys.argv) > 1 else "host"
    if stage == "host":
        host_stage(os.path.join(ROOT, "tests", "golden"))
    else:
        gpu_stage(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "golden"))
