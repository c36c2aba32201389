"""tools/convert_bench.py -- device format conversions on the
BASELINE config-2 matrix (2^20 rows / 2^25 nnz), timed with the reference's own
host converters (oracle/_ref, the unmodified headers) beside them.

GPU: median of 5 calls after one warm-up, wall clock around the synchronous C-ABI
call (each call allocates its temporaries and synchronises, as a user sees it).
CPU: one call of the reference converter on a 2^17-row / 2^22-nnz matrix from the
same generator (the full size takes minutes on the host); both are reported as
nnz/s so the sample size cancels. `GB/s` counts the arrays a conversion must read
and write once (algorithmic bytes), not its temporaries."""
import ctypes as C
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from loops_b200 import convert, csr_t, generate as g

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = lambda a: a.ctypes.data_as(C.c_void_p)


def gpu_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(); torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
        del out
    return statistics.median(ts)


def main():
    rows = cols = 1 << 20; nnz = 1 << 25
    off, idx, val = g.synth_csr(rows, cols, nnz, device="cuda")
    A = csr_t.from_tensors(rows, cols, off, idx, val)
    coo = convert.csr_to_coo(A)
    pitch = convert.csr_max_degree(A)
    nb = convert.csr_to_bcsr(A, 4, 4).num_blocks
    res = {}

    def rec(name, sec, nbytes, note=""):
        res[name] = {"gpu_ms": sec * 1e3, "gpu_nnz_per_s": nnz / sec, "gpu_GBps_algorithmic": nbytes / sec / 1e9,
                     "note": note}

    csr_b = nnz * 8 + (rows + 1) * 4
    rec("csr_to_coo", gpu_time(lambda: convert.csr_to_coo(A)), (rows + 1) * 4 + nnz * 4)
    rec("coo_to_csr", gpu_time(lambda: convert.coo_to_csr(coo)), nnz * 12 + csr_b, "input already (row, col)-sorted")
    rec("csr_to_csc", gpu_time(lambda: convert.csr_to_csc(A)), 2 * csr_b)
    rec("csr_to_ell", gpu_time(lambda: convert.csr_to_ell(A), reps=3), csr_b + rows * pitch * 8, f"pitch {pitch}")
    rec("csr_to_bcsr4x4_f32", gpu_time(lambda: convert.csr_to_bcsr(A, 4, 4)), csr_b + nb * 68, f"{nb} blocks")
    rec("csr_to_bcsr4x4_bf16", gpu_time(lambda: convert.csr_to_bcsr(A, 4, 4, torch.bfloat16)), csr_b + nb * 36)

    # ---- the reference's host converters on a smaller matrix of the same family ----
    so = os.path.join(ROOT, "oracle", "_ref", "libloopsref_host.so")
    if os.path.exists(so):
        L = C.CDLL(so)
        r2 = c2 = 1 << 17; n2 = 1 << 22
        o2, i2, v2 = (t.cpu().numpy() for t in g.synth_csr(r2, c2, n2, device="cpu"))

        def cpu(name, fn):
            t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
            res[name].update({"ref_cpu_s_on_sample": dt, "ref_cpu_nnz_per_s": n2 / dt,
                              "speedup_nnz_per_s": res[name]["gpu_nnz_per_s"] / (n2 / dt)})

        rid = np.zeros(n2, np.int32)
        cpu("csr_to_coo", lambda: L.ref_csr_to_coo_rows(r2, c2, n2, P(o2), P(i2), P(v2), P(rid)))
        co, cr, cv = np.zeros(c2 + 1, np.int32), np.zeros(n2, np.int32), np.zeros(n2, np.float32)
        cpu("csr_to_csc", lambda: L.ref_csr_to_csc(r2, c2, n2, P(o2), P(i2), P(v2), P(co), P(cr), P(cv)))
        L.ref_ell_pitch.restype = C.c_int
        p2 = L.ref_ell_pitch(r2, c2, n2, P(o2), P(i2), P(v2))
        ei, ev = np.zeros(r2 * p2, np.int32), np.zeros(r2 * p2, np.float32)
        cpu("csr_to_ell", lambda: L.ref_csr_to_ell(r2, c2, n2, P(o2), P(i2), P(v2), P(ei), P(ev)))
        nbk = C.c_int(0)
        cpu("csr_to_bcsr4x4_f32", lambda: L.ref_csr_to_bcsr(4, r2, c2, n2, P(o2), P(i2), P(v2), C.byref(nbk),
                                                             None, None, None, 0))
        res["_cpu_sample"] = {"rows": r2, "nnz": n2, "host_cores_used": 1,
                              "what": "reference converters as shipped (single-threaded host loops / thrust host sorts)"}
    res["_workload"] = {"rows": rows, "nnz": nnz, "ell_pitch": pitch, "bcsr4x4_blocks": nb}
    for k, v in res.items():
        if not k.startswith("_"):
            print(f"{k:22s} {v['gpu_ms']:9.3f} ms  {v['gpu_nnz_per_s']/1e9:7.2f} Gnnz/s  {v['gpu_GBps_algorithmic']:8.1f} GB/s"
                  + (f"   ref CPU {v['ref_cpu_nnz_per_s']/1e6:8.1f} Mnnz/s  x{v['speedup_nnz_per_s']:.0f}" if "ref_cpu_nnz_per_s" in v else "")
                  + (f"   ({v['note']})" if v["note"] else ""))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
