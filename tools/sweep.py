"""tools/sweep.py -- BASELINE configs 3 and 4 on one B200:
  config 3: {thread_mapped, group_mapped, work_oriented, merge_path_flat} x
            {csr, coo, ell} on the 2^20-row / 2^25-nnz synthetic matrix: all 12 cells
            (the reference itself ships kernels for 7 of them),
  config 4: BCSR 4x4 bf16 on tcgen05, 262,144 block-rows / 8,388,608 blocks,
and, for context, the reference's own kernels (oracle/_ref/libloopsref_gpu.so,
built from its unmodified headers with -DLOOPS_TARGET_ARCH=100) on the same box.
Every cell is checked against y of merge_path_flat/CSR (exact inputs -> equal
bits) before it is timed. Prints one JSON document.
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from loops_b200 import _lib, csr_t, generate as g
from loops_b200.algorithms import spmv
from loops_b200.container import bcsr_t, csr_to_coo_device, csr_to_ell_device

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def time_ms(fn, warm=5, reps=5, batch=20):
    """Median / best over `reps` batches of `batch` back-to-back launches, one CUDA event pair
    per batch on the current stream (how bench.py times its step): launch latency is hidden
    behind the previous launch, as in any real use, instead of being added to every sample."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(batch):
            fn()
        b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) / batch)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def run(small=False, reference=True, log=sys.stderr):
    """All of BASELINE configs[2] and configs[3] on the current device; returns the result document.
    `reference=False` leaves out the reference's own kernels (bench.py's `extra` block: product code only)."""
    rows = cols = (1 << 16) if small else (1 << 20)
    nnz = rows * 32
    out = {"rows": rows, "nnz": nnz, "peak_gbs": PEAK, "cells": [], "reference_gpu": [], "bcsr": None,
           "timing": "median of 5 batches of 20 back-to-back launches, one CUDA event pair per batch"}
    off, idx, val = g.synth_csr(rows, cols, nnz, device="cuda")
    A = csr_t.from_tensors(rows, cols, off, idx, val)
    x = g.x_recipe(cols, device="cuda")
    y0 = torch.empty(rows, device="cuda")
    spmv.merge_path_flat(A, x, y0, tiled=False)
    bytes_csr = nnz * 8 + (rows + 1) * 4 + cols * 4 + rows * 4
    bytes_coo = nnz * 12 + cols * 4 + rows * 4

    def cell(layout, sched, fn, container, nbytes):
        y = torch.full((rows,), float("nan"), device="cuda")
        fn(container, x, y)
        ok = bool(torch.equal(y, y0))
        med, best = time_ms(lambda: fn(container, x, y, sync=False))
        out["cells"].append({"layout": layout, "schedule": sched, "ms_median": med, "ms_min": best,
                             "gnnz_per_s": nnz / med / 1e6, "algorithmic_gb_per_s": nbytes / med / 1e6,
                             "roofline_frac": nbytes / med / 1e6 / PEAK, "bit_equal_to_merge_csr": ok})
        print(f"{layout:4s} {sched:16s} {med*1e3:9.1f} us  {nnz/med/1e6:7.1f} Gnnz/s  {nbytes/med/1e6:7.0f} GB/s  ok={ok}",
              file=log)

    scheds = ("merge_path_flat", "work_oriented", "group_mapped", "thread_mapped")
    for name in scheds:
        fn = spmv.CELLS[("csr", name)]
        if name == "merge_path_flat":     # the kernel that reads the CSR arrays; the plan-owned tiled copy is its own line
            fn = lambda c, xx, yy, sync=True: spmv.merge_path_flat(c, xx, yy, sync=sync, tiled=False)
        cell("csr", name, fn, A, bytes_csr)
    cell("csr", "merge_path_flat+band_tiled_plan",
         lambda c, xx, yy, sync=True: spmv.merge_path_flat(c, xx, yy, sync=sync, tiled="auto"), A, bytes_csr)
    spmv.merge_path_flat(A, x, y0, tiled=False)   # drop the tiled copy again (290 MB)
    coo = csr_to_coo_device(A)
    for name in scheds:
        cell("coo", name, spmv.CELLS[("coo", name)], coo, bytes_coo)
    ell = csr_to_ell_device(A)
    bytes_ell = rows * ell.pitch * 8 + cols * 4 + rows * 4
    out["ell_pitch"] = ell.pitch
    for name in scheds:
        cell("ell", name, spmv.CELLS[("ell", name)], ell, bytes_ell)
    del ell, coo
    torch.cuda.empty_cache()

    # ---- reference kernels on the same GPU (context: the kernels to beat) ----
    so = os.path.join(ROOT, "oracle", "_ref", "libloopsref_gpu.so")
    if reference and os.path.exists(so):
        G = C.CDLL(so)
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        ho, hi, hv, hx = off.cpu().numpy(), idx.cpu().numpy(), val.cpu().numpy(), x.cpu().numpy()
        y0h = y0.cpu().numpy()
        names = ["merge_path_flat", "work_oriented", "thread_mapped", "group_mapped", "coo_thread_mapped"]
        for which, nm in enumerate(names):
            yr = np.zeros(rows, np.float32)
            ms, inner = C.c_float(), C.c_float()
            rc = G.ref_gpu_spmv(which, rows, cols, nnz, P(ho), P(hi), P(hv), P(hx), P(yr), 10, C.byref(ms), C.byref(inner))
            t = inner.value if inner.value > 0 else ms.value
            out["reference_gpu"].append({"kernel": nm, "rc": rc, "ms_best_wrapper": ms.value, "ms_best_timer_t": inner.value,
                                         "gnnz_per_s": nnz / t / 1e6, "y_equal": bool(np.array_equal(yr, y0h))})
            print(f"reference {nm:18s} {t*1e3:9.1f} us  {nnz/t/1e6:7.1f} Gnnz/s  y_equal={np.array_equal(yr, y0h)}", file=log)

    # ---- config 4: BCSR 4x4 bf16 on tcgen05 ----
    nbr = (1 << 12) if small else (1 << 18)
    nb = nbr * 32
    b_off, b_col, _ = g.synth_csr(nbr, nbr, nb, device="cuda")
    e = torch.arange(nb * 16, device="cuda", dtype=torch.int64)
    b_val = (((g._lsr(g.mix64(e ^ 0x5151), 33) % 16) + 1).to(torch.float32) / 8.0).to(torch.bfloat16)
    B = bcsr_t.from_tensors(4, 4, nbr * 4, nbr * 4, nb * 16, b_off, b_col, b_val)
    xb = g.x_recipe(nbr * 4, device="cuda").to(torch.bfloat16)
    yb = torch.full((nbr * 4,), float("nan"), device="cuda")
    spmv.bcsr_thread_mapped(B, xb, yb)
    # independent check with torch ops in float64 (exact inputs)
    rowb = torch.repeat_interleave(torch.arange(nbr, device="cuda"), (b_off[1:] - b_off[:-1]).long(), output_size=nb)
    xs = xb.double()[(b_col.long()[:, None] * 4 + torch.arange(4, device="cuda")[None, :])]      # [nb,4]
    prod = (b_val.double().view(nb, 4, 4) * xs[:, None, :]).sum(2)                                  # [nb,4]
    ref = torch.zeros(nbr, 4, dtype=torch.float64, device="cuda").index_add_(0, rowb, prod).view(-1)
    ok = bool(torch.equal(yb.double(), ref))
    med, best = time_ms(lambda: spmv.bcsr_thread_mapped(B, xb, yb, sync=False))
    nbytes = nb * 32 + nb * 4 + (nbr + 1) * 4 + nbr * 4 * 2 + nbr * 4 * 4
    out["bcsr"] = {"block_rows": nbr, "blocks": nb, "ms_median": med, "ms_min": best, "blocks_per_s": nb / med * 1e3,
                   "algorithmic_bytes": nbytes, "algorithmic_gb_per_s": nbytes / med / 1e6,
                   "roofline_frac": nbytes / med / 1e6 / PEAK, "exact_vs_float64": ok}
    print(f"bcsr4x4 bf16 tcgen05: {med*1e3:.1f} us  {nbytes/med/1e6:.0f} GB/s  frac {nbytes/med/1e6/PEAK:.3f}  ok={ok}", file=log)
    del B, xb, yb, ref
    torch.cuda.empty_cache()

    # ---- SpMM (reference algorithms/spmm/thread_mapped.cuh:28-53): C = A B on the config-2 matrix,
    # B and C dense row-major with n = 32 columns. Last, and guarded: it is a SIMT kernel that carries no
    # roofline claim -- the figure is here so that it is on record, not because it is tuned. ----
    try:
        from loops_b200.algorithms import spmm
        n = 32
        col = torch.arange(cols * n, device="cuda", dtype=torch.int64)
        Bd = ((g._lsr(g.mix64(col ^ 0x7171), 33) % 10) + 1).to(torch.float32).view(cols, n)      # integers 1..10: exact sums
        del col
        Cd = torch.full((rows, n), float("nan"), device="cuda")
        spmm.thread_mapped(A, Bd, Cd)
        # independent check, one dense column at a time through the SpMV path (merge_path_flat/CSR)
        ok = True
        yk = torch.empty(rows, device="cuda")
        for k in (0, n // 2, n - 1):
            spmv.merge_path_flat(A, Bd[:, k].contiguous(), yk, tiled=False)
            ok = ok and bool(torch.equal(yk, Cd[:, k]))
        med, best = time_ms(lambda: spmm.thread_mapped(A, Bd, Cd, sync=False), warm=2, reps=3, batch=5)
        nbytes = nnz * 8 + (rows + 1) * 4 + cols * n * 4 + rows * n * 4
        out["spmm"] = {"n": n, "ms_median": med, "ms_min": best, "gflop_per_s": 2.0 * nnz * n / med / 1e6,
                       "algorithmic_bytes": nbytes, "algorithmic_gb_per_s": nbytes / med / 1e6,
                       "roofline_frac": nbytes / med / 1e6 / PEAK, "columns_equal_to_spmv": ok,
                       "kernel": "sk::spmm_csr_row_warp (SIMT, warp per row x 32 columns; not tuned)"}
        print(f"spmm n={n}: {med*1e3:.1f} us  {nbytes/med/1e6:.0f} GB/s  frac {nbytes/med/1e6/PEAK:.3f}  ok={ok}", file=log)
    except Exception as e:          # never lose the cells above to this one
        out["spmm"] = {"error": repr(e)[:200]}
    return out


def main():
    print(json.dumps(run(small="--small" in sys.argv)))


if __name__ == "__main__":
    main()
