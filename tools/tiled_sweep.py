"""tools/tiled_sweep.py -- band-tiled merge_path_flat on BASELINE config 2 over a
list of geometries ("nb,q,warps,cb,xb,es" each; default list below): builds the
plan, checks y bit for bit against the plain CSR merge-path kernel (exact
inputs), times it with CUDA events. Prints one line per geometry and a JSON
document at the end."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from loops_b200 import _lib, csr_t, generate as g
from loops_b200.algorithms import spmv

DEFAULT = ["0,0,0,0,0,0", "37,4,16,4096,4,3", "74,2,16,12288,2,3", "148,1,16,16384,2,3", "37,4,12,8192,2,4",
           "37,4,8,8192,2,6", "37,4,16,5440,3,3"]


def time_ms(fn, warm=5, reps=40):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def time_b2b_ms(fn, warm=5, reps=100):
    """Back-to-back launches between one event pair (what bench.py's `value` sees)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / reps


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    dmax = 1024
    for a in sys.argv[1:]:
        if a.startswith("--dmax="):
            dmax = int(a.split("=")[1])   # heaviest row of the power law (hub rows exercise the long-run path)
    small = "--small" in sys.argv
    rows = cols = (1 << 16) if small else (1 << 20)
    nnz = rows * 32
    geoms = args or DEFAULT
    peak = 6548.2
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    off, idx, val = g.synth_csr(rows, cols, nnz, device="cuda", d_max=dmax)
    x = g.x_recipe(cols, device="cuda")
    A0 = csr_t.from_tensors(rows, cols, off, idx, val)
    y0 = torch.empty(rows, device="cuda")
    spmv.merge_path_flat(A0, x, y0, tiled=False)
    med0, min0 = time_ms(lambda: spmv.merge_path_flat(A0, x, y0, sync=False, tiled=False))
    nbytes = nnz * 8 + (rows + 1) * 4 + cols * 4 + rows * 4
    print(f"plain csr merge-path: {med0*1e3:.1f} us median, {min0*1e3:.1f} min, {nbytes/med0/1e6:.0f} GB/s", flush=True)
    out = {"rows": rows, "nnz": nnz, "peak_gbs": peak, "plain_ms": med0, "cells": []}
    for geo in geoms:
        os.environ["LOOPSB_TILED_GEOM"] = geo
        A = csr_t.from_tensors(rows, cols, off, idx, val)
        t0 = time.time()
        try:
            plan = A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled=True)
        except Exception as e:
            print(f"{geo}: plan failed: {e}", flush=True)
            continue
        info = plan.tiled_info()
        if info is None:
            print(f"{geo}: declined: {getattr(plan, 'tile_declined', '?')}", flush=True)
            continue
        build_s = time.time() - t0
        y = torch.full((rows,), float("nan"), device="cuda")
        spmv.merge_path_flat(A, x, y, tiled=True)
        ok = bool(torch.equal(y, y0))
        med, best = time_ms(lambda: spmv.merge_path_flat(A, x, y, sync=False, tiled=True))
        b2b = time_b2b_ms(lambda: spmv.merge_path_flat(A, x, y, sync=False, tiled=True))
        ok2 = bool(torch.equal(y, y0))
        gbs = nbytes / med / 1e6
        print(f"{geo}: {med*1e3:.1f} us median, {best*1e3:.1f} min, back-to-back {b2b*1e3:.1f} us ({nbytes/b2b/1e6/peak:.3f} of peak), "
              f"{gbs:.0f} GB/s ({gbs/peak:.3f} of peak) ok={ok and ok2} "
              f"geom=({info['nb']},{info['q']},{info['warps']},{info['cb']},{info['xb']},{info['es']}) smem={info['smem_bytes']} "
              f"pad={info['pad_entries']/info['real_entries']:.4f} flagged_steps={info['flagged_steps']/max(info['total_steps'],1):.3f} "
              f"build={build_s:.1f}s", flush=True)
        out["cells"].append({"geometry": geo, "ms_median": med, "ms_min": best, "ms_back_to_back": b2b, "gbs": gbs, "frac": gbs / peak,
                             "ok": ok and ok2, "info": info, "build_s": build_s})
        A.drop_plans()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
