"""tools/heuristic_sweep.py -- which SpMV schedule wins where on B200.

The reference's paper picks thread-mapped / group-mapped / merge-path per matrix
from (rows, cols, nnz) (plots/data/heuristics.csv: the `kernel` column over 4831
SuiteSparse matrices; the rule itself is not in the tree). This tool measures the
same decision for THIS library's kernels on synthetic power-law matrices over a
grid of sizes, so that `loopsb_select_schedule` can be fitted to measurements:
back-to-back time of each CSR schedule (and of the band-tiled merge-path plan where
its cost model accepts the matrix), best schedule per cell."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from loops_b200 import _lib, csr_t, generate as g
from loops_b200.algorithms import spmv


def b2b_us(fn, reps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def main():
    out = []
    for lr in (8, 10, 12, 14, 16, 18, 20, 22):
        for deg in (2, 8, 32):
            rows = cols = 1 << lr
            nnz = rows * deg
            if nnz > (1 << 27):
                continue
            off, idx, val = g.synth_csr(rows, cols, nnz, device="cuda", d_max=min(1024, cols))
            A = csr_t.from_tensors(rows, cols, off, idx, val)
            x = g.x_recipe(cols, device="cuda")
            y = torch.empty(rows, device="cuda")
            reps = 200 if nnz < (1 << 22) else 40
            cell = {"rows": rows, "nnz": nnz, "avg_degree": deg, "us": {}}
            A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled=False)
            cell["us"]["merge_path_flat"] = b2b_us(lambda: spmv.merge_path_flat(A, x, y, sync=False, tiled=False), reps)
            for name in ("thread_mapped", "group_mapped", "work_oriented"):
                fn = spmv.BY_NAME[name]
                cell["us"][name] = b2b_us(lambda: fn(A, x, y, sync=False), reps)
            A.drop_plans()
            p = A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled="auto")
            if p.tiled_info():
                cell["us"]["merge_path_flat+tiled_plan"] = b2b_us(
                    lambda: spmv.merge_path_flat(A, x, y, sync=False, tiled="auto"), reps)
            cell["best"] = min(cell["us"], key=cell["us"].get)
            out.append(cell)
            print(f"rows 2^{lr:<2d} deg {deg:<3d} nnz {nnz:>10d}: " +
                  "  ".join(f"{k} {v:8.1f}" for k, v in cell["us"].items()) + f"   -> {cell['best']}", flush=True)
            A.drop_plans()
            del A, off, idx, val, x, y
    print(json.dumps(out))


if __name__ == "__main__":
    main()
