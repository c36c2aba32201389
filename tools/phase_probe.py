"""tools/phase_probe.py -- per-phase cycle breakdown of the merge-path kernel
(LOOPSB_DEBUG_PHASES counters) on the config-2 workload, for each geometry
variant given on the command line."""
import os, sys, json
os.environ["LOOPSB_DEBUG_PHASES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from loops_b200 import _lib, csr_t, generate as g
from loops_b200.algorithms import spmv
rows = cols = 1 << 20; nnz = 1 << 25
off, idx, val = g.synth_csr(rows, cols, nnz, device="cuda")
x = g.x_recipe(cols, device="cuda"); y = torch.empty(rows, device="cuda")
names = ["wait", "gather+search", "walk", "scan+refill", "combine+store", "-"]
for v in sys.argv[1:] or ["0"]:
    os.environ["LOOPSB_MERGE_VARIANT"] = v
    A = csr_t.from_tensors(rows, cols, off, idx, val)
    for _ in range(5): spmv.merge_path_flat(A, x, y)
    plan = A.plan(_lib.SCHED_MERGE_PATH_FLAT); info = plan.info()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); spmv.merge_path_flat(A, x, y, sync=False); e1.record(); torch.cuda.synchronize()
    out = np.zeros((info.grid_blocks, 8), np.int64)
    _lib.check(_lib.load().loopsb_plan_debug_phases_host(plan.handle, out.ctypes.data, info.grid_blocks), "phases")
    tiles = out[:, 6].mean(); tot = out[:, :6].sum(1)
    print(f"variant {v}: grid {info.grid_blocks} x{info.cta_threads}, {tiles:.1f} tiles/CTA, step {e0.elapsed_time(e1)*1e3:.1f} us, "
          f"CTA total cycles mean {tot.mean():.0f} max {tot.max()}")
    print("   per-tile cycles: " + "  ".join(f"{n} {out[:, i].mean()/tiles:7.0f}" for i, n in enumerate(names)))
    A.drop_plans()
