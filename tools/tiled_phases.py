"""tools/tiled_phases.py -- per-phase cycle breakdown of the band-tiled kernel
(LOOPSB_DEBUG_PHASES counters: per consumer warp {stream wait, band wait,
gather, y update fast, y update flagged, flagged steps, total, steps}) on the
config-2 workload for the geometries given on the command line."""
import os, sys
os.environ["LOOPSB_DEBUG_PHASES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from loops_b200 import _lib, csr_t, generate as g
from loops_b200.algorithms import spmv
rows = cols = 1 << 20; nnz = 1 << 25
off, idx, val = g.synth_csr(rows, cols, nnz, device="cuda")
x = g.x_recipe(cols, device="cuda"); y = torch.empty(rows, device="cuda")
for geo in sys.argv[1:] or ["0,0,0,0,0,0"]:
    os.environ["LOOPSB_TILED_GEOM"] = geo
    A = csr_t.from_tensors(rows, cols, off, idx, val)
    plan = A.plan(_lib.SCHED_MERGE_PATH_FLAT, tiled=True)
    info = plan.tiled_info()
    for _ in range(5): spmv.merge_path_flat(A, x, y, tiled=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); spmv.merge_path_flat(A, x, y, sync=False, tiled=True); e1.record(); torch.cuda.synchronize()
    ns = info["nb"] * info["q"] * info["warps"]
    grid = info["nb"] * info["q"]
    raw = np.zeros(ns * 8 + ((grid * 4 + 7) // 8) * 8, np.int64)
    _lib.check(_lib.load().loopsb_plan_debug_phases_host(plan.handle, raw.ctypes.data, raw.size // 8), "phases")
    out = raw[: ns * 8].reshape(ns, 8)
    st = raw[ns * 8: ns * 8 + grid * 4].reshape(grid, 4).astype(float)
    t0 = st[:, 0].min()
    print(f"   wall clock (us from first CTA entry): entry max {(st[:,0].max()-t0)/1e3:.1f}; init done mean {(st[:,1]-t0).mean()/1e3:.1f} max {(st[:,1]-t0).max()/1e3:.1f}; "
          f"loop done mean {(st[:,2]-t0).mean()/1e3:.1f} min {(st[:,2]-t0).min()/1e3:.1f} max {(st[:,2]-t0).max()/1e3:.1f}; CTA end mean {(st[:,3]-t0).mean()/1e3:.1f} max {(st[:,3]-t0).max()/1e3:.1f}")
    if os.environ.get("STAMPS_ONLY"):   # library built with -DLOOPSB_STAMPS_ONLY=1: wall-clock stamps only
        W = info["warps"]; o = out.astype(float)
        first, wend = o[:, 0] - t0, o[:, 1] - t0
        pub, vis = o[::W, 3] - t0, o[::W, 2] - t0
        print(f"{geo}: launch {e0.elapsed_time(e1)*1e3:.1f} us; per warp: first step ready mean {first.mean()/1e3:.2f} max {first.max()/1e3:.2f}; "
              f"stream done mean {wend.mean()/1e3:.2f} min {wend.min()/1e3:.2f} max {wend.max()/1e3:.2f}")
        print(f"   per CTA: partials stored mean {pub.mean()/1e3:.2f} max {pub.max()/1e3:.2f}; peers visible mean {vis.mean()/1e3:.2f} max {vis.max()/1e3:.2f}; "
              f"end mean {(st[:,3]-t0).mean()/1e3:.2f} max {(st[:,3]-t0).max()/1e3:.2f}", flush=True)
        A.drop_plans()
        continue
    steps = out[:, 7].astype(float); m = steps > 0
    print(f"{geo}: launch {e0.elapsed_time(e1)*1e3:.1f} us (profiled build); warps {ns}, steps/warp {steps[m].mean():.1f}, "
          f"warp total cycles mean {out[m,6].mean():.0f} max {out[:,6].max()}")
    names = ["stream wait", "band wait", "gather", "y fast", "y flagged"]
    print("   per-step cycles: " + "  ".join(f"{n} {out[m,i].sum()/steps[m].sum():7.0f}" for i, n in enumerate(names)) +
          f"  | flagged steps {out[m,5].sum()/steps[m].sum():.3f}, cycles per flagged step {out[m,4].sum()/max(out[m,5].sum(),1):.0f}", flush=True)
    A.drop_plans()
    W = info["warps"]
    for cta in (0, 77):
        blk = out[cta * W:(cta + 1) * W]
        print(f"   cta {cta}: steps  " + " ".join(f"{int(v):5d}" for v in blk[:, 7]))
        print(f"   cta {cta}: total  " + " ".join(f"{int(v/1000):5d}" for v in blk[:, 6]) + "  (k cycles)")
        print(f"   cta {cta}: bandw  " + " ".join(f"{int(v/1000):5d}" for v in blk[:, 1]))
        print(f"   cta {cta}: strmw  " + " ".join(f"{int(v/1000):5d}" for v in blk[:, 0]))
    tot = out[:, 6].reshape(-1, W)
    print(f"   per-CTA slowest warp cycles: mean {tot.max(1).mean():.0f} max {tot.max(1).max()} min {tot.max(1).min()}; steps/CTA-max-warp mean {out[:,7].reshape(-1,W).max(1).mean():.1f}")
