"""tools/merge_probe.py [cfg2|shard8|shard4|shard2 ...] -- kernel time of the plain CSR
merge-path kernels (no band-tiled copy) for every (generation, variant) pair, on
BASELINE configs[1] (2^20 rows / 2^25 nnz) and on ONE rank's shard of configs[4]
(2^24 / 2^29 cut into N row ranges, global column ids, x = 64 MB). One process;
LOOPSB_MERGE_KERNEL / LOOPSB_MERGE2_VARIANT / LOOPSB_MERGE_VARIANT are read at plan
creation. Prints y checksums so variants can be compared with each other."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loops_b200 import csr_t, generate as g
from loops_b200.algorithms import spmv

COMBOS = os.environ.get("PROBE_COMBOS", "1:9 2:0 2:1 2:2 2:3 2:4 2:5 2:6").split()
REPS = int(os.environ.get("PROBE_REPS", "50"))


def time_one(A, x, y):
    for _ in range(5):
        spmv.merge_path_flat(A, x, y, sync=False, tiled=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REPS):
        spmv.merge_path_flat(A, x, y, sync=False, tiled=False)
    b.record(); b.synchronize()
    return a.elapsed_time(b) / REPS


def run(label, rows_total, cols, nnz_total, r1):
    deg = g.powerlaw_degrees(rows_total, nnz_total, d_max=1024)
    off, idx, val = g.synth_csr(rows_total, cols, nnz_total, device="cuda", degrees=deg, row_begin=0, row_end=r1)
    x = g.x_recipe(cols, device="cuda")
    y = torch.empty(r1, device="cuda")
    lnnz = int(idx.numel())
    floor_us = lnnz / 148 / 1.965e3
    ref = None
    for combo in COMBOS:
        gen, var = combo.split(":")
        os.environ["LOOPSB_MERGE_KERNEL"] = gen
        os.environ["LOOPSB_MERGE_VARIANT" if gen == "1" else "LOOPSB_MERGE2_VARIANT"] = var
        A = csr_t.from_tensors(r1, cols, off, idx, val)
        y.fill_(float("nan"))
        ms = time_one(A, x, y)
        chk = float(y.double().sum().item())
        if ref is None:
            ref = y.clone()
        same = bool(torch.equal(ref, y))
        byts = lnnz * 8 + (r1 + 1) * 4 + cols * 4 + r1 * 4
        print(f"{label}: gen {gen} variant {var}: {ms*1e3:7.1f} us  {lnnz/ms/1e6:7.1f} Gnnz/s  "
              f"{byts/ms/1e6:7.0f} GB/s  (gather floor {floor_us:.0f} us -> {floor_us/(ms*1e3):.2f})  "
              f"y==first {same}  chk {chk:.6e}", flush=True)
        A.drop_plans()


for what in (sys.argv[1:] or ["cfg2", "shard8"]):
    if what == "cfg2":
        run("cfg2", 1 << 20, 1 << 20, 1 << 25, 1 << 20)
    elif what.startswith("shard"):
        n = int(what[5:])
        run(f"shard 1/{n}", 1 << 24, 1 << 24, 1 << 29, (1 << 24) // n)
