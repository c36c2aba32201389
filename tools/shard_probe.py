"""tools/shard_probe.py [N] -- the local SpMV of ONE rank of the multi-GPU run
(BASELINE configs[4]: 2^24 rows / 2^29 nnz cut into N row shards, global column
ids, x = 64 MB) on a single GPU: what the per-shard kernel costs without the
all-gather. Prints the kernel time for the variants in LOOPSB_MERGE_VARIANT."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loops_b200 import _lib, csr_t, generate as g
from loops_b200.algorithms import spmv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rows, cols, nnz = 1 << 24, 1 << 24, 1 << 29
deg = g.powerlaw_degrees(rows, nnz, d_max=1024)
r1 = rows // N
off, idx, val = g.synth_csr(rows, cols, nnz, device="cuda", degrees=deg, row_begin=0, row_end=r1)
x = g.x_recipe(cols, device="cuda")
y = torch.empty(r1, device="cuda")
lnnz = int(idx.numel())
for variant in (sys.argv[2:] or ["9"]):
    os.environ["LOOPSB_MERGE_VARIANT"] = variant
    A = csr_t.from_tensors(r1, cols, off, idx, val)
    for _ in range(5):
        spmv.merge_path_flat(A, x, y, sync=False, tiled=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        spmv.merge_path_flat(A, x, y, sync=False, tiled=False)
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 50
    print(f"shard 1/{N}: {r1} rows, {lnnz} nnz, variant {variant}: {ms*1e3:.1f} us  {lnnz/ms/1e6:.1f} Gnnz/s "
          f"(gather-rate bound {lnnz/148/1.965e3:.0f} us)", flush=True)
    A.drop_plans()
