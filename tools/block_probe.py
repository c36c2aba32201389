"""tools/block_probe.py N "g1,g2,..." -- ONE rank's shard of configs[4] (2^24 / 2^29 over N ranks)
on a single GPU, cut into the column blocks loopsb_dist_create would hold for rank 0 with the
given phase groups; times the unsplit SpMV and every block's SpMV (y = A_0 x, then y += A_b x)
without any communication. Shows what the split itself costs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loops_b200 import _lib, csr_t, generate as g
from loops_b200.algorithms import spmv
from loops_b200.convert import csr_split_columns
from loops_b200.dist import ring_blocks

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
groups = [int(t) for t in (sys.argv[2] if len(sys.argv) > 2 else "2,2,3").split(",")]
REPS = int(os.environ.get("PROBE_REPS", "20"))
rows, cols, nnz = 1 << 24, 1 << 24, 1 << 29
deg = g.powerlaw_degrees(rows, nnz, d_max=1024)
r1 = rows // N
off, idx, val = g.synth_csr(rows, cols, nnz, device="cuda", degrees=deg, row_begin=0, row_end=r1)
x = g.x_recipe(cols, device="cuda")
y = torch.empty(r1, device="cuda")
A = csr_t.from_tensors(r1, cols, off, idx, val)
lib = _lib.load()


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REPS):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / REPS * 1e3


t_full = timed(lambda: spmv.merge_path_flat(A, x, y, sync=False, tiled=False))
y_full = y.clone()
print(f"shard 1/{N}: {r1} rows, {A.nnzs} nnz unsplit: {t_full:.1f} us  {A.nnzs / t_full / 1e3:.1f} Gnnz/s", flush=True)
blocks = csr_split_columns(A, cols // N, ring_blocks(N, 0, groups))
plans = [B.plan(_lib.SCHED_MERGE_PATH_FLAT, None, tiled=False) for B in blocks]
S = _lib.stream_ptr(None)


NOACC = os.environ.get("PROBE_NOACC", "0") == "1"     # every block writes its own y_b (no read-modify-write)
ys = [torch.empty(r1, device="cuda") for _ in blocks] if NOACC else None


def run_block(b):
    B = blocks[b]
    if NOACC:
        _lib.check(lib.loopsb_spmv_f32(plans[b].handle, _lib.ptr(B.values), _lib.ptr(B.indices), None, _lib.ptr(x),
                                       _lib.ptr(ys[b]), r1, cols, S), "spmv")
        return
    if b == 0:
        _lib.check(lib.loopsb_spmv_f32(plans[b].handle, _lib.ptr(B.values), _lib.ptr(B.indices), None, _lib.ptr(x),
                                       _lib.ptr(y), r1, cols, S), "spmv")
    else:
        _lib.check(lib.loopsb_spmv_acc_f32(plans[b].handle, _lib.ptr(B.values), _lib.ptr(B.indices), _lib.ptr(x),
                                           _lib.ptr(y), r1, cols, S), "acc")


total = 0.0
for b in range(len(blocks)):
    t = timed(lambda: run_block(b))
    total += t
    print(f"  block {b}: {blocks[b].nnzs} nnz: {t:.1f} us  {blocks[b].nnzs / t / 1e3:.1f} Gnnz/s", flush=True)
t_all = timed(lambda: [run_block(b) for b in range(len(blocks))])
for b in range(len(blocks)):
    run_block(b)
torch.cuda.synchronize()
if NOACC:
    t_sum = timed(lambda: torch.sum(torch.stack(ys), 0, out=y))
    print(f"  combine of {len(ys)} partial y: {t_sum:.1f} us", flush=True)
print(f"  all blocks back to back: {t_all:.1f} us (sum of singles {total:.1f}); y == unsplit: {bool(torch.equal(y, y_full))}", flush=True)
