"""tools/bcsr_uniform.py -- BCSR 4x4 bf16 tcgen05 on matrices whose block-rows all have the
same length (8 / 32 / 64 blocks): separates the per-work-item cost from the per-K-step cost
(DESIGN 4.3). Back-to-back time of the packed and the direct kernel."""
import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from loops_b200 import generate as g
from loops_b200.algorithms import spmv
from loops_b200.container import bcsr_t
nbr = 1 << 18
for label, deg in (("uniform 32 blocks/row", np.full(nbr, 32, np.int64)), ("uniform 64", np.full(nbr, 64, np.int64)), ("uniform 8", np.full(nbr, 8, np.int64))):
    nb = int(deg.sum())
    b_off, b_col, _ = g.synth_csr(nbr, nbr, nb, device="cuda", degrees=deg)
    e = torch.arange(nb * 16, device="cuda", dtype=torch.int64)
    b_val = (((g._lsr(g.mix64(e ^ 0x5151), 33) % 16) + 1).to(torch.float32) / 8.0).to(torch.bfloat16)
    B = bcsr_t.from_tensors(4, 4, nbr * 4, nbr * 4, nb * 16, b_off, b_col, b_val)
    xb = g.x_recipe(nbr * 4, device="cuda").to(torch.bfloat16)
    yb = torch.empty(nbr * 4, device="cuda")
    for mode in ("1", "0"):
        os.environ["LOOPSB_BCSR_PACKED"] = mode
        B.drop_plans()
        for _ in range(5): spmv.bcsr_thread_mapped(B, xb, yb, sync=False)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50): spmv.bcsr_thread_mapped(B, xb, yb, sync=False)
        b.record(); b.synchronize()
        us = a.elapsed_time(b) / 50 * 1e3
        nbytes = nb * 36 + nbr * 4 + nbr * 4 * 6
        print(f"{label:24s} packed={mode}: {us:7.1f} us  {nbytes/us/1e3:7.0f} GB/s  ({nb} blocks, {us*1e3/ (nb/128):.1f} ns per 128-block step per GPU)")
