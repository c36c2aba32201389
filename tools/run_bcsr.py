"""tools/run_bcsr.py -- a few launches of the BCSR 4x4 bf16 tcgen05 kernel at
BASELINE config 4 size (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loops_b200 import generate as g
from loops_b200.algorithms import spmv
from loops_b200.container import bcsr_t
nbr = 1 << 18; nb = nbr * 32
b_off, b_col, _ = g.synth_csr(nbr, nbr, nb, device="cuda")
e = torch.arange(nb * 16, device="cuda", dtype=torch.int64)
b_val = (((g._lsr(g.mix64(e ^ 0x5151), 33) % 16) + 1).to(torch.float32) / 8.0).to(torch.bfloat16)
B = bcsr_t.from_tensors(4, 4, nbr * 4, nbr * 4, nb * 16, b_off, b_col, b_val)
xb = g.x_recipe(nbr * 4, device="cuda").to(torch.bfloat16)
yb = torch.empty(nbr * 4, device="cuda")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 6):
    spmv.bcsr_thread_mapped(B, xb, yb)
print("checksum", float(yb.double().sum()))
