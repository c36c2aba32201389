"""tools/bcsr_bench.py -- BASELINE config 4 alone: BCSR 4x4 bf16 on tcgen05,
262,144 block-rows / 8,388,608 blocks; exact check against float64 torch ops,
CUDA-event timing (isolated launches and back to back)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from loops_b200 import generate as g
from loops_b200.algorithms import spmv
from loops_b200.container import bcsr_t

PEAK = 6548.2
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
small = "--small" in sys.argv
nbr = (1 << 12) if small else (1 << 18)
nb = nbr * 32
b_off, b_col, _ = g.synth_csr(nbr, nbr, nb, device="cuda")
e = torch.arange(nb * 16, device="cuda", dtype=torch.int64)
b_val = (((g._lsr(g.mix64(e ^ 0x5151), 33) % 16) + 1).to(torch.float32) / 8.0).to(torch.bfloat16)
B = bcsr_t.from_tensors(4, 4, nbr * 4, nbr * 4, nb * 16, b_off, b_col, b_val)
xb = g.x_recipe(nbr * 4, device="cuda").to(torch.bfloat16)
yb = torch.full((nbr * 4,), float("nan"), device="cuda")
spmv.bcsr_thread_mapped(B, xb, yb)
rowb = torch.repeat_interleave(torch.arange(nbr, device="cuda"), (b_off[1:] - b_off[:-1]).long(), output_size=nb)
xs = xb.double()[(b_col.long()[:, None] * 4 + torch.arange(4, device="cuda")[None, :])]
prod = (b_val.double().view(nb, 4, 4) * xs[:, None, :]).sum(2)
ref = torch.zeros(nbr, 4, dtype=torch.float64, device="cuda").index_add_(0, rowb, prod).view(-1)
ok = bool(torch.equal(yb.double(), ref))
fn = lambda: spmv.bcsr_thread_mapped(B, xb, yb, sync=False)
for _ in range(5):
    fn()
torch.cuda.synchronize()
ts = []
for _ in range(30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
ts.sort()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(100):
    fn()
b.record(); b.synchronize()
b2b = a.elapsed_time(b) / 100
ok2 = bool(torch.equal(yb.double(), ref))
nbytes = nb * 32 + nb * 4 + (nbr + 1) * 4 + nbr * 4 * 2 + nbr * 4 * 4
print(f"bcsr4x4 bf16 tcgen05: median {ts[15]*1e3:.1f} us, min {ts[0]*1e3:.1f}, back-to-back {b2b*1e3:.1f} us = "
      f"{nbytes/b2b/1e6:.0f} GB/s ({nbytes/b2b/1e6/PEAK:.3f} of peak)  exact={ok and ok2}")
print(json.dumps({"block_rows": nbr, "blocks": nb, "ms_median": ts[15], "ms_min": ts[0], "ms_back_to_back": b2b,
                  "algorithmic_bytes": nbytes, "roofline_frac_b2b": nbytes / b2b / 1e6 / PEAK,
                  "roofline_frac_median": nbytes / ts[15] / 1e6 / PEAK, "exact_vs_float64": ok and ok2}))
