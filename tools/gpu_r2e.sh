#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, "persisting max", getattr(p, "persisting_l2_cache_max_size", None), "window max", getattr(p, "access_policy_max_window_size", None))
PY
python tools/block_probe.py 2 1 2>&1 | tee gpurun_out/block_probe_n2.txt
LOOPSB_NO_L2_PIN=1 python tools/block_probe.py 2 1 2>&1 | sed 's/^/nopin /'
python tools/block_probe.py 8 2,2,3 2>&1 | tee gpurun_out/block_probe_n8.txt
LOOPSB_NO_L2_PIN=1 python tools/block_probe.py 8 2,2,3 2>&1 | sed 's/^/nopin /'
python tools/block_probe.py 8 1,1,1,1,1,1,1 2>&1 | tee gpurun_out/block_probe_n8_7.txt
