#!/bin/bash
# finer LOOPSB_TILED_L2AHEAD sweep of the band-tiled kernel (profiles/tiled_l2ahead_r02.txt)
mkdir -p gpurun_out
for A in 6 2 3 4 6 3; do
  LOOPSB_TILED_L2AHEAD=$A timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_ahead$A.json 2> gpurun_out/bench_ahead$A.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_ahead$A.json"))
    print("L2AHEAD=$A: %.2f us/step  frac %.4f  (event-pair %.2f, cold %.2f us)" % (d["ms_per_step"]*1e3, d["roofline"]["frac"], d["roofline"]["kernel_ms_event_pair_mean"]*1e3, d["roofline"]["cold_l2"]["ms_median"]*1e3))
except Exception as e:
    print("failed", e)
PY
done
