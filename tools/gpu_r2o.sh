#!/bin/bash
# BCSR tcgen05 kernel: alternating work counters (no memset per launch) + programmatic dependent launch; L2-ahead sweep of the tiled kernel
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_bcsr_tc.py -x -q > gpurun_out/pytest_bcsr.log 2>&1; echo "pytest bcsr rc=$?"; tail -3 gpurun_out/pytest_bcsr.log
for P in 1 0; do
  LOOPSB_BCSR_PDL=$P timeout 150 python tools/bcsr_bench.py > gpurun_out/bcsr_bench_pdl$P.log 2>&1; echo "PDL=$P: $(head -1 gpurun_out/bcsr_bench_pdl$P.log)"
done
for A in 4 8 12; do
  LOOPSB_TILED_L2AHEAD=$A timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_ahead$A.json 2> gpurun_out/bench_ahead$A.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_ahead$A.json"))
    print("L2AHEAD=$A: %.2f us/step  frac %.4f  (cold %.2f us)" % (d["ms_per_step"]*1e3, d["roofline"]["frac"], d["roofline"]["cold_l2"]["ms_median"]*1e3))
except Exception as e:
    print("failed", e)
PY
done
