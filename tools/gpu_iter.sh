#!/bin/bash
# tools/gpu_iter.sh -- quick iteration loop: parity on the merge kernel, phase
# breakdown and bench for the variants given as arguments (default "0").
mkdir -p gpurun_out
VARS="${@:-0}"
timeout 600 python -m pytest tests/test_gpu_spmv.py -x -q > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_iter.log)"
grep -E "FAILED|Error|assert" gpurun_out/pytest_iter.log | head -10
timeout 300 python tools/phase_probe.py $VARS
for v in $VARS; do
  LOOPSB_MERGE_VARIANT=$v timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_v$v.json"))
    print("variant $v: %.1f Gnnz/s  step %.1f us  kernel %.1f us  frac %.3f  e2e %.1f Gnnz/s" % (d["value"]/1e9, d["ms_per_step"]*1e3, d["roofline"]["kernel_ms_mean"]*1e3, d["roofline"]["frac"], d["e2e"]["value"]/1e9))
except Exception as e:
    print("variant $v bench failed", e); print(open("gpurun_out/bench_v$v.err").read()[-800:])
PY
done
