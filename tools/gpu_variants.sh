#!/bin/bash
# tools/gpu_variants.sh -- parity + bench for every merge-kernel geometry variant.
mkdir -p gpurun_out
for v in 0 1 2 3 4 5 6; do
  export LOOPSB_MERGE_VARIANT=$v
  timeout 600 python -m pytest tests/test_gpu_spmv.py -x -q -k "merge or chesapeake or edge or degenerate or exact or overwritten or full_size" > gpurun_out/pytest_v$v.log 2>&1
  echo "variant $v pytest rc=$? $(tail -1 gpurun_out/pytest_v$v.log)"
  timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_v$v.json"))
    print("variant $v: %.1f Gnnz/s  step %.1f us  kernel %.1f us  frac %.3f  e2e %.1f Gnnz/s  grid %d smem %d" % (d["value"]/1e9, d["ms_per_step"]*1e3, d["roofline"]["kernel_ms_mean"]*1e3, d["roofline"]["frac"], d["e2e"]["value"]/1e9, d["plan"]["grid_blocks"], d["plan"]["smem_bytes"]))
except Exception as e:
    print("variant $v bench failed", e); print(open("gpurun_out/bench_v$v.err").read()[-800:])
PY
done
