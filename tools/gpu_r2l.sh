#!/bin/bash
# A/B of the consumer-warp row split of the band-tiled plan (LOOPSB_TILED_SPLIT=0 = round-1 rule) + parity
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_tiled_build.py -x -q > gpurun_out/pytest_tiled.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tiled.log
for i in 1 2; do for S in 0 1; do
  LOOPSB_TILED_SPLIT=$S timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_split${S}_$i.json 2> gpurun_out/bench_split${S}_$i.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_split${S}_$i.json"))
    print("SPLIT=$S run $i: %.2f us/step  frac %.4f  (event-pair %.2f us, cold %.2f us)  steps %d" % (d["ms_per_step"]*1e3, d["roofline"]["frac"],
          d["roofline"]["kernel_ms_event_pair_mean"]*1e3, d["roofline"]["cold_l2"]["ms_median"]*1e3, d["plan"]["band_tiled"]["total_steps"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_split${S}_$i.err").read()[-800:])
PY
done; done
