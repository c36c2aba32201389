#!/bin/bash
# A/B of the un-rounded stream lengths of the band-tiled plan (LOOPSB_TILED_ROUND=1 = round-1 format) + its parity tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_tiled_build.py -x -q > gpurun_out/pytest_tiled.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_tiled.log
for i in 1 2 3; do for R in 1 0; do
  LOOPSB_TILED_ROUND=$R timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_round${R}_$i.json 2> gpurun_out/bench_round${R}_$i.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_round${R}_$i.json"))
print("ROUND=$R run $i: %.2f us/step  frac %.4f  (event-pair %.2f us, cold %.2f us)  steps %d  e2e %.1f" % (d["ms_per_step"]*1e3, d["roofline"]["frac"],
      d["roofline"]["kernel_ms_event_pair_mean"]*1e3, d["roofline"]["cold_l2"]["ms_median"]*1e3, d["plan"]["band_tiled"]["total_steps"], d["e2e"]["value"]/1e9))
PY
done; done
