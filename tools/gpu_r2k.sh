#!/bin/bash
# (historic) L2-resident head of the band-tiled plan's matrix copy: the LOOPSB_TILED_PIN_MB experiment of round 2 -- measured (profiles/tiled_ab_r02.txt), dropped, and the env knob no longer exists in the library
mkdir -p gpurun_out
LOOPSB_TILED_PIN_MB=64 timeout 300 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_tiled_build.py -x -q > gpurun_out/pytest_tiled.log 2>&1; echo "pytest(pin on) rc=$?"; tail -3 gpurun_out/pytest_tiled.log
for MB in ${PIN_LIST:-0 24 40 56 72 96 0}; do
  LOOPSB_TILED_PIN_MB=$MB timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_pin$MB.json 2> gpurun_out/bench_pin$MB.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_pin$MB.json"))
    print("PIN_MB=$MB: %.2f us/step  frac %.4f  (event-pair %.2f us, cold %.2f us)  y_checksum %.3f" % (d["ms_per_step"]*1e3, d["roofline"]["frac"],
          d["roofline"]["kernel_ms_event_pair_mean"]*1e3, d["roofline"]["cold_l2"]["ms_median"]*1e3, d["y_checksum"]))
except Exception as e:
    print("PIN_MB=$MB failed", e); print(open("gpurun_out/bench_pin$MB.err").read()[-800:])
PY
done
