#!/bin/bash
# PDL as the default (with the multi-stream tracker): parity incl. the two-stream test, tracking overhead, full bench line
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_tiled.py tests/test_gpu_cpp_dropin.py -x -q > gpurun_out/pytest_tiled.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_tiled.log
for P in 1 2 0 1; do
  LOOPSB_TILED_PDL=$P timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_pdl${P}.json 2> gpurun_out/bench_pdl${P}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_pdl${P}.json"))
    print("PDL=$P: %.2f us/step  frac %.4f  (event-pair %.2f us, cold %.2f us)  chk %.3f  e2e %.1f serial %.1f equal %s" % (d["ms_per_step"]*1e3, d["roofline"]["frac"],
          d["roofline"]["kernel_ms_event_pair_mean"]*1e3, d["roofline"]["cold_l2"]["ms_median"]*1e3, d["y_checksum"], d["e2e"]["value"]/1e9, d["e2e"]["serial_value"]/1e9, d["e2e_y_equal_device_y"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_pdl${P}.err").read()[-800:])
PY
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_n1.json
