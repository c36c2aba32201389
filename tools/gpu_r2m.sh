#!/bin/bash
# programmatic dependent launch of the band-tiled kernel (LOOPSB_TILED_PDL=1): parity, then A/B
mkdir -p gpurun_out
LOOPSB_TILED_PDL=1 timeout 150 python -m pytest tests/test_gpu_tiled.py -x -q > gpurun_out/pytest_tiled_pdl.log 2>&1; echo "pytest(PDL) rc=$?"; tail -3 gpurun_out/pytest_tiled_pdl.log
for i in 1 2; do for P in 0 1; do
  LOOPSB_TILED_PDL=$P timeout 120 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_pdl${P}_$i.json 2> gpurun_out/bench_pdl${P}_$i.err
  echo "bench rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_pdl${P}_$i.json"))
    print("PDL=$P run $i: %.2f us/step  frac %.4f  (event-pair %.2f us, cold %.2f us)  chk %.3f  e2e %.1f equal %s" % (d["ms_per_step"]*1e3, d["roofline"]["frac"],
          d["roofline"]["kernel_ms_event_pair_mean"]*1e3, d["roofline"]["cold_l2"]["ms_median"]*1e3, d["y_checksum"], d["e2e"]["value"]/1e9, d["e2e_y_equal_device_y"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_pdl${P}_$i.err").read()[-800:])
PY
done; done
