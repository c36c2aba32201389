#!/bin/bash
# tools/gpu_check.sh -- one gpurun call: smoke, GPU parity tests, golden stream
# capture from the reference templates, a short bench, and the ncu launch list.
# Usage (from the repo root, on the GPU box):  bash tools/gpu_check.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" | tee gpurun_out/smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
if [ "$1" != "quick" ]; then
  echo "== golden streams from the reference templates"
  timeout 600 python tests/golden/make_golden.py gpu gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?" | tee -a gpurun_out/golden.log
  tail -3 gpurun_out/golden.log
fi
echo "== bench"
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -12 gpurun_out/launches.csv
