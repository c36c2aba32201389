#!/bin/bash
# validation pass: smoke, GPU parity tests, driver-style bench, config-3/4 sweep, ncu launch list of the SpMV kernels
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
LOOPSB_TILED=0 timeout 600 python tools/sweep.py > gpurun_out/sweep.json 2> gpurun_out/sweep.log; echo "sweep rc=$?"; tail -25 gpurun_out/sweep.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmv|bcsr|merge" -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
tail -5 gpurun_out/launches.csv
