#!/bin/bash
# final evidence pass of round 2: smoke, all GPU tests, the driver's bench command, ncu launch list + one full capture of the tiled kernel
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench_n1.json; tail -16 gpurun_out/bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmv|bcsr|merge" -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spmv_bt_kernel -s 6 -c 1 -f -o gpurun_out/prof_tiled_r02 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/ncu_full_tiled.log 2>&1; echo "ncu tiled rc=$?"
