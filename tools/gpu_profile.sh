#!/bin/bash
# tools/gpu_profile.sh -- the evidence pass of a round: full bench (both arms),
# ncu launch list of the bench command, full ncu captures of the two hot kernels.
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmv_|merge_|bcsr" -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_bt_kernel -s 4 -c 2 -f -o gpurun_out/prof_tiled \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_tiled.log 2>&1; echo "ncu tiled rc=$?"
LOOPSB_TILED=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_merge_kernel -s 4 -c 2 -f -o gpurun_out/prof_merge \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu merge rc=$?"
LOOPSB_TILED=0 timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_plain.json 2> gpurun_out/bench_n1_plain.err; echo "bench plain rc=$?"; cut -c1-300 gpurun_out/bench_n1_plain.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bcsr4x4 -s 2 -c 2 -f -o gpurun_out/prof_bcsr \
    python tools/run_bcsr.py 5 > gpurun_out/ncu_bcsr.log 2>&1; echo "ncu bcsr rc=$?"
timeout 300 python tools/bcsr_bench.py > gpurun_out/bcsr_bench.log 2>&1; head -1 gpurun_out/bcsr_bench.log
timeout 600 python tools/tiled_sweep.py 0,0,0,0,0,0 37,4,24,6144,4,3 37,4,20,7168,3,3 37,4,16,4096,4,3 > gpurun_out/tiled_sweep.log 2>&1; grep -v "^{" gpurun_out/tiled_sweep.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
