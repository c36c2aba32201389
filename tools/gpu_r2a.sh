#!/bin/bash
# round 2, call A: gen-2 merge kernel parity + variant timing + ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest merge2 + spmv"
timeout 900 python -m pytest tests/test_gpu_merge2.py tests/test_gpu_spmv.py -x -q > gpurun_out/pytest_a.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_a.log
echo "== probe"
timeout 600 python tools/merge_probe.py cfg2 shard8 2>&1 | tee gpurun_out/merge_probe.txt
echo "== ncu gen2"
LOOPSB_TILED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_merge2_kernel -s 4 -c 1 -f -o gpurun_out/prof_merge2 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
