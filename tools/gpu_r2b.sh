#!/bin/bash
# round 2, call B: gen-2b merge kernel parity + timing + wavefront counters
mkdir -p gpurun_out
echo "== pytest merge2 + spmv"
timeout 900 python -m pytest tests/test_gpu_merge2.py tests/test_gpu_spmv.py -x -q > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_b.log
echo "== probe"
PROBE_COMBOS="2:0 2:1 2:2 2:3 2:4 2:5 2:6" timeout 600 python tools/merge_probe.py cfg2 shard8 2>&1 | tee gpurun_out/merge_probe_b.txt
echo "== ncu counters gen2"
LOOPSB_TILED=0 timeout 600 ncu --clock-control none -k regex:spmv_merge2_kernel -s 4 -c 1 \
  --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "spmv_merge2|gpu__time|smsp__|l1tex|sm__cycles" | tee gpurun_out/ncu_counters_b.txt
