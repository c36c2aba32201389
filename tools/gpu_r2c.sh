#!/bin/bash
# round 2, call C: ncu of the shard kernel (x = 64 MB) -- where do its bytes come from
mkdir -p gpurun_out
PROBE_COMBOS="2:3" PROBE_REPS=3 timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_merge2_kernel -s 6 -c 1 -f -o gpurun_out/prof_shard8_r02 \
   python tools/merge_probe.py shard8 > gpurun_out/ncu_shard.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_shard.log
LOOPSB_NO_L2_PIN=1 PROBE_COMBOS="2:3 2:0" timeout 300 python tools/merge_probe.py shard8 2>&1 | sed 's/^/nopin /'
