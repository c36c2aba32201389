#!/bin/bash
# tools/gpu_prof.sh -- microbench ceilings + ncu captures of the merge kernel.
mkdir -p gpurun_out
echo "== microbench"
timeout 300 ./tools/microbench | tee gpurun_out/microbench.txt
echo "== bench"
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"
cat gpurun_out/bench2.json | python -c "import sys,json; d=json.load(sys.stdin); print(d['value']/1e9,'Gnnz/s', d['ms_per_step'],'ms', 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value']/1e9)"
echo "== ncu launch list (our kernels)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmv_|merge_" -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/launches.csv | cut -d, -f5,15 | tail -14
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_merge_kernel -s 4 -c 2 -f -o gpurun_out/prof_merge \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
