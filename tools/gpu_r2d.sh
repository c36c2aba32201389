#!/bin/bash
# round 2, multi-GPU call: parity under torchrun, then bench at N = number of visible GPUs
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -k "real_gpus" 2>&1 | tail -6
for G in ${GROUPS_LIST:-default 0}; do
  if [ "$G" = "default" ]; then unset LOOPSB_DIST_GROUPS; else export LOOPSB_DIST_GROUPS=$G; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps ${STEPS:-50} --warmup 5 $([ "$G" = "default" ] || echo --no-same-workload) > gpurun_out/bench_n${N}_g${G}.json 2> gpurun_out/bench_n${N}_g${G}.err
  echo "bench N=$N groups=$G rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_n${N}_g${G}.json"))
    print("N=%d groups=%s: %.1f Gnnz/s  step %.3f ms  comm %.3f ms  kernel %.3f ms  blocks %s  e2e %.1f  y_ok %s  same-workload-1gpu %s" % (
        d["n_gpus"], d["breakdown"]["groups"], d["value"]/1e9, d["ms_per_step"], d["comm_ms"], d["kernel_ms"],
        ["%.3f" % v for v in d["breakdown"]["block_ms_rank0"]], d["e2e"]["value"]/1e9, d["y_matches_oracle"],
        d["single_gpu_same_workload"]))
    print(json.dumps(d["dist_check"]))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_n${N}_g${G}.err").read()[-2500:])
PY
done
