#!/bin/bash
# lean multi-GPU bench sweep: GROUPS_LIST of LOOPSB_DIST_GROUPS values, no pytest, no same-workload leg
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for G in ${GROUPS_LIST:-default}; do
  if [ "$G" = "default" ]; then unset LOOPSB_DIST_GROUPS; else export LOOPSB_DIST_GROUPS=$G; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps ${STEPS:-40} --warmup 5 --no-same-workload > gpurun_out/bench_n${N}_g${G}.json 2> gpurun_out/bench_n${N}_g${G}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_n${N}_g${G}.json"))
    print("N=%d groups=%s: %.1f Gnnz/s  step %.3f ms  comm %.3f ms  kernel %.3f ms  blocks %s  e2e %.1f  y_ok %s" % (
        d["n_gpus"], d["breakdown"]["groups"], d["value"]/1e9, d["ms_per_step"], d["comm_ms"], d["kernel_ms"],
        ["%.3f" % v for v in d["breakdown"]["block_ms_rank0"]], d["e2e"]["value"]/1e9, d["y_matches_oracle"]), "graphs", d["dist"].get("graphs_cached"), d["dist"].get("transport"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_n${N}_g${G}.err").read()[-1500:])
PY
done
