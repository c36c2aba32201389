#!/bin/bash
# tools/gpu_scale.sh N -- bench.py on N GPUs of one box (torchrun), JSON to gpurun_out/bench_nN.json
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value %.1f Gnnz/s  step %.3f ms  kernel %.3f ms  e2e %.1f Gnnz/s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms_mean"], d["e2e"]["value"]/1e9))
PY
grep -i -E "error|NVLS|Traceback" gpurun_out/bench_n$N.err | head -5
