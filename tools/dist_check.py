"""tools/dist_check.py -- run under torchrun on N GPUs of one box: every rank builds its row
shard of a small synthetic matrix, runs loopsb_dist_spmv (C ABI) with (a) one ncclAllGather and
(b) the phased send/recv all-gather over column blocks, and compares its y shard with the
oracle's SpMV of the same rows (test infrastructure: the oracle is the checker). Rank 0 prints
one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from helpers import Oracle
    from loops_b200 import csr_t, generate as g
    from loops_b200.dist import DistPlan, default_groups, row_range
    oracle = Oracle(os.path.join(ROOT, "oracle", "libloops_oracle.so"))
    rows = cols = 1 << 17
    nnz = rows * 20
    deg = g.powerlaw_degrees(rows, nnz, d_max=512)
    r0, r1 = row_range(rows, rank, world)
    off, idx, val = g.synth_csr(rows, cols, nnz, device=dev, degrees=deg, row_begin=r0, row_end=r1)
    A = csr_t.from_tensors(r1 - r0, cols, off, idx, val)
    x = g.x_recipe(cols, device=dev)
    n = cols // world
    ref = oracle.spmv(off.cpu().numpy(), idx.cpu().numpy(), val.cpu().numpy(), x.cpu().numpy())
    results = {}
    for name, groups, transport in (("single_allgather", [], None), ("phased", default_groups(world), None),
                                    ("phased_one_chunk_each", [1] * (world - 1), None),
                                    ("phased_nccl_transport", default_groups(world), "nccl")):
        if name == "phased_one_chunk_each" and world > 8:
            continue
        if transport:
            os.environ["LOOPSB_DIST_TRANSPORT"] = transport
        else:
            os.environ.pop("LOOPSB_DIST_TRANSPORT", None)
        dp = DistPlan.from_process_group(A, groups=groups)
        ok = True
        xs = torch.empty(n, device=dev)
        y = torch.empty(r1 - r0, device=dev)
        side = torch.cuda.Stream(device=dev)     # a real stream: the phased step is replayed as a CUDA graph
        torch.cuda.synchronize()
        for it in range(7):       # repeated steps re-use x_full, the staging buffers, flags, graphs
            use = side if it >= 2 else torch.cuda.current_stream()
            with torch.cuda.stream(use):
                xs.copy_(x[rank * n:(rank + 1) * n] * float(it + 1))
                y.fill_(float("nan"))
                dp(xs, y, use)
            torch.cuda.synchronize()
            ok = ok and bool(np.array_equal(y.cpu().numpy(), ref * np.float32(it + 1)))
            ok = ok and bool(torch.equal(dp.x_full(cols), x * float(it + 1)))
        flag = torch.tensor([0.0 if ok else 1.0], device=dev)
        dist.all_reduce(flag)
        results[name] = {"ok": bool(flag.item() == 0.0), "groups": groups, "blocks": dp.info()["num_blocks"],
                         "transport": dp.info()["transport"]}
        dp.close()
        dist.barrier()
    if rank == 0:
        print(json.dumps({"world": world, "all_ok": all(v["ok"] for v in results.values()), "results": results}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
