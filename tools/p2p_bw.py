"""tools/p2p_bw.py -- copy-engine peer copy bandwidth between GPU 0 and GPU 1 of one box:
issued on the DESTINATION device's stream (pull) vs on the SOURCE device's stream (push),
8 MB and 32 MB, 1..3 copies in flight."""
import torch
assert torch.cuda.device_count() >= 2
for mb in (8, 32):
    n = mb << 18
    src = [torch.ones(n, device="cuda:0") for _ in range(3)]
    dst = [torch.empty(n, device="cuda:1") for _ in range(3)]
    for mode, dev in (("pull", 1), ("push", 0)):
        for conc in (1, 2, 3):
            torch.cuda.set_device(dev)
            streams = [torch.cuda.Stream(device=dev) for _ in range(conc)]
            def run(reps):
                for r in range(reps):
                    for j, st in enumerate(streams):
                        with torch.cuda.stream(st):
                            dst[j].copy_(src[j], non_blocking=True)
            run(3)
            torch.cuda.synchronize(0); torch.cuda.synchronize(1)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            import time
            t0 = time.perf_counter()
            run(20)
            torch.cuda.synchronize(0); torch.cuda.synchronize(1)
            dt = time.perf_counter() - t0
            print(f"{mb:3d} MB {mode} x{conc}: {20 * conc * n * 4 / dt / 1e9:7.1f} GB/s  ({dt / 20 * 1e6:.1f} us per round)", flush=True)
