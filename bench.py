#!/usr/bin/env python
"""bench.py -- CSR SpMV throughput (nnz/s) of the merge_path_flat hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one SpMV  y = A x  over the whole matrix.
  N == 1 : BASELINE.json configs[1] -- synthetic power-law CSR, 2^20 rows,
           2^25 nnz, fp32, merge_path_flat, one B200.
  N  > 1 : BASELINE.json configs[4] -- 2^24 rows / 2^29 nnz row-partitioned over
           the N ranks (contiguous row ranges, global column ids); every step is
           loopsb_dist_spmv (C ABI, loops_b200/csrc/dist.cu): the NCCL all-gather of
           the dense x shards, phased so that it overlaps the local merge-path SpMV
           of the shard's column blocks. Total work is fixed as N grows ("strong").
Inputs are generated in HBM (deterministic counter-based generator, x = the
reference's recipe) and exceed the 126 MB L2 (272 MB of matrix per step), so
consecutive timed steps cannot be served from cache ("l2": "inputs_exceed_l2").

One JSON line on stdout (rank 0). `value` = device-timed whole-job nnz/s with
inputs resident in HBM; `e2e` = the same through the public API with x coming
from pinned host memory and y going back every step; `roofline` = algorithmic
bytes / CUDA-event duration of the merge-path kernel alone (probes on the
launching stream) against MEASURED_PEAKS.json; `cpu_baseline` = the reference's
own CPU validator (oracle/_ref, built from its unmodified headers) on this
box's host cores.

`--impl reference` times that CPU validator as the reference arm, on all host
threads (each thread runs the unmodified reference::spmv on a contiguous row
slice), for the same config / metric / unit.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "csr_spmv_nnz_per_s"
UNIT = "nnz/s"
CFG2 = dict(rows=1 << 20, cols=1 << 20, nnz=1 << 25)
CFG5 = dict(rows=1 << 24, cols=1 << 24, nnz=1 << 29)
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def algorithmic_bytes(rows, cols, nnz):
    """nnz*(4+4) + (rows+1)*4 + cols*4 + rows*4 (x and y once)."""
    return nnz * 8 + (rows + 1) * 4 + cols * 4 + rows * 4


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)   # the timed region is ~12 ms at N=1: several samples inside it

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def report(self):
        if not self.samples:
            return None
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def ref_host_lib():
    p = os.path.join(ROOT, "oracle", "_ref", "libloopsref_host.so")
    return C.CDLL(p) if os.path.exists(p) else None


def cpu_spmv_seconds(off, idx, val, x, rows, cols, reps, threads):
    """Best-of-`reps` seconds per SpMV of the reference CPU validator
    (oracle/_ref) or, if that library is absent, of the oracle port."""
    import numpy as np
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    nnz = len(idx)
    L = ref_host_lib()
    if L is not None:
        L.ref_time_spmv.restype = C.c_double
        y = np.zeros(rows, np.float32)
        return L.ref_time_spmv(rows, cols, nnz, P(off), P(idx), P(val), P(x), reps, threads, P(y)), "reference", y
    so = os.path.join(ROOT, "oracle", "libloops_oracle.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    O = C.CDLL(so)
    y = np.zeros(rows, np.float32)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        O.orc_spmv_f32(rows, P(off), P(idx), P(val), P(x), P(y))
        best = min(best, time.perf_counter() - t0)
    return best, "port", y


def make_inputs(cfg, device, row_begin=0, row_end=None):
    from loops_b200 import generate as g
    deg = g.powerlaw_degrees(cfg["rows"], cfg["nnz"], d_max=min(1024, cfg["cols"]))
    off, idx, val = g.synth_csr(cfg["rows"], cfg["cols"], cfg["nnz"], device=device, degrees=deg,
                                row_begin=row_begin, row_end=row_end)
    return deg, off, idx, val


def run_reference_arm(args, rank, world):
    """Reference arm: the reference's CPU SpMV on the host cores (rank 0 only)."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from loops_b200 import generate as g
    cfg = CFG2 if args.gpus == 1 else CFG5
    # bounded sample: the first 2^20 rows / 2^25 nnz of the workload (that IS
    # the whole N=1 workload; for N>1 it is 1/16 of the rows, same generator)
    sample = dict(CFG2) if args.gpus == 1 else dict(rows=1 << 20, cols=cfg["cols"], nnz=None)
    if args.gpus == 1:
        _, off, idx, val = make_inputs(cfg, "cpu")
    else:
        deg = g.powerlaw_degrees(cfg["rows"], cfg["nnz"], d_max=1024)
        off, idx, val = g.synth_csr(cfg["rows"], cfg["cols"], cfg["nnz"], degrees=deg, row_begin=0,
                                    row_end=sample["rows"])
    x = g.x_recipe(cfg["cols"]).numpy()
    off, idx, val = off.numpy(), idx.numpy(), val.numpy()
    rows, nnz = len(off) - 1, len(idx)
    threads = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_spmv_seconds(off, idx, val, x, rows, cfg["cols"], 1, threads)
    t0 = time.perf_counter()
    sec, kind, _ = cpu_spmv_seconds(off, idx, val, x, rows, cfg["cols"], max(args.steps, 1), threads)
    wall = time.perf_counter() - t0
    value = nnz / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{rows} rows / {nnz} nnz of the workload, best of {args.steps} "
                                   f"passes, {threads} threads x unmodified reference::spmv on row slices"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    emit(line)


def workload_config(n, groups=None):
    if n == 1:
        return {"workload": "synthetic power-law CSR 2^20 rows / 2^25 nnz fp32, merge_path_flat (BASELINE configs[1])",
                "rows": CFG2["rows"], "nnz": CFG2["nnz"], "schedule": "merge_path_flat", "layout": "csr",
                "degree_law": "P(d)~d^-2.1, d in [1,1024], seeded row order", "columns": "stratified hash, unique ascending",
                "l2": "inputs_exceed_l2", "partition": "none"}
    if groups is None:
        from loops_b200.dist import default_groups
        groups = default_groups(n)
    return {"workload": "synthetic power-law CSR 2^24 rows / 2^29 nnz fp32, merge_path_flat, row-partitioned, "
                        "NCCL all-gather of x per step (BASELINE configs[4])",
            "rows": CFG5["rows"], "nnz": CFG5["nnz"], "schedule": "merge_path_flat", "layout": "csr",
            "l2": "inputs_exceed_l2", "partition": f"row{n}",
            "collective": ("nccl all-gather of x issued as ring-shifted send/recv phases " + str(groups) +
                           " (chunks per phase) overlapping the SpMV of the shard's column blocks"
                           if groups else "one ncclAllGather(x), then one SpMV"),
            "api": "loopsb_dist_spmv (C ABI; NCCL driven by libloopsb200.so)",
            "note": "N>1 runs BASELINE configs[4], 16x the N=1 workload (configs[1]): total work is fixed across "
                    "N=2/4/8 (strong scaling among them), but it is NOT the N=1 matrix -- its x is 64 MB, so the "
                    "band-tiled plan's cost model declines and every shard runs the CSR merge-path kernel; "
                    "`single_gpu_same_workload` gives this workload on one GPU"}


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# Libraries (NCCL's version banner, torchrun notices) must not pollute stdout:
# keep a private handle on it and point fd 1 at stderr for everything else.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="N = 1: skip the `extra` block (BASELINE configs[2] sweep and configs[3] BCSR, timed after the headline)")
    ap.add_argument("--no-same-workload", action="store_true",
                    help="N > 1: skip timing the whole configs[4] matrix on rank 0's GPU alone")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 5          # CPU passes are ~0.1-1 s each; keep the arm within minutes
        args.warmup = min(args.warmup, 2)
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from loops_b200 import _lib, csr_t, generate as g
    from loops_b200.algorithms import spmv

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: loops-b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    N = args.gpus
    if world != N:
        if world == 1 and N > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    if N > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = CFG2 if N == 1 else CFG5
    rows, cols, nnz = cfg["rows"], cfg["cols"], cfg["nnz"]
    r0, r1 = (rows * rank) // N, (rows * (rank + 1)) // N
    deg, off, idx, val = make_inputs(cfg, dev, r0, r1)
    A = csr_t.from_tensors(r1 - r0, cols, off, idx, val)
    x_full = g.x_recipe(cols, device=dev)
    x_shard = x_full[(cols * rank) // N: (cols * (rank + 1)) // N].clone()
    y = torch.empty(r1 - r0, dtype=torch.float32, device=dev)
    if N > 1:
        # a real (non-legacy) stream: loopsb_dist_spmv replays its step as a CUDA graph on it
        torch.cuda.synchronize()
        torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    stream = torch.cuda.current_stream()
    t_plan = time.perf_counter()
    plan = A.plan(_lib.SCHED_MERGE_PATH_FLAT, stream)   # preprocess (not timed, as in the reference:
    torch.cuda.synchronize()                            # merge_path_flat.cuh:111 vs :121-122)
    plan_s = time.perf_counter() - t_plan
    info = plan.info()
    tiled = plan.tiled_info()            # None when the cost model kept the plain CSR kernel
    launches_per_spmv = 1 if tiled else int(info.launches_per_spmv)

    # N > 1: the whole step is loopsb_dist_spmv (C ABI). torch.distributed only hands
    # rank 0's NCCL id to the other ranks and reduces the timings.
    dp = None
    groups = []
    if N > 1:
        from loops_b200.dist import DistPlan, default_groups
        groups = default_groups(N)
        t_dp = time.perf_counter()
        dp = DistPlan.from_process_group(A, groups=groups, stream=stream)
        torch.cuda.synchronize()
        plan_s += time.perf_counter() - t_dp
        dinfo = dp.info()
        launches_per_spmv = 2 * dinfo["num_blocks"] + (len(groups) if groups else 1)

    # N = 1: the step is ONE call of the C ABI, made the way a C/C++ host makes it (handle and
    # pointers resolved once; ~3 us of host time per call instead of ~20 us through the Python
    # mirror, which matters for the first launch after the opening barrier of a 20-step run).
    # spmv.merge_path_flat(A, x, y) is the same call behind argument checks; the e2e block and
    # the parity guard below go through it.
    lib = _lib.load()
    abi_args = (plan.handle, _lib.ptr(A.values), _lib.ptr(A.indices), None, _lib.ptr(x_full), _lib.ptr(y),
                r1 - r0, cols, _lib.stream_ptr(stream))

    def step():
        if N > 1:
            dp(x_shard, y, stream)
        else:
            rc = lib.loopsb_spmv_f32(*abi_args)
            if rc:
                _lib.check(rc, "loopsb_spmv_f32")

    def barrier():
        if N > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    if N > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = nnz / (ms_step * 1e-3)

    # ---- kernel-only probes (second pass; not part of `value`) ----
    breakdown = None
    if N > 1:
        # per step: CUDA events around the all-gather (first send/recv issued -> last chunk
        # landed) and around every column block's SpMV; one step at a time (probe_read syncs)
        dp.probe(True)
        reads = []
        for _ in range(min(args.steps, 20)):
            barrier()
            step()
            reads.append(dp.probe_read())
        dp.probe(False)
        kernel_ms = np.array([r["kernel_ms"] for r in reads], np.float32)
        comm = torch.tensor([float(np.mean([r["comm_ms"] for r in reads])), float(np.mean(kernel_ms))], device=dev)
        cmax = comm.clone(); dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
        csum = comm.clone(); dist.all_reduce(csum, op=dist.ReduceOp.SUM)
        breakdown = {"comm_ms": float(cmax[0].item()), "kernel_ms": float(cmax[1].item()),
                     "comm_ms_mean_over_ranks": float(csum[0].item()) / N,
                     "kernel_ms_mean_over_ranks": float(csum[1].item()) / N,
                     "block_ms_rank0": [float(v) for v in np.mean([r["block_ms"] for r in reads], axis=0)],
                     "block_nnz_rank0": dinfo["block_nnz"], "groups": groups,
                     "x_bytes_received_per_rank": int(cols * 4 * (N - 1) // N),
                     "allgather_GBps_in": cols * 4 * (N - 1) / N / (float(cmax[0].item()) * 1e-3) / 1e9,
                     "how": "max over ranks of the per-rank means of 20 single steps with probes on: comm = first "
                            "send/recv issued -> last chunk landed (side stream), kernel = sum of the column "
                            "blocks' SpMV launches; with phases the two overlap, so comm + kernel > step"}
    else:
        plan.probe_begin(args.steps)
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        kernel_ms = plan.probe_collect(args.steps)
    local_nnz = int(idx.numel())
    local_bytes = algorithmic_bytes(r1 - r0, cols, local_nnz)
    peak, peak_src = measured_peak()
    k_pair_ms = float(np.mean(kernel_ms))
    if N == 1 and launches_per_spmv == 1:
        # The step IS one launch of this kernel, so the timed region (one event pair around
        # K back-to-back launches) divided by K is its average duration, launch gaps included.
        # An event pair around every single launch adds ~2 us of record/start latency to a
        # 60 us kernel and keeps consecutive launches from overlapping; it is kept below as
        # kernel_ms_event_pair_*.
        k_ms, k_src = ms_step, ("timed region / K (step = 1 launch of this kernel; CUDA events on the launching stream). "
                                "Launches after the first are programmatic dependent launches: the prologue of launch n+1 "
                                "runs under the write-out of launch n, so this average period is shorter than one launch "
                                "timed alone (kernel_ms_event_pair_*, the serialised ncu list); LOOPSB_TILED_PDL=0 turns it off")
    else:
        k_ms, k_src = k_pair_ms, "mean of per-launch CUDA event pairs on the launching stream (second pass)"
    achieved = local_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": (f"spmv_bt_kernel (band-tiled; CTA {tiled['cta_threads']} thr, grid {tiled['grid_blocks']}, "
                           f"smem {tiled['smem_bytes']} B)" if tiled else
                           f"spmv_merge2_kernel (CTA {info.cta_threads} thr, grid {info.grid_blocks}, smem {info.smem_bytes} B"
                           + (f"; {dinfo['num_blocks']} column blocks per step" if N > 1 else "") + ")"),
                "kernel_ms_mean": k_ms, "duration_source": k_src,
                "kernel_ms_event_pair_mean": k_pair_ms, "kernel_ms_event_pair_min": float(np.min(kernel_ms)),
                "frac_event_pair": local_bytes / (k_pair_ms * 1e-3) / 1e9 / peak,
                "algorithmic_bytes_per_launch": local_bytes,
                "frac_of_nominal_8TBs": achieved / 8000.0}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if N == 1:   # the capture is of the N=1 workload
            roofline["traffic"] = (tj.get("tiled") if tiled else tj.get("plain", tj)).get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- cold-L2 variant (reported beside the back-to-back figure, SURVEY 8d "Timing") ----
    # Between launches a 512 MB buffer is overwritten (4x the 126 MB L2), so neither
    # x nor the head of the matrix stream is resident; one event pair per launch.
    if N == 1:
        flush = torch.empty(128 << 20, dtype=torch.float32, device=dev)
        cold = []
        for i in range(12):
            flush.fill_(float(i))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(); b.record()
            torch.cuda.synchronize()
            cold.append(a.elapsed_time(b))
        del flush
        cold = sorted(cold[2:])
        cold_ms = cold[len(cold) // 2]
        roofline["cold_l2"] = {"ms_median": cold_ms, "frac": local_bytes / (cold_ms * 1e-3) / 1e9 / peak,
                               "how": "512 MB written between launches; CUDA event pair per launch (includes its ~2 us)"}

    # ---- e2e: public API, x from pinned host memory, y back to the host ----
    # Every step uploads its x (pinned host -> HBM) and downloads its y. Two
    # flavours: `serial` = copy-in, SpMV, copy-out strictly one after another
    # (a dependent iteration); `value` = the same per-step work with the copies
    # of neighbouring steps overlapped on side streams (double-buffered x / y),
    # i.e. the throughput a caller streaming independent right-hand sides gets.
    x_src = x_shard if N > 1 else x_full
    x_host = x_src.cpu().pin_memory()
    NBUF = 4     # depth of the copy/compute pipeline of the overlapped e2e flavour
    y_host = [torch.empty(r1 - r0, dtype=torch.float32).pin_memory() for _ in range(NBUF)]

    def e2e_serial_step():
        x_src.copy_(x_host, non_blocking=True)
        if N > 1:
            dp(x_shard, y, stream)
        else:
            spmv.merge_path_flat(A, x_full, y, stream=stream, sync=False)
        y_host[0].copy_(y, non_blocking=True)

    def timed(fn_loop):
        barrier()
        t0 = time.perf_counter()
        fn_loop()
        barrier()
        dt = time.perf_counter() - t0
        if N > 1:
            tt = torch.tensor([dt], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return dt

    for _ in range(3):
        e2e_serial_step()
    serial_s = timed(lambda: [e2e_serial_step() for _ in range(args.steps)])

    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    xd = [x_src.clone() for _ in range(NBUF)]
    yd = [y] + [y.clone() for _ in range(NBUF - 1)]
    ev_in = [torch.cuda.Event() for _ in range(NBUF)]
    ev_done = [torch.cuda.Event() for _ in range(NBUF)]
    ev_out = [torch.cuda.Event() for _ in range(NBUF)]
    for e in ev_done + ev_out:
        e.record(stream)

    def e2e_pipelined_torch(steps):
        for k in range(steps):
            b = k % NBUF
            with torch.cuda.stream(s_in):
                s_in.wait_event(ev_done[b])           # the SpMV that last read xd[b] is finished
                xd[b].copy_(x_host, non_blocking=True)
                ev_in[b].record(s_in)
            stream.wait_event(ev_in[b])
            stream.wait_event(ev_out[b])              # yd[b] has been drained to the host
            if N > 1:
                dp(xd[b], yd[b], stream)
            else:
                spmv.merge_path_flat(A, xd[b], yd[b], stream=stream, sync=False)
            ev_done[b].record(stream)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[b])
                y_host[b].copy_(yd[b], non_blocking=True)
                ev_out[b].record(s_out)

    # N == 1: the same pipeline issued through the CUDA runtime and the C ABI directly
    # (what a C/C++ caller of libloopsb200.so does) -- ten driver calls per step
    # instead of ten Python-dispatched torch ops, so the loop is bound by PCIe and
    # the kernel rather than by the interpreter.
    rt = None
    if N == 1:
        try:
            rt = C.CDLL("libcudart.so.12")
            for fn in ("cudaMemcpyAsync", "cudaEventRecord", "cudaStreamWaitEvent"):
                getattr(rt, fn).restype = C.c_int
            rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
            rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
            rt.cudaStreamWaitEvent.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
        except OSError:
            rt = None

    def e2e_pipelined_rt(steps):
        lib = _lib.load()
        h = plan.handle
        S, SI, SO = stream.cuda_stream, s_in.cuda_stream, s_out.cuda_stream
        e_in = [e.cuda_event for e in ev_in]
        e_done = [e.cuda_event for e in ev_done]
        e_out = [e.cuda_event for e in ev_out]
        xb_, yb_ = x_host.numel() * 4, y_host[0].numel() * 4
        xptr, yptr = [t.data_ptr() for t in xd], [t.data_ptr() for t in yd]
        hx, hy = x_host.data_ptr(), [t.data_ptr() for t in y_host]
        vptr, iptr = A.values.data_ptr(), A.indices.data_ptr()
        rc = 0
        for k in range(steps):
            b = k % NBUF
            rc |= rt.cudaStreamWaitEvent(SI, e_done[b], 0)
            rc |= rt.cudaMemcpyAsync(xptr[b], hx, xb_, 1, SI)            # cudaMemcpyHostToDevice
            rc |= rt.cudaEventRecord(e_in[b], SI)
            rc |= rt.cudaStreamWaitEvent(S, e_in[b], 0)
            rc |= rt.cudaStreamWaitEvent(S, e_out[b], 0)
            rc |= lib.loopsb_spmv_f32(h, vptr, iptr, None, xptr[b], yptr[b], r1 - r0, cols, S)
            rc |= rt.cudaEventRecord(e_done[b], S)
            rc |= rt.cudaStreamWaitEvent(SO, e_done[b], 0)
            rc |= rt.cudaMemcpyAsync(hy[b], yptr[b], yb_, 2, SO)         # cudaMemcpyDeviceToHost
            rc |= rt.cudaEventRecord(e_out[b], SO)
        if rc:
            raise RuntimeError("CUDA runtime / loopsb call failed in the e2e pipeline")

    e2e_pipelined = e2e_pipelined_rt if rt is not None else e2e_pipelined_torch
    for e in ev_in:
        e.record(stream)          # events must exist (be recorded once) before their raw handles are used
    torch.cuda.synchronize()

    e2e_pipelined(2 * NBUF)
    torch.cuda.synchronize()
    # wall-clock over K steps is sensitive to what else the host and the PCIe link are
    # doing; the loop is repeated three times and the MEDIAN repetition is reported
    e2e_reps = []
    for _ in range(3):
        e2e_reps.append(timed(lambda: e2e_pipelined(args.steps)))
        torch.cuda.synchronize()
    e2e_s = sorted(e2e_reps)[1]
    e2e = {"value": nnz / (e2e_s / args.steps), "unit": UNIT,
           "h2d_bytes_per_step": int(x_host.numel() * 4), "d2h_bytes_per_step": int(y_host[0].numel() * 4),
           "ms_per_step": e2e_s / args.steps * 1e3,
           "ms_per_step_repetitions": [t / args.steps * 1e3 for t in e2e_reps],
           "serial_value": nnz / (serial_s / args.steps), "serial_ms_per_step": serial_s / args.steps * 1e3,
           "note": "matrix resident in HBM (the reference API's csr_t is device-resident); every step uploads "
                   "x from pinned host memory and downloads y; `value` overlaps the copies of neighbouring "
                   "steps on side streams (4 buffers in flight; at N=1 issued through the CUDA runtime and the "
                   "C ABI directly, as a C caller would), `serial_value` runs copy-in/SpMV/copy-out back to "
                   "back; wall clock, final synchronize on all streams; median of three repetitions of the K-step loop. On the pool's (virtualised, one NUMA node) boxes this figure is bimodal from PROCESS to process -- ~95 us or ~130 us per step, stable within a process, unaffected by CPU affinity or GC -- i.e. it follows the host's PCIe path, not this code"}
    y_e2e_ok = bool(torch.equal(torch.from_numpy(y_host[(args.steps - 1) % NBUF].numpy()).to(dev), yd[(args.steps - 1) % NBUF]))

    # correctness guard on the timed configuration (exact inputs -> exact sums)
    step()
    torch.cuda.synchronize()
    chk = float(y.double().sum().item())

    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        o, i_, v, xx = off.cpu().numpy(), idx.cpu().numpy(), val.cpu().numpy(), x_full.cpu().numpy()
        sec, kind, y_cpu = cpu_spmv_seconds(o, i_, v, xx, rows, cols, 5, 1)
        cpu = {"value": nnz / sec, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": "the full 2^20 x 2^25 workload, best of 5 passes of reference::spmv as shipped "
                         "(single thread, includes its internal host copies)",
               "host_cores_available": os.cpu_count(),
               "y_matches_gpu_bit_exact": bool(np.array_equal(y_cpu, y.cpu().numpy()))}

    # ---- N > 1: parity on the hardware that ran the timed steps ----
    #  * every rank: the first 2^16 rows of its y shard against the reference's CPU SpMV
    #    (the cpu_baseline leg's library) on the same rows, bit for bit;
    #  * globally: sum(y) over all ranks (fp64; every term is a multiple of 1/8, so the sum
    #    is exact) against sum_j x_j * (column sum_j of A), formed with torch ops only.
    dist_check = None
    same_workload = None
    if N > 1:
        prefix = min(1 << 16, r1 - r0)
        pn = int(off[prefix].item())
        o, i_, v = off[: prefix + 1].cpu().numpy(), idx[:pn].cpu().numpy(), val[:pn].cpu().numpy()
        xx = x_full.cpu().numpy()
        _, kind, y_cpu = cpu_spmv_seconds(o, i_, v, xx, prefix, cols, 1, 1)
        ok_prefix = bool(np.array_equal(y_cpu, y[:prefix].cpu().numpy()))
        colsum = torch.zeros(cols, dtype=torch.float64, device=dev).index_add_(0, idx.long(), val.double())
        sums = torch.stack([y.double().sum(), (colsum * x_full.double()).sum(),
                            torch.tensor(0.0 if ok_prefix else 1.0, dtype=torch.float64, device=dev)])
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        gathered_ok = bool(torch.equal(dp.x_full(cols), x_full))
        flag = torch.tensor([0.0 if gathered_ok else 1.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.SUM)
        dist_check = {"y_matches_oracle": bool(sums[2].item() == 0.0 and sums[0].item() == sums[1].item()),
                      "prefix_rows_per_rank": prefix, "prefix_checker": kind, "ranks_with_prefix_mismatch": int(sums[2].item()),
                      "global_checksum": float(sums[0].item()), "global_checksum_expected": float(sums[1].item()),
                      "gathered_x_equals_x_on_all_ranks": bool(flag.item() == 0.0)}
        del colsum
        # the same workload on ONE GPU (rank 0, after the timed region): the whole 2^24 x 2^29
        # matrix generated in 8 row slices, plain merge-path kernel (what loopsb_spmv_f32 runs
        # for it; x = 64 MB does not fit the band-tiled plan)
        if rank == 0 and not args.no_same_workload:
            try:
                offs, idxs, vals = [], [], []
                base = 0
                for sl in range(8):
                    a0, a1 = rows * sl // 8, rows * (sl + 1) // 8
                    o_, i__, v_ = g.synth_csr(rows, cols, nnz, device=dev, degrees=deg, row_begin=a0, row_end=a1)
                    offs.append(o_[:-1].long() + base if sl < 7 else o_.long() + base)
                    base += int(o_[-1].item())
                    idxs.append(i__); vals.append(v_)
                Afull = csr_t.from_tensors(rows, cols, torch.cat(offs).to(torch.int32), torch.cat(idxs), torch.cat(vals))
                del offs, idxs, vals
                yf = torch.empty(rows, dtype=torch.float32, device=dev)
                for _ in range(3):
                    spmv.merge_path_flat(Afull, x_full, yf, stream=stream, sync=False, tiled=False)
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record(stream)
                for _ in range(10):
                    spmv.merge_path_flat(Afull, x_full, yf, stream=stream, sync=False, tiled=False)
                b_.record(stream); b_.synchronize()
                ms1 = a_.elapsed_time(b_) / 10
                same_workload = {"value": nnz / (ms1 * 1e-3), "unit": UNIT, "ms_per_step": ms1, "n_gpus": 1,
                                 "y_checksum": float(yf.double().sum().item()),
                                 "how": "rank 0, after the timed region: whole configs[4] matrix resident on one "
                                        "B200, 10 back-to-back merge_path_flat SpMVs, CUDA events"}
                del Afull, yf
            except Exception as e:          # out of memory on a shared box: report, do not fail the bench
                same_workload = {"error": repr(e)[:200]}
        dist.barrier()

    # ---- N = 1: what the plan-owned tiled copy costs, and BASELINE configs[2] / configs[3] on the same box ----
    amort = None
    extra = None
    if N == 1 and rank == 0:
        try:
            if tiled:
                for _ in range(3):
                    spmv.merge_path_flat(A, x_full, y, stream=stream, sync=False, tiled=False)
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record(stream)
                for _ in range(20):
                    spmv.merge_path_flat(A, x_full, y, stream=stream, sync=False, tiled=False)
                b_.record(stream); b_.synchronize()
                plain_ms = a_.elapsed_time(b_) / 20
                t_re = time.perf_counter()
                spmv.merge_path_flat(A, x_full, y, stream=stream, sync=True, tiled="auto")   # rebuilds the copy (warm allocator)
                retile_s = time.perf_counter() - t_re - ms_step * 1e-3
                amort = {"plain_csr_kernel_ms": plain_ms, "band_tiled_kernel_ms": ms_step,
                         "first_plan_s": plan_s, "retile_s_warm": retile_s,
                         "breakeven_spmv_calls_measured": (int(retile_s / ((plain_ms - ms_step) * 1e-3)) + 1) if plain_ms > ms_step else None,
                         "breakeven_spmv_calls_model": plan.tile_breakeven(cols),
                         "extra_hbm_bytes": int(tiled["bytes"]),
                         "note": "the copy is built once per matrix outside the timed region, like the reference's "
                                 "preprocess (merge_path_flat.cuh:111 vs :121-122); a single cold SpMV is faster on the "
                                 "plain CSR kernel -- the C++ mirror tiles a cached plan only after the model's call count"}
        except Exception as e:
            amort = {"error": repr(e)[:200]}
        if not args.no_extra:
            try:
                A.drop_plans()
                torch.cuda.empty_cache()
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import sweep
                doc = sweep.run(small=False, reference=False, log=sys.stderr)
                extra = {"config3_sweep": [{k: c[k] for k in ("layout", "schedule", "ms_median", "gnnz_per_s",
                                                               "roofline_frac", "bit_equal_to_merge_csr")}
                                           for c in doc["cells"]],
                         "config4_bcsr4x4_bf16_tcgen05": doc["bcsr"], "ell_pitch": doc.get("ell_pitch"),
                         "spmm_csr_n32": doc.get("spmm"),
                         "timing": doc["timing"],
                         "note": "same box, after the headline; every cell checked bit-equal to merge_path_flat/CSR "
                                 "before it is timed; roofline_frac = the layout's algorithmic bytes / time / measured HBM peak"}
            except Exception as e:
                extra = {"error": repr(e)[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": N, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(N, groups), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches_per_spmv * args.steps,
            "clocks": clocks.report(),
            "plan": {"grid_blocks": info.grid_blocks, "cta_threads": info.cta_threads,
                     "smem_bytes": info.smem_bytes, "merge_tiles": int(info.num_merge_tiles),
                     "preprocess_s_untimed": plan_s, "band_tiled": tiled,
                     "kernel": "band-tiled (plan-owned re-ordered copy of the matrix)" if tiled else "csr merge-path"},
            "y_checksum": chk, "e2e_y_equal_device_y": y_e2e_ok,
        }
        if amort is not None:
            line["plan"]["amortisation"] = amort
        if extra is not None:
            line["extra"] = extra
        if N > 1:
            line["comm_ms"] = breakdown["comm_ms"]
            line["kernel_ms"] = breakdown["kernel_ms"]
            line["breakdown"] = breakdown
            line["y_matches_oracle"] = dist_check["y_matches_oracle"]
            line["dist_check"] = dist_check
            line["single_gpu_same_workload"] = same_workload
            line["dist"] = dp.info()
        emit(line)
    if N > 1:
        dp.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
