"""GPU parity of the multi-GPU building blocks (loops_b200/csrc/dist.cu):
the column-block split, y += A x, loopsb_dist_* on one rank, and -- when the box has at
least two GPUs -- the whole phased all-gather + SpMV step under torchrun, every rank's y
against the oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import random_csr
from merge2_emul import spmv_merge2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exact_case(rows=3000, cols=4096, density=0.004, seed=3):
    off, idx, val = random_csr(rows, cols, density, seed=seed, exact=True, empty_every=11)
    x = np.random.default_rng(seed).integers(1, 11, cols).astype(np.float32)
    return off, idx, val, x


def test_split_columns_matches_host_statement():
    from loops_b200 import csr_t
    from loops_b200.convert import csr_split_columns
    from loops_b200.dist import ring_blocks, split_column_blocks
    off, idx, val, _ = _exact_case()
    A = csr_t(3000, 4096, off, idx, val)
    for world, rank, groups in ((8, 3, [2, 2, 3]), (4, 0, [1, 2]), (2, 1, [1]), (8, 7, [7])):
        boc = ring_blocks(world, rank, groups)
        got = csr_split_columns(A, 4096 // world, boc)
        want = split_column_blocks(off, idx, val, 4096 // world, boc)
        assert len(got) == len(want)
        for g, (w_off, w_idx, w_val) in zip(got, want):
            np.testing.assert_array_equal(g.offsets.cpu().numpy(), w_off)
            np.testing.assert_array_equal(g.indices.cpu().numpy(), w_idx)
            np.testing.assert_array_equal(g.values.cpu().numpy(), w_val)
            assert g.nnzs == 0 or g.indices.data_ptr() % 16 == 0       # every block on a 16-byte boundary


def test_accumulate_over_column_blocks(oracle):
    """y = A_0 x, then y += A_b x block by block (loopsb_spmv_acc_f32): exact on exact
    inputs, and bit-equal to the kernel's host emulation on floats."""
    import ctypes as C
    from loops_b200 import _lib, csr_t
    from loops_b200.algorithms import spmv
    from loops_b200.convert import csr_split_columns
    from loops_b200.dist import ring_blocks
    lib = _lib.load()
    for exact in (True, False):
        off, idx, val = random_csr(5000, 8192, 0.003, seed=8, exact=exact, empty_every=13)
        rng = np.random.default_rng(2)
        x = rng.integers(1, 11, 8192).astype(np.float32) if exact else rng.uniform(-1, 1, 8192).astype(np.float32)
        A = csr_t(5000, 8192, off, idx, val)
        blocks = csr_split_columns(A, 1024, ring_blocks(8, 2, [2, 2, 3]))
        xd = torch.as_tensor(x).cuda()
        y = torch.full((5000,), float("nan"), device="cuda")
        emu = None
        for b, B in enumerate(blocks):
            plan = B.plan(_lib.SCHED_MERGE_PATH_FLAT, None, tiled=False)
            if b == 0:
                spmv.merge_path_flat(B, xd, y, tiled=False)
                emu = spmv_merge2(B.offsets.cpu().numpy(), B.indices.cpu().numpy(), B.values.cpu().numpy(), x)
            else:
                _lib.check(lib.loopsb_spmv_acc_f32(plan.handle, _lib.ptr(B.values), _lib.ptr(B.indices), _lib.ptr(xd),
                                                   _lib.ptr(y), 5000, 8192, _lib.stream_ptr(None)), "acc")
                emu = spmv_merge2(B.offsets.cpu().numpy(), B.indices.cpu().numpy(), B.values.cpu().numpy(), x,
                                  y_init=emu)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(y.cpu().numpy(), emu)
        if exact:
            np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv(off, idx, val, x))


def test_dist_plan_world_1(oracle):
    from loops_b200 import csr_t
    from loops_b200.dist import DistPlan
    off, idx, val, x = _exact_case()
    A = csr_t(3000, 4096, off, idx, val)
    dp = DistPlan(A, 1, 0)
    y = torch.full((3000,), float("nan"), device="cuda")
    dp(torch.as_tensor(x).cuda(), y)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv(off, idx, val, x))
    assert dp.info()["num_blocks"] == 1
    np.testing.assert_array_equal(dp.x_full(4096).cpu().numpy(), x)
    dp.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_dist_step_on_real_gpus(world):
    """The whole step on `world` GPUs of this box (skipped when it has fewer): torchrun +
    tools/dist_check.py, which compares every rank's y shard with the oracle for the
    single-all-gather path and for the phased column-block path."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    port = 29600 + world
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, timeout=900,
                       cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-1500:], p.stderr[-3000:])
    line = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["world"] == world and line["all_ok"] is True, line
