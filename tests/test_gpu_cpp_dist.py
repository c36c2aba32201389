"""GPU: BASELINE configs[4] driven from a plain C++ host -- tests/cpp/dist_spmv.cu uses only
include/loopsb.h (loopsb_dist_unique_id / _create / _spmv / _info / _x_full / _destroy) and the CUDA
runtime: one host thread per GPU, no Python, PyTorch or MPI on the data path. It is compiled by
__graft_entry__.build() (so the C ABI of the multi-GPU step is proven to be usable from C++ as
declared). The program was written after the round's GPU budget was spent: it has been compiled and
its no-device error path run, NOT yet executed on hardware -- so it only runs on request
(LOOPSB_RUN_CPP_DIST=1), never as part of the default suite. The same C calls are exercised on the
GPU through the ctypes mirror by tests/test_gpu_dist.py and bench.py --gpus N."""
import os
import subprocess

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "dist_spmv")


@pytest.mark.parametrize("world,groups", [(1, ""), (2, ""), (2, "1"), (4, "1,2"), (8, "")])
def test_cpp_host_runs_the_multi_gpu_step(world, groups):
    if os.environ.get("LOOPSB_RUN_CPP_DIST") != "1":
        pytest.skip("on request only (LOOPSB_RUN_CPP_DIST=1): not yet executed on hardware, see the module docstring")
    if not os.path.exists(EXE):
        pytest.skip("tests/cpp/dist_spmv not built (run __graft_entry__.build())")
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [EXE, "--world", str(world)] + (["--groups", groups] if groups else [])
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    oks = [l for l in p.stdout.splitlines() if l.startswith("OK rank")]
    assert len(oks) == world, p.stdout
