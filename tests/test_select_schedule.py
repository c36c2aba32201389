"""Schedule selection: `loopsb_select_schedule` is a pure host
function, so it is tested without a GPU -- against hand cases that restate the
thresholds measured in profiles/heuristic_sweep_r01.log, and against the outcomes
of the reference's own heuristic (plots/data/heuristics.csv -> tests/golden/
heuristics_ref.npz, made by tests/golden/make_heuristics_fixture.py)."""
import ctypes as C
import os

import numpy as np

from loops_b200 import _lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MERGE, THREAD = _lib.SCHED_MERGE_PATH_FLAT, _lib.SCHED_THREAD_MAPPED


def pick(rows, cols, nnz, max_degree=-1):
    out = C.c_int32(-1)
    _lib.check(_lib.load().loopsb_select_schedule(rows, cols, nnz, max_degree, C.byref(out)), "select")
    return int(out.value)


def test_thresholds():
    assert pick(39, 39, 340) == THREAD                      # chesapeake: launch-latency regime
    assert pick(1 << 20, 1 << 20, 1 << 25, 1024) == MERGE   # BASELINE config 2
    assert pick(1 << 20, 1 << 20, 1 << 21, 12) == THREAD    # 2 nonzeros per row, no heavy row
    assert pick(1 << 20, 1 << 20, 1 << 21, 100000) == MERGE  # same sizes with a hub row
    assert pick(1 << 20, 1 << 20, 1 << 21) == MERGE          # widest row unknown: sizes only
    assert pick(0, 0, 0) == THREAD


def test_rejects_bad_arguments():
    out = C.c_int32(0)
    lib = _lib.load()
    assert lib.loopsb_select_schedule(-1, 4, 4, -1, C.byref(out)) == _lib.ERR_INVALID
    assert lib.loopsb_select_schedule(4, 4, 4, -1, None) == _lib.ERR_INVALID


def test_agrees_with_the_reference_heuristic_on_its_4831_matrices():
    """With the widest row unknown only nnz decides, as in the reference: its `kernel`
    column is merge-path exactly when nnz >= 10,000 on all but 4 matrices."""
    z = np.load(os.path.join(GOLDEN, "heuristics_ref.npz"))
    keep = (z["rows"] < 2**31) & (z["cols"] < 2**31)
    ours = np.array([pick(int(r), int(c), int(n)) for r, c, n in zip(z["rows"][keep], z["cols"][keep], z["nnz"][keep])])
    ref_merge = z["picked"][keep] == MERGE
    agree = float(((ours == MERGE) == ref_merge).mean())
    assert agree >= 0.999, agree
