"""CPU part of the drop-in proof: the reference's own HOST-ONLY unit tests (layout contract for
every in-tree layout, flat partitioner, range / math helpers, format round trips, the Matrix
Market loader), compiled unchanged against this repo's include/ tree by
tests/cpp/reference_suite.mk, run here without a GPU. The device-side ones run under `-m gpu`
(tests/test_gpu_reference_suite.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, "tests", "_refsuite")
HOST_ONLY = ["test_layout_bcsr", "test_layout_coo", "test_layout_csc", "test_layout_csr", "test_layout_dia",
             "test_layout_ell", "test_layout_flat_partitioner", "test_util_math", "test_util_range",
             "test_format_round_trip", "test_market_loader"]


@pytest.mark.parametrize("name", HOST_ONLY)
def test_reference_host_unit_test(name):
    b = os.path.join(SUITE, "unit." + name)
    if not os.path.exists(b):
        pytest.skip("tests/_refsuite/ not built (the reference is not mounted where the build ran)")
    p = subprocess.run([b], capture_output=True, text=True, timeout=300)
    last = (p.stdout.strip().splitlines() or ["<no output>"])[-1]
    assert p.returncode == 0 and "failed: 0" in last, (p.stdout[-1500:], p.stderr[-1500:])
