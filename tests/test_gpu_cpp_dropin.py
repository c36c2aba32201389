"""GPU: the C++ side of the boundary. tests/cpp/dropin_spmv.cu is user-style
code written against the reference's API (containers, algorithms::spmv::*,
kernels over schedule::setup<> incl. a custom layout and the reference's own
merge-path kernel body); it is compiled against include/loops/ by
__graft_entry__.build() and must report zero failures."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_dropin_program():
    exe = os.path.join(ROOT, "tests", "cpp", "dropin_spmv")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp/dropin_spmv not built (run __graft_entry__.build())")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(p.stdout, p.stderr)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith(("OK", "FAIL"))]
    assert len(lines) == 25 and all(l.startswith("OK") for l in lines), p.stdout
