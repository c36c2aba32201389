"""The drop-in proof, run on the GPU: the REFERENCE's own unit tests (unittests/test_*.cu) and
example mains (examples/spmv/*.cu as .f32 and .f64, examples/{range,saxpy,spmm}), compiled
UNCHANGED against this repo's include/ tree and libloopsb200.so by tests/cpp/reference_suite.mk
(in the build container, where /root/reference is mounted; the binaries travel in
tests/_refsuite/). Every Catch2 TEST_CASE must pass; every example must print the reference's
--validate lines with `Errors: 0` on BASELINE config 1 (chesapeake: `39 x 39 (340)`,
site/content/experimentation.md:19-37) and a Wilkinson verdict of NOT_A_BUG."""
import glob
import os
import subprocess

import numpy as np
import pytest

from helpers import load_chesapeake, random_csr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, "tests", "_refsuite")


def _binaries(prefix):
    return sorted(p for p in glob.glob(os.path.join(SUITE, prefix + "*")) if os.access(p, os.X_OK))


def _need_suite():
    if not _binaries("unit."):
        pytest.skip("tests/_refsuite/ not built (the reference is not mounted where the build ran)")


def _write_mtx(path, rows, cols, off, idx, val, pattern=False):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate " + ("pattern" if pattern else "real") + " general\n")
        f.write(f"{rows} {cols} {len(idx)}\n")
        for r in range(rows):
            for a in range(off[r], off[r + 1]):
                f.write(f"{r + 1} {idx[a] + 1}" + ("" if pattern else f" {val[a]:.9g}") + "\n")


@pytest.fixture(scope="module")
def chesapeake_mtx(tmp_path_factory):
    c = load_chesapeake()
    p = tmp_path_factory.mktemp("mtx") / "chesapeake.mtx"
    _write_mtx(str(p), 39, 39, c["off"], c["idx"], c["val"], pattern=True)
    return str(p)


@pytest.fixture(scope="module")
def ragged_mtx(tmp_path_factory):
    off, idx, val = random_csr(700, 650, 0.02, seed=13, empty_every=9, heavy_row=(300, 600))
    p = tmp_path_factory.mktemp("mtx") / "ragged.mtx"
    _write_mtx(str(p), 700, 650, off, idx, val)
    return str(p)


def test_every_reference_unit_test_passes():
    _need_suite()
    bins = _binaries("unit.")
    assert len(bins) >= 26, bins      # one binary per unittests/test_*.cu
    summary = []
    for b in bins:
        p = subprocess.run([b], capture_output=True, text=True, timeout=600)
        last = (p.stdout.strip().splitlines() or ["<no output>"])[-1]
        summary.append((os.path.basename(b), p.returncode, last))
        assert p.returncode == 0, (os.path.basename(b), p.stdout[-1500:], p.stderr[-1500:])
        assert "failed: 0" in last, (os.path.basename(b), last)
    cases = sum(int(s[2].split("test cases:")[1].split("|")[0]) for s in summary)
    assert cases >= 90          # the reference has 92 TEST_CASEs


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_reference_examples_validate_on_config_1(chesapeake_mtx, ragged_mtx, precision):
    _need_suite()
    bins = _binaries("loops.spmv.")
    mine = [b for b in bins if b.endswith("." + precision)]
    assert len(mine) >= 13, mine
    for b in mine:
        for mtx, dims in ((chesapeake_mtx, "39 x 39 (340)"), (ragged_mtx, "700 x 650")):
            p = subprocess.run([b, "-m", mtx, "--validate", "--rigorous"], capture_output=True, text=True, timeout=600)
            name = os.path.basename(b)
            assert p.returncode == 0, (name, p.stdout[-1200:], p.stderr[-1200:])
            out = p.stdout
            assert "Dimensions:\t" + dims in out, (name, out)
            assert "Errors:\t\t0" in out, (name, out)
            assert "Verdict:\tNOT_A_BUG" in out, (name, out)


def test_other_reference_examples_run():
    _need_suite()
    for name in ("loops.range", "loops.saxpy"):
        b = os.path.join(SUITE, name)
        assert os.path.exists(b), name
        p = subprocess.run([b], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, (name, p.stdout[-800:], p.stderr[-800:])
