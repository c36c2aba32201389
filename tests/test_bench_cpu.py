"""CPU checks of bench.py (the driver-facing contract): the pieces that run without a GPU --
the workload/config blocks for every N the driver uses, the algorithmic-byte figures of
SURVEY 8d, and the reference arm end to end (it times the CPU path, so it runs here)."""
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_workload_config_for_every_n(monkeypatch):
    b = _bench()
    for groups_env in (None, "0", "1,1,1,2,2"):
        if groups_env is None:
            monkeypatch.delenv("LOOPSB_DIST_GROUPS", raising=False)
        else:
            monkeypatch.setenv("LOOPSB_DIST_GROUPS", groups_env)
        for n in (1, 2, 4, 8):
            cfg = b.workload_config(n)
            assert "workload" in cfg and "model" not in cfg
            assert cfg["partition"] == ("none" if n == 1 else f"row{n}")
            json.dumps(cfg)
    from loops_b200.dist import default_groups
    monkeypatch.delenv("LOOPSB_DIST_GROUPS", raising=False)
    for n in (2, 4, 8):
        g = default_groups(n)
        assert g == [] or sum(g) == n - 1               # no split, or phases covering every remote chunk once
    monkeypatch.setenv("LOOPSB_DIST_GROUPS", "0")
    assert default_groups(8) == []                       # one ncclAllGather, no split


def test_algorithmic_bytes_match_the_survey():
    b = _bench()
    assert b.algorithmic_bytes(1 << 20, 1 << 20, 1 << 25) == 281_018_372          # config 2
    assert b.algorithmic_bytes(1 << 24, 1 << 24, 1 << 29) == 4_496_293_892        # config 5, whole problem


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "csr_spmv_nnz_per_s" and line["unit"] == "nnz/s"
    assert line["n_gpus"] == 1 and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["e2e"]["value"] == line["value"]
    assert line["config"]["workload"].startswith("synthetic power-law CSR 2^20 rows / 2^25 nnz")
