#!/bin/bash
# tools/gpu_tiled.sh -- parity of the band-tiled kernel, then the geometry sweep on config 2.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tiled.py -x -q --timeout 120 > gpurun_out/pytest_tiled.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_tiled.log)"
grep -E "FAILED|Error|assert|Timeout" gpurun_out/pytest_tiled.log | head -20
timeout 600 python tools/tiled_sweep.py "$@" > gpurun_out/tiled_sweep.log 2>&1; echo "sweep rc=$?"
grep -v '^{' gpurun_out/tiled_sweep.log | tail -20
