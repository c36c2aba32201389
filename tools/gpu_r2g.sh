#!/bin/bash
mkdir -p gpurun_out
for V in 0 2; do
  echo "== variant $V"
  LOOPSB_MERGE2_VARIANT=$V PROBE_NOACC=1 python tools/block_probe.py 8 3,4 2>&1 | sed 's/^/noacc /'
  LOOPSB_MERGE2_VARIANT=$V PROBE_NOACC=1 python tools/block_probe.py 8 7 2>&1 | sed 's/^/noacc /'
done
LOOPSB_MERGE2_VARIANT=0 PROBE_NOACC=1 python tools/block_probe.py 4 1,2 2>&1 | sed 's/^/noacc /'
LOOPSB_MERGE2_VARIANT=2 PROBE_NOACC=1 python tools/block_probe.py 4 1,2 2>&1 | sed 's/^/noacc /'
