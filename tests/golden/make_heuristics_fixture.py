"""tests/golden/make_heuristics_fixture.py -- run in the build container (the
reference is mounted there): turns the reference's published heuristic outcomes
(plots/data/heuristics.csv: one row per SuiteSparse matrix, `kernel` = what its
heuristic picked, `oracle-speed-up-kernel` = what was actually fastest) into the
small fixture tests/golden/heuristics_ref.npz used by tests/test_select_schedule.py."""
import csv
import os
import sys

import numpy as np

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/plots/data/heuristics.csv"
names = {"merge-path": 0, "thread-mapped": 2, "group-mapped": 3}   # loopsb schedule ids
rows, cols, nnz, picked, fastest = [], [], [], [], []
with open(src) as f:
    for r in csv.DictReader(f):
        rows.append(int(r["rows"])); cols.append(int(r["cols"])); nnz.append(int(r["nnzs"]))
        picked.append(names[r["kernel"]]); fastest.append(names[r["oracle-speed-up-kernel"]])
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "heuristics_ref.npz")
np.savez_compressed(out, rows=np.array(rows, np.int64), cols=np.array(cols, np.int64), nnz=np.array(nnz, np.int64),
                    picked=np.array(picked, np.int8), fastest=np.array(fastest, np.int8))
print(out, len(rows), "matrices")
