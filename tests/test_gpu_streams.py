"""GPU parity of the schedule iterators: loopsb_emit_schedule (the loops-b200
schedule::setup<> classes) vs the oracle, and -- when the reference-built
recorder travels with the snapshot -- vs the reference's own templates run on
the same GPU. Bit-exact integer comparison."""
import ctypes as C

import numpy as np
import pytest

from helpers import Oracle, SCHED, P

pytestmark = pytest.mark.gpu

FIELDS = ("visitor", "step", "tile", "visits", "commit", "map", "dense_tile", "dense_atom", "dense_emit",
          "thread_start")


def _layouts(b):
    from loops_b200 import layout
    import torch
    off_d = torch.as_tensor(b["off"]).cuda()
    nnz = len(b["idx"])
    return [
        ("csr", layout.csr(off_d, b["rows"], nnz), Oracle.csr_layout(b["off"])),
        ("coo", layout.coo(nnz), Oracle.coo_layout(nnz)),
        ("ell", layout.ell(b["rows"], b["ell_pitch"]), Oracle.ell_layout(b["rows"], b["ell_pitch"])),
    ]


def _compare(mine, theirs, label):
    for f in FIELDS:
        a = getattr(mine, f, None) if not isinstance(mine, dict) else mine.get(f)
        b = theirs.get(f)
        if a is None or b is None:
            continue
        np.testing.assert_array_equal(np.asarray(a).ravel(), np.asarray(b).ravel(), err_msg=f"{label}:{f}")


CELLS = [("thread_mapped", 2, 32, 1), ("thread_mapped", 5, 128, 1), ("group_mapped", 0, 128, 1),
         ("work_oriented", 1, 128, 1), ("work_oriented", 3, 128, 1),
         ("merge_path_flat", 0, 128, 8), ("merge_path_flat", 0, 128, 7), ("merge_path_flat", 0, 128, 5)]


@pytest.mark.parametrize("sched,grid,tpb,ipt", CELLS)
def test_emit_matches_oracle_on_battery(oracle, battery, sched, grid, tpb, ipt):
    from loops_b200.emit import emit_schedule
    for b in battery:
        for lname, lay, olay in _layouts(b):
            mine = emit_schedule(lay, SCHED[sched], grid, tpb, ipt)
            theirs = oracle.emit(olay, SCHED[sched], grid, tpb, ipt)
            _compare(mine, theirs, (b["name"], lname, sched))
            assert np.all(mine.visits == 1)


@pytest.mark.parametrize("sched,grid,tpb,ipt", [("thread_mapped", 512, 128, 1), ("group_mapped", 0, 128, 1),
                                                ("work_oriented", 296, 128, 1), ("merge_path_flat", 0, 128, 8)])
def test_emit_matches_oracle_64k_rows(oracle, sched, grid, tpb, ipt):
    """SURVEY 8d parity gate: index-stream equality on a 64K-row sample of the
    benchmark generator (preprocess coordinates materialised: 2113 merge tiles)."""
    import torch
    from loops_b200 import generate as g, layout
    from loops_b200.emit import emit_schedule
    rows = 1 << 16
    off, _, _ = g.synth_csr(rows, rows, rows * 32)
    lay = layout.csr(off.cuda(), rows, rows * 32)
    mine = emit_schedule(lay, SCHED[sched], grid, tpb, ipt)
    theirs = oracle.emit(Oracle.csr_layout(off.numpy()), SCHED[sched], grid, tpb, ipt)
    _compare(mine, theirs, ("synthetic-64k", sched))


def test_plan_coordinates_match_oracle(oracle, battery):
    """loopsb_plan_merge_coords_host == the reference's
    generate_search_coordinates values (merge_path_flat.hxx:45-76)."""
    import torch
    from loops_b200 import csr_t, generate as g, _lib
    rows = 1 << 16
    off, idx, val = g.synth_csr(rows, rows, rows * 32)
    A = csr_t(rows, rows, off.numpy(), idx.numpy(), val.numpy())
    plan = A.plan(_lib.SCHED_MERGE_PATH_FLAT)
    torch.cuda.synchronize()
    theirs = oracle.emit(Oracle.csr_layout(off.numpy()), SCHED["merge_path_flat"], 0, 128, 8)
    np.testing.assert_array_equal(plan.merge_coords(), theirs["coords"])
    info = plan.info()
    assert info.num_merge_tiles == (rows + rows * 32 + 1023) // 1024
    assert info.launches_per_spmv == 2 and info.grid_blocks > 0


def _ref_emit(G, kind, off, T, A, pitch, sched, grid, tpb, ipt):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import ref_gpu_emit
    return ref_gpu_emit(G, kind, off, T, A, pitch, sched, grid, tpb, ipt)


@pytest.mark.parametrize("sched,grid,tpb,ipt", CELLS)
def test_emit_matches_reference_templates_on_gpu(ref_gpu, oracle, battery, sched, grid, tpb, ipt):
    """The reference's schedule::setup<> classes, run right here, must hand out
    the same (thread, step, tile, atom) stream as ours AND the oracle's."""
    from loops_b200.emit import emit_schedule
    for b in battery:
        nnz = len(b["idx"])
        for (lname, lay, olay), kind in zip(_layouts(b), (0, 1, 2)):
            T = b["rows"] if kind != 1 else nnz
            g = grid
            if g == 0:
                g = (T + 127) // 128 if sched == "group_mapped" else 1
            ref = _ref_emit(ref_gpu, kind, b["off"] if kind == 0 else None, T, nnz if kind != 2 else 0,
                            b["ell_pitch"] if kind == 2 else (1 if kind == 1 else 0), SCHED[sched], g, tpb, ipt)
            mine = emit_schedule(lay, SCHED[sched], grid, tpb, ipt)
            _compare(mine, ref, (b["name"], lname, sched, "ours-vs-reference"))
            _compare(oracle.emit(olay, SCHED[sched], grid, tpb, ipt), ref,
                     (b["name"], lname, sched, "oracle-vs-reference"))


def test_reference_kernels_agree_on_y(ref_gpu, oracle, battery):
    """The reference's own SpMV kernels (the kernels to beat) produce the same
    y as ours within the tolerance, on its battery."""
    import torch
    from loops_b200 import csr_t
    from loops_b200.algorithms import spmv
    for b in battery:
        for which, name in enumerate(("merge_path_flat", "work_oriented", "thread_mapped", "group_mapped")):
            yr = np.zeros(b["rows"], np.float32)
            ms, inner = C.c_float(), C.c_float()
            rc = ref_gpu.ref_gpu_spmv(which, b["rows"], b["cols"], len(b["idx"]), P(b["off"]), P(b["idx"]),
                                      P(b["val"]), P(b["x"]), P(yr), 1, C.byref(ms), C.byref(inner))
            assert rc == 0
            A = csr_t(b["rows"], b["cols"], b["off"], b["idx"], b["val"])
            y = torch.empty(b["rows"], device="cuda")
            spmv.BY_NAME[name](A, torch.as_tensor(b["x"]).cuda(), y)
            np.testing.assert_allclose(y.cpu().numpy(), yr, rtol=1e-5, atol=1e-5, err_msg=(b["name"], name))
