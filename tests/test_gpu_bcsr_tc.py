"""GPU parity of the BCSR 4x4 bf16 tcgen05 path (BASELINE config 4) against the
oracle's fp32-accumulating restatement of the reference kernel's loop
(algorithms/spmv/bcsr_thread_mapped.cuh:48-62) on bf16-rounded inputs."""
import numpy as np
import pytest
import torch

from helpers import random_csr

pytestmark = pytest.mark.gpu


def _run(oracle, rows, cols, off, idx, val, x, exact):
    from loops_b200 import csr_t, bcsr_t
    from loops_b200.algorithms import spmv
    A = csr_t(rows, cols, off, idx, val)
    B = bcsr_t.from_csr(A, 4, 4, value_dtype=torch.bfloat16)
    xb = B.padded_x(torch.as_tensor(x).cuda().to(torch.bfloat16))
    y = torch.full((rows,), float("nan"), device="cuda")
    spmv.bcsr_thread_mapped(B, xb, y)
    y = y.cpu().numpy()
    # oracle on the SAME bf16-rounded numbers, fp32 accumulate
    b_off = B.block_offsets.cpu().numpy()
    b_col = B.block_col_indices.cpu().numpy()
    b_val = B.values.float().cpu().numpy()
    xr = xb.float().cpu().numpy()
    ref = oracle.spmv_bcsr(4, 4, rows, b_off, b_col, b_val, xr)
    assert np.all(np.isfinite(y))
    if exact:
        np.testing.assert_array_equal(y, ref)
    else:
        # products of two bf16 are exact in fp32; only the order of the fp32
        # adds differs -> 1e-6 relative to the row's L1 mass
        l1 = np.zeros(rows, np.float64)
        nbr = len(b_off) - 1
        for br in range(nbr):
            for b in range(b_off[br], b_off[br + 1]):
                blk = np.abs(b_val[b * 16:(b + 1) * 16].reshape(4, 4).astype(np.float64))
                xs = np.abs(xr[b_col[b] * 4: b_col[b] * 4 + 4].astype(np.float64))
                for i in range(4):
                    if br * 4 + i < rows:
                        l1[br * 4 + i] += float(blk[i] @ xs)
        err = np.abs(y.astype(np.float64) - ref.astype(np.float64)) / np.maximum(l1, 1e-30)
        assert err.max() <= 1e-6, float(err.max())


def test_battery_bf16(oracle, battery):
    for b in battery:
        _run(oracle, b["rows"], b["cols"], b["off"], b["idx"], b["val"], b["x"], exact=False)


@pytest.mark.parametrize("rows,cols,dens,seed,empty,heavy", [
    (1, 7, 0.9, 1, 0, None), (130, 130, 0.05, 2, 0, None), (257, 1000, 0.01, 3, 3, (100, 1000)),
    (4096, 4096, 0.004, 4, 0, (7, 4096)), (1000, 64, 0.0, 5, 0, (500, 64)), (333, 4001, 0.02, 6, 2, None)])
def test_shapes_exact(oracle, rows, cols, dens, seed, empty, heavy):
    off, idx, val = random_csr(rows, cols, dens, seed, empty, heavy, exact=True)
    x = np.random.default_rng(seed).integers(1, 11, cols).astype(np.float32)
    _run(oracle, rows, cols, off, idx, val, x, exact=True)


def test_powerlaw_blocks_exact(oracle):
    """Config-4-shaped input at 1/64 scale: power-law block-row lengths."""
    from loops_b200 import generate as g
    nbr = 4096
    off, bcol, _ = g.synth_csr(nbr, nbr, nbr * 32)         # block structure
    off, bcol = off.numpy(), bcol.numpy()
    rows = cols = nbr * 4
    # expand every block to a dense 4x4 of k/8 values -> CSR
    nb = len(bcol)
    rng = np.random.default_rng(9)
    vals = (rng.integers(1, 17, size=(nb, 4, 4)) / 8.0).astype(np.float32)
    deg = np.diff(off)
    csr_off = np.zeros(rows + 1, np.int64)
    csr_off[1:] = np.cumsum(np.repeat(deg * 4, 4))
    idx = np.empty(nb * 16, np.int32); val = np.empty(nb * 16, np.float32)
    p = 0
    for br in range(nbr):
        blks = np.arange(off[br], off[br + 1])
        for i in range(4):
            n = len(blks) * 4
            idx[p:p + n] = (bcol[blks][:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
            val[p:p + n] = vals[blks, i, :].reshape(-1)
            p += n
    x = g.x_recipe(cols).numpy()
    _run(oracle, rows, cols, csr_off.astype(np.int32), idx, val, x, exact=True)


def test_fp32_blocks_still_use_thread_mapped(oracle, battery):
    from loops_b200 import csr_t, bcsr_t
    from loops_b200.algorithms import spmv
    b = battery[3]
    A = csr_t(b["rows"], b["cols"], b["off"], b["idx"], b["val"])
    B = bcsr_t.from_csr(A, 4, 4)
    y = torch.empty(b["rows"], device="cuda")
    spmv.bcsr_thread_mapped(B, B.padded_x(torch.as_tensor(b["x"]).cuda()), y)
    g_off, g_col, g_val = b["bcsr4"]
    xp = np.zeros(((b["cols"] + 3) // 4) * 4, np.float32); xp[: b["cols"]] = b["x"]
    np.testing.assert_array_equal(y.cpu().numpy(), oracle.spmv_bcsr(4, 4, b["rows"], g_off, g_col, g_val, xp))


def test_packed_and_direct_paths_are_bit_identical(monkeypatch):
    """The plan's packed copy (A-tile images + block columns, one TMA bulk copy per
    K-step, dynamic work items) and the kernel that reads the BCSR arrays directly
    run the same MMAs in the same order: y must be equal bit for bit -- also after
    the values change in place and the copy is rebuilt."""
    from loops_b200 import csr_t, bcsr_t
    from loops_b200.algorithms import spmv
    rows, cols = 4103, 3001                     # ragged last block-row / block-column
    off, idx, val = random_csr(rows, cols, 0.012, 77, empty_every=6, heavy_row=(9, 2500))
    A = csr_t(rows, cols, off, idx, val)
    B = bcsr_t.from_csr(A, 4, 4, value_dtype=torch.bfloat16)
    xb = B.padded_x(torch.as_tensor(np.random.default_rng(5).uniform(-1, 1, cols).astype(np.float32)).cuda()
                    .to(torch.bfloat16))
    yd = torch.full((rows,), float("nan"), device="cuda")
    yp = torch.full((rows,), float("nan"), device="cuda")
    monkeypatch.setenv("LOOPSB_BCSR_PACKED", "0")
    spmv.bcsr_thread_mapped(B, xb, yd)
    monkeypatch.setenv("LOOPSB_BCSR_PACKED", "1")
    spmv.bcsr_thread_mapped(B, xb, yp)
    assert torch.equal(yd, yp) and bool(torch.isfinite(yp).all())
    B.values.mul_(0.5)                          # in place: same pointer, new numbers
    spmv.bcsr_thread_mapped(B, xb, yp, repack=True)
    assert torch.equal(yp, yd * 0.5)
    for _ in range(3):                          # the work counter is re-armed every launch
        spmv.bcsr_thread_mapped(B, xb, yp)
    assert torch.equal(yp, yd * 0.5)
    B.values.mul_(4.0)                          # in place WITHOUT repack=True: torch's version counter is watched
    spmv.bcsr_thread_mapped(B, xb, yp)
    assert torch.equal(yp, yd * 2.0)


def test_full_size_config4_exact():
    """BASELINE configs[3] at FULL size -- 262,144 block-rows / 8,388,608 4x4 bf16 blocks (the workload
    tools/bcsr_bench.py and bench.py's `extra` time): y equal bit for bit to a float64 evaluation with
    torch ops (values k/8, x integers 1..10: every product and sum is exact), on the first launch (fully
    ordered) and after several back-to-back launches (programmatic dependent launches, alternating work
    counters)."""
    from loops_b200 import generate as g
    from loops_b200.algorithms import spmv
    from loops_b200.container import bcsr_t
    nbr = 1 << 18
    nb = nbr * 32
    b_off, b_col, _ = g.synth_csr(nbr, nbr, nb, device="cuda")
    e = torch.arange(nb * 16, device="cuda", dtype=torch.int64)
    b_val = (((g._lsr(g.mix64(e ^ 0x5151), 33) % 16) + 1).to(torch.float32) / 8.0).to(torch.bfloat16)
    del e
    B = bcsr_t.from_tensors(4, 4, nbr * 4, nbr * 4, nb * 16, b_off, b_col, b_val)
    xb = g.x_recipe(nbr * 4, device="cuda").to(torch.bfloat16)
    yb = torch.full((nbr * 4,), float("nan"), device="cuda")
    spmv.bcsr_thread_mapped(B, xb, yb)
    rowb = torch.repeat_interleave(torch.arange(nbr, device="cuda"), (b_off[1:] - b_off[:-1]).long(), output_size=nb)
    xs = xb.double()[(b_col.long()[:, None] * 4 + torch.arange(4, device="cuda")[None, :])]      # [nb, 4]
    prod = (b_val.double().view(nb, 4, 4) * xs[:, None, :]).sum(2)                                  # [nb, 4]
    ref = torch.zeros(nbr, 4, dtype=torch.float64, device="cuda").index_add_(0, rowb, prod).view(-1)
    del xs, prod, rowb
    assert torch.equal(yb.double(), ref)
    for _ in range(5):
        yb.fill_(float("nan"))
        spmv.bcsr_thread_mapped(B, xb, yb, sync=False)
    torch.cuda.synchronize()
    assert torch.equal(yb.double(), ref)
