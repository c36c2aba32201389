"""CPU, world_size 2 over gloo: the host logic of the row-partitioned path
(shard boundaries, offset rebasing, the single all-gather of x, y sharding).
The per-shard arithmetic is checked with the oracle -- the product kernels need
a GPU and have no CPU fallback."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from helpers import Oracle
        from loops_b200 import generate as g
        from loops_b200.dist import row_range, shard_csr, nnz_imbalance
        rows = cols = 4096
        nnz = rows * 16
        off, idx, val = (t.numpy() for t in g.synth_csr(rows, cols, nnz))
        x = g.x_recipe(cols)
        oracle = Oracle(os.path.join(ROOT, "oracle", "libloops_oracle.so"))
        y_ref = oracle.spmv(off, idx, val, x.numpy())

        r0, r1 = row_range(rows, rank, world)
        l_off, l_idx, l_val = shard_csr(off, idx, val, rank, world)
        assert l_off[0] == 0 and l_off[-1] == len(l_idx) == off[r1] - off[r0]
        # the generator's own sharding produces the same shard
        s_off, s_idx, s_val = g.synth_csr(rows, cols, nnz, row_begin=r0, row_end=r1)
        np.testing.assert_array_equal(s_off.numpy(), l_off)
        np.testing.assert_array_equal(s_idx.numpy(), l_idx)

        # the one collective: all-gather of the dense x shards
        x_shard = x[r0:r1].clone()
        x_full = torch.empty(cols)
        dist.all_gather_into_tensor(x_full, x_shard)
        assert torch.equal(x_full, x)

        y_shard = oracle.spmv(l_off, np.ascontiguousarray(l_idx), np.ascontiguousarray(l_val), x_full.numpy())
        np.testing.assert_array_equal(y_shard, y_ref[r0:r1])

        # the overlap split: own-column part (needs only x_shard) + the rest (needs x_full)
        from loops_b200.dist import split_columns
        c0, c1 = row_range(cols, rank, world)
        (o_off, o_idx, o_val), (r_off, r_idx, r_val) = split_columns(
            torch.from_numpy(l_off), torch.from_numpy(np.ascontiguousarray(l_idx)),
            torch.from_numpy(np.ascontiguousarray(l_val)), c0, c1)
        assert int(o_off[-1]) + int(r_off[-1]) == len(l_idx)
        assert o_idx.numel() == 0 or (int(o_idx.min()) >= 0 and int(o_idx.max()) < c1 - c0)
        y_own = oracle.spmv(o_off.numpy(), o_idx.numpy(), o_val.numpy(), x_shard.numpy())
        y_rest = oracle.spmv(r_off.numpy(), r_idx.numpy(), r_val.numpy(), x_full.numpy())
        np.testing.assert_array_equal(y_own + y_rest, y_ref[r0:r1])      # exact inputs: any order of the adds

        # y shards concatenate to the global y (and are the next x shards)
        y_all = torch.empty(rows)
        dist.all_gather_into_tensor(y_all, torch.from_numpy(y_shard))
        np.testing.assert_array_equal(y_all.numpy(), y_ref)
        assert nnz_imbalance(off, world) < 1.1
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_row_partition_world2_gloo():
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "libloops_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_row_range_covers_everything():
    from loops_b200.dist import row_range
    for rows in (0, 1, 7, 1 << 20, (1 << 24)):
        for world in (1, 2, 4, 8):
            cuts = [row_range(rows, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == rows
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))


def test_ring_blocks_and_column_split(oracle):
    """Host logic of the phased all-gather (loops_b200/csrc/dist.cu): every rank's block map
    covers each remote chunk once in ring order, and the column blocks add up to the shard."""
    from loops_b200 import generate as g
    from loops_b200.dist import ring_blocks, split_column_blocks, shard_csr
    for world, groups in ((2, [1]), (4, [1, 2]), (8, [2, 2, 3]), (8, [7])):
        for rank in range(world):
            m = ring_blocks(world, rank, groups)
            assert m[rank] == 0 and sorted(m).count(0) == 1
            order = [m[(rank + k) % world] for k in range(1, world)]
            assert order == sorted(order)                       # blocks arrive in ring order
            assert [order.count(b + 1) for b in range(len(groups))] == groups
    with pytest.raises(ValueError):
        ring_blocks(4, 0, [1, 1])
    rows = cols = 2048
    off, idx, val = (t.numpy() for t in g.synth_csr(rows, cols, rows * 12))
    x = g.x_recipe(cols).numpy()
    world, rank = 4, 1
    l_off, l_idx, l_val = shard_csr(off, idx, val, rank, world)
    y = np.zeros(len(l_off) - 1, np.float32)
    total = 0
    for b_off, b_idx, b_val in split_column_blocks(l_off, l_idx, l_val, cols // world, ring_blocks(world, rank, [1, 2])):
        y += oracle.spmv(b_off, b_idx, b_val, x)               # exact inputs: any grouping is exact
        total += len(b_idx)
    assert total == len(l_idx)
    np.testing.assert_array_equal(y, oracle.spmv(l_off, l_idx, l_val, x))
