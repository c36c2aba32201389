"""CPU: synthetic generator and x recipe (host logic of the product)."""
import os

import numpy as np
import torch

from helpers import GOLDEN
from loops_b200 import generate as g


def test_x_recipe_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN, "xrecipe.npz"))
    for seed in (42, 7):
        np.testing.assert_array_equal(g.x_recipe(4096, 1, 10, seed).numpy(), z[f"int_1_10_seed{seed}"])
    h = g.hash32(torch.arange(1024, dtype=torch.int64)).numpy().astype(np.uint32)
    np.testing.assert_array_equal(h, z["hash_0_1023"])
    assert g.x_recipe(39).tolist()[:5] == [1, 10, 6, 2, 10]


def test_mix64_twins_agree():
    e = torch.arange(0, 5000, dtype=torch.int64) * 7919 - 12345
    a = g.mix64(e).numpy().astype(np.uint64)
    b = g.mix64_np(e.numpy().astype(np.uint64))
    np.testing.assert_array_equal(a, b)


def test_powerlaw_degrees_properties():
    deg = g.powerlaw_degrees(1 << 14, 1 << 19)
    assert deg.sum() == 1 << 19 and deg.min() >= 1 and deg.max() <= 1024
    assert deg.std() / deg.mean() > 1.0          # skewed, not uniform
    np.testing.assert_array_equal(deg, g.powerlaw_degrees(1 << 14, 1 << 19))   # deterministic
    tiny = g.powerlaw_degrees(10, 10)
    assert tiny.tolist() == [1] * 10


def test_synth_csr_is_canonical_and_shardable():
    rows, cols, nnz = 3000, 3000, 3000 * 24
    off, idx, val = g.synth_csr(rows, cols, nnz)
    off, idx, val = off.numpy(), idx.numpy(), val.numpy()
    assert off[0] == 0 and off[-1] == nnz and np.all(np.diff(off) >= 1)
    assert idx.min() >= 0 and idx.max() < cols
    for r in range(rows):
        assert np.all(np.diff(idx[off[r]:off[r + 1]]) > 0)      # unique + ascending
    assert set(np.unique(val * 8)).issubset(set(range(1, 17)))  # k/8, exact in bf16
    # shards concatenate to the whole (config 5 generates per rank)
    deg = g.powerlaw_degrees(rows, nnz, d_max=min(1024, cols))
    parts = [g.synth_csr(rows, cols, nnz, degrees=deg, row_begin=a, row_end=b)
             for a, b in ((0, 1000), (1000, 2100), (2100, 3000))]
    np.testing.assert_array_equal(np.concatenate([p[1].numpy() for p in parts]), idx)
    np.testing.assert_array_equal(np.concatenate([p[2].numpy() for p in parts]), val)
    assert [int(p[0][-1]) for p in parts] == [off[1000], off[2100] - off[1000], off[3000] - off[2100]]
