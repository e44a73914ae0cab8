"""CPU: the binary columnar interaction cache (dataloader.py) -- round trip, id / score
narrowing, staleness against the CSV it was made from, and loaders that give the same arrays with and without it."""
import os

import numpy as np
import torch

from invpref_kdd_2022_b200 import dataloader as dl


def _csv(path, data):
    with open(path, "w") as f:
        f.write("user_id,item_id,score\n")
        for u, i, s in data.tolist():
            f.write(f"{u},{i},{s}\n")


def test_round_trip_and_narrowing(tmp_path):
    rng = np.random.default_rng(0)
    for U, I, hi in ((300, 290, 6), (70000, 40000, 2), (5, 3_000_000, 6)):
        data = np.stack([rng.integers(0, U, 5000), rng.integers(0, I, 5000), rng.integers(0, hi, 5000)], axis=1)
        p = str(tmp_path / f"c{U}.bin")
        dl.write_interaction_cache(p, data)
        got = dl.read_interaction_cache(p)
        assert got.dtype == np.int64 and np.array_equal(got, data)
        per_row = (os.path.getsize(p) - 64) / 5000
        assert per_row <= (2 if U < 32768 else 4) + (2 if I < 32768 else 4) + 1 + 0.01
    # empty file, foreign file
    dl.write_interaction_cache(str(tmp_path / "e.bin"), np.zeros((0, 3), dtype=np.int64))
    assert dl.read_interaction_cache(str(tmp_path / "e.bin")).shape == (0, 3)
    (tmp_path / "junk.bin").write_bytes(b"x" * 100)
    assert dl.read_interaction_cache(str(tmp_path / "junk.bin")) is None
    assert dl.read_interaction_cache(str(tmp_path / "missing.bin")) is None


def test_loader_uses_and_invalidates_the_cache(tmp_path):
    rng = np.random.default_rng(1)
    tr = np.stack([rng.integers(0, 50, 400), rng.integers(0, 30, 400), rng.integers(1, 6, 400)], axis=1)
    te = np.stack([rng.integers(0, 50, 100), rng.integers(0, 30, 100), rng.integers(1, 6, 100)], axis=1)
    tr[0, :2] = (49, 29)
    d = tmp_path / "ds"
    d.mkdir()
    _csv(d / "train.csv", tr)
    _csv(d / "test.csv", te)
    cpu = torch.device("cpu")
    a = dl.ExplicitDataLoader(str(d), cpu)
    assert (d / "train.csv.invpref.bin").exists() and (d / "test.csv.invpref.bin").exists()
    b = dl.ExplicitDataLoader(str(d), cpu)                      # second construction reads the cache
    assert np.array_equal(a.train_data_np, tr) and np.array_equal(b.train_data_np, tr)
    assert np.array_equal(b.test_data_np, te) and (b.user_num, b.item_num) == (50, 30)
    # the CSV changes: the stale cache is ignored and rewritten
    tr2 = tr.copy()
    tr2[5, 2] = 5 if tr2[5, 2] != 5 else 4
    _csv(d / "train.csv", tr2)
    os.utime(d / "train.csv", ns=(1, 1_700_000_000_000_000_000))
    c = dl.ExplicitDataLoader(str(d), cpu)
    assert np.array_equal(c.train_data_np, tr2)
    assert np.array_equal(dl.read_interaction_cache(str(d / "train.csv.invpref.bin")), tr2)
    # implicit loader through the same reader
    imp = np.stack([rng.integers(0, 50, 300), rng.integers(0, 30, 300), rng.integers(0, 2, 300)], axis=1)
    _csv(d / "train.csv", imp)
    _csv(d / "test.csv", imp[imp[:, 2] > 0][:40])
    e = dl.YahooImplicitBCELossDataLoader(str(d), cpu)
    f = dl.YahooImplicitBCELossDataLoader(str(d), cpu)
    assert np.array_equal(e.train_data_np, imp) and np.array_equal(f.train_data_np, imp)
    assert e.all_test_users_by_sorted_list == f.all_test_users_by_sorted_list
