"""CPU: pins the two oracle restatements against the golden fixtures (outputs of the reference's own
train.py / models.py, see tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from _golden import CASES, Golden
from oracle import invpref_numpy as on
from oracle import invpref_torch_cpu as ot


def nerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def _hyper(g, alpha):
    return on.Hyper(alpha=alpha, lr=g.lr, use_class_rw=g.crw, use_rec_rw=g.rrw, **g.coef)


@pytest.mark.parametrize("case", CASES)
def test_numpy_oracle_step0(case):
    g = Golden(case)
    B = min(g.B, g.N)
    d = g.data
    flags = on.Flags(g.implicit, g.roe, g.ree)
    e0, w0 = g["envs0"][:B].astype(np.int64), g["sample_weights0"][:B]
    # forward, fp32
    p32 = on.params_from_state_dict(g.group("init"), np.float32)
    s_inv, s_env, logp, _ = on.forward(p32, d[:B, 0], d[:B, 1], e0, flags, np.float32)
    assert nerr(s_inv, g["fwd0/s_inv"]) <= 1e-5 and nerr(s_env, g["fwd0/s_env"]) <= 1e-5
    assert nerr(logp, g["fwd0/logp"]) <= 1e-5
    # gradients, fp64: must equal the reference's fp64 autograd to rounding
    p64 = on.params_from_state_dict(g.group("init"), np.float64)
    st = on.new_adam_state(p64, np.float64)
    lo, gr = on.train_step(p64, st, d[:B, 0], d[:B, 1], d[:B, 2], e0, w0, _hyper(g, float(g["alpha0"])), flags,
                           np.float64)
    g64 = g.group("grad0_f64")
    for k, sk in on.STATE_KEYS.items():
        assert nerr(gr[k], g64[sk]) <= 1e-12, k
    for j, k in enumerate(on.LOSS_KEYS):
        assert abs(float(lo[k]) - g["epoch_losses"][0][j]) <= 1e-5 * abs(g["epoch_losses"][0][j]), k
    # Adam state after one step: m = 0.1 g, v = 0.001 g^2 -> tight; params within the fp32 noise
    m1, v1 = g.group("step1_m"), g.group("step1_v")
    for k, sk in on.STATE_KEYS.items():
        assert nerr(st["m"][k], m1[sk]) <= 2e-5, k
        assert nerr(st["v"][k], v1[sk]) <= 4e-5, k


@pytest.mark.parametrize("case", CASES)
def test_numpy_oracle_epoch_fp32(case):
    g = Golden(case)
    d = g.data
    flags = on.Flags(g.implicit, g.roe, g.ree)
    p = on.params_from_state_dict(g.group("init"), np.float32)
    st = on.new_adam_state(p, np.float32)
    bounds = on.mini_batch_bounds(g.N, g.B)
    losses = []
    for bi, (lo, hi) in enumerate(bounds):
        alpha = g.alpha if g.alpha is not None else on.alpha_schedule(bi, 0, len(bounds))
        ld, _ = on.train_step(p, st, d[lo:hi, 0], d[lo:hi, 1], d[lo:hi, 2], g["envs0"][lo:hi].astype(np.int64),
                              g["sample_weights0"][lo:hi], _hyper(g, alpha), flags, np.float32)
        losses.append([float(ld[k]) for k in on.LOSS_KEYS])
    assert np.abs(np.asarray(losses) - g["epoch_losses"]).max() <= 1e-4 * np.abs(g["epoch_losses"]).max()
    assert np.allclose(np.mean(np.asarray(losses), axis=0), g["epoch_mean_loss"], rtol=1e-4)
    for k, sk in on.STATE_KEYS.items():
        assert nerr(p[k], g.group("epoch1")[sk]) <= 5e-4, k


@pytest.mark.parametrize("case", CASES)
def test_numpy_oracle_cluster_and_stat(case):
    g = Golden(case)
    d = g.data
    flags = on.Flags(g.implicit, g.roe, g.ree)
    assert np.array_equal(on.init_eps(g.K), g["eps_table"])          # train.py:763-769, bit-exact
    sd = g.group("epoch1")
    sd.update(g.group("sep"))
    p = on.params_from_state_dict(sd, np.float32)
    new, dist = on.cluster_batch(p, d[:, 0], d[:, 1], d[:, 2], flags, g["sep_perm_idx"].astype(np.int64),
                                 g["eps_table"])
    mism = new != g["sep_envs"]
    assert not (mism & ~on.near_tie_mask(dist)).any()
    assert mism.sum() <= int(g["sep_near_ties"]) + 2
    cnt, cw, sw = on.stat_envs(g["cluster_envs"].astype(np.int64), g.K, g.N)
    assert np.array_equal(cnt, g["stat_counts"])
    assert np.array_equal(cw, g["stat_class_weights"])
    cnt0, cw0, sw0 = on.stat_envs(g["envs0"].astype(np.int64), g.K, g.N)
    assert np.array_equal(cw0, g["class_weights0"]) and np.array_equal(sw0, g["sample_weights0"])


@pytest.mark.parametrize("case", CASES)
def test_torch_cpu_port_matches_reference(case):
    """The torch-eager port issues the reference's op sequence, so it reproduces the fixtures closely."""
    g = Golden(case)
    d = g.data
    flags = on.Flags(g.implicit, g.roe, g.ree)
    torch.set_num_threads(8)
    P = ot.params_from_state_dict(g.group("init"))
    tr = ot.CpuTrainer(P, flags, _hyper(g, 0.0))
    assert torch.equal(tr.eps_table, torch.tensor(g["eps_table"]))
    bounds = on.mini_batch_bounds(g.N, g.B)
    u, i = torch.tensor(d[:, 0]), torch.tensor(d[:, 1])
    y = torch.tensor(d[:, 2]).float()
    e, w = torch.tensor(g["envs0"].astype(np.int64)), torch.tensor(g["sample_weights0"])
    losses = []
    for bi, (lo, hi) in enumerate(bounds):
        alpha = g.alpha if g.alpha is not None else on.alpha_schedule(bi, 0, len(bounds))
        ld = tr.train_a_batch(u[lo:hi], i[lo:hi], y[lo:hi], e[lo:hi], w[lo:hi], alpha)
        losses.append([ld[k] for k in on.LOSS_KEYS])
        if bi == 0:
            for k, sk in on.STATE_KEYS.items():
                assert nerr(P[k].detach().numpy(), g.group("step1")[sk]) <= 1e-6, k
    assert np.abs(np.asarray(losses) - g["epoch_losses"]).max() <= 1e-5 * np.abs(g["epoch_losses"]).max()
    for k, sk in on.STATE_KEYS.items():
        assert nerr(P[k].detach().numpy(), g.group("epoch1")[sk]) <= 1e-4, k
    # cluster on the separated parameters
    sd = g.group("epoch1")
    sd.update(g.group("sep"))
    tr2 = ot.CpuTrainer(ot.params_from_state_dict(sd), flags, _hyper(g, 0.0))
    new = []
    for lo, hi in bounds:
        new.append(tr2.cluster_a_batch(u[lo:hi], i[lo:hi], y[lo:hi], torch.tensor(g["sep_perm_idx"][lo:hi].astype(np.int64))))
    new = torch.cat(new).numpy()
    assert (new != g["sep_envs"]).sum() <= int(g["sep_near_ties"]) + 2
    cnts, cw, sw = ot.CpuTrainer.stat_envs(torch.tensor(g["cluster_envs"].astype(np.int64)), g.K, g.N)
    assert [cnts[k] for k in range(g.K)] == g["stat_counts"].tolist()
    assert np.array_equal(cw.numpy(), g["stat_class_weights"])


def test_alpha_schedule_and_batches():
    # train.py:891-894: p in [1, 2) so alpha is ~1
    assert abs(on.alpha_schedule(0, 0, 3) - (2.0 / (1.0 + np.exp(-10.0)) - 1.0)) < 1e-15
    assert on.alpha_schedule(2, 4, 3) == 2.0 / (1.0 + np.exp(-10.0 * (17.0 / 15.0))) - 1.0
    assert on.mini_batch_bounds(10, 4) == [(0, 4), (4, 8), (8, 10)]
    assert on.mini_batch_bounds(8, 4) == [(0, 4), (4, 8)]
    perm, rows, off = on.stable_segments(np.array([3, 1, 3, 0, 1, 3]))
    assert perm.tolist() == [3, 1, 4, 0, 2, 5] and rows.tolist() == [0, 1, 3] and off.tolist() == [0, 1, 3, 6]
