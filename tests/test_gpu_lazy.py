"""GPU: lazy dense Adam for the user tables must be BIT-IDENTICAL to the plain dense path (same arithmetic in
the same order: skipped zero-gradient steps are replayed with each step's own bias corrections), for every
table, the Adam state, the losses and the env assignments -- including rows skipped for many steps, long
(chunked) user segments and the buffer/flush bookkeeping of the model / trainer classes."""
import numpy as np
import pytest
import torch

from _golden import Golden
from oracle import invpref_numpy as on

pytestmark = pytest.mark.gpu


def synth(U, I, N, K, D, implicit, seed):
    rng = np.random.default_rng(seed)
    u = np.floor(U * rng.random(N) ** 1.5).astype(np.int64)
    i = np.floor(I * rng.random(N) ** 3).astype(np.int64)
    y = (rng.integers(0, 2, N) if implicit else rng.integers(1, 6, N)).astype(np.float32)
    e = rng.integers(0, K, N).astype(np.int64)
    w = rng.random(N).astype(np.float32)
    p = {"Uinv": rng.normal(0, 0.1, (U, D)), "Iinv": rng.normal(0, 0.1, (I, D)), "Uenv": rng.normal(0, 0.3, (U, D)),
         "Ienv": rng.normal(0, 0.3, (I, D)), "E": rng.normal(0, 0.5, (K, D)), "W": rng.normal(0, 0.3, (K, D)),
         "b": rng.normal(0, 0.1, (K,))}
    return u, i, y, e, w, {k: v.astype(np.float32) for k, v in p.items()}


KW = dict(c_inv=0.8, c_ea=1.7, c_env=1.1, c_L2=0.6, c_L1=0.03, alpha=1.3, use_class_rw=True, use_rec_rw=True)


@pytest.mark.parametrize("implicit,K,D,U,I", [(False, 4, 64, 4000, 300), (True, 6, 40, 700, 90), (False, 2, 30, 50, 20)])
def test_lazy_is_bit_identical_to_dense(implicit, K, D, U, I):
    from invpref_kdd_2022_b200.engine import HotPath
    dev = torch.device("cuda:0")
    u, i, y, e, w, p = synth(U, I, 40000, K, D, implicit, 3)
    t = lambda a: torch.tensor(a, device=dev)
    # batches of very different sizes: many users are skipped for several steps, U=50 gives chunked segments
    cuts = [0, 9000, 9100, 9130, 20000, 20010, 33000, 33001, 40000]
    runs = {}
    for lazy in (False, True):
        hp = HotPath({k: t(v) for k, v in p.items()}, implicit, False, True, lr=1e-2, lazy=lazy)
        losses = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            losses.append(hp.train_step(t(u[a:b]), t(i[a:b]), t(y[a:b]), t(e[a:b]), t(w[a:b]), **KW).clone())
        if lazy:
            # before the flush some user rows are still behind ...
            stale = hp.params["Uinv"].clone()
            behind = bool((hp.last_step < hp.step).any())
            assert hp._dirty
            hp.flush()
            assert torch.equal(stale, hp.params["Uinv"]) != behind
            assert int(hp.last_step.min()) == hp.step == len(cuts) - 1
        runs[lazy] = (losses, {k: hp.params[k].clone() for k in on.PARAM_ORDER},
                      {k: hp.m[k].clone() for k in on.PARAM_ORDER}, {k: hp.v[k].clone() for k in on.PARAM_ORDER}, hp)
    for la, lb in zip(runs[False][0], runs[True][0]):
        assert torch.equal(la, lb)
    for which in (1, 2, 3):
        for k in on.PARAM_ORDER:
            assert torch.equal(runs[False][which][k], runs[True][which][k]), (which, k)
    # readers flush on their own: run two more steps, then cluster WITHOUT an explicit flush
    outs = []
    for lazy in (False, True):
        hp = runs[lazy][4]
        hp.train_step(t(u[:300]), t(i[:300]), t(y[:300]), t(e[:300]), t(w[:300]), **KW)
        hp.train_step(t(u[300:350]), t(i[300:350]), t(y[300:350]), t(e[300:350]), t(w[300:350]), **KW)
        new, hist, diff = hp.cluster(t(u[:5000]), t(i[:5000]), t(y[:5000]), None, None, t(e[:5000]))
        outs.append((new, hist, diff, hp.params["Uenv"].clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("case", ["coat_explicit", "implicit_k6"])
def test_trainer_lazy_equals_dense(case):
    """Same driver-style run with lazy_adam on and off: identical losses, parameters, env assignments."""
    from test_gpu_trainer import build
    g = Golden(case)
    res = []
    for lazy in (False, True):
        model, tm = build(g, epochs=3, cluster_interval=2, lazy_adam=lazy)
        out = tm.train(silent=True, auto=True)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        res.append((out, sd, tm.envs.clone(), tm.engine.lazy))
    assert res[0][3] is False and res[1][3] is True
    assert res[0][0][0] == res[1][0][0]                     # per-epoch loss dicts
    assert res[0][0][2] == res[1][0][2]                     # cluster diffs / env counts
    assert torch.equal(res[0][2], res[1][2])
    for k in res[0][1]:
        assert torch.equal(res[0][1][k], res[1][1][k]), k


def test_train_a_batch_flushes_and_state_dict_is_current():
    from test_gpu_trainer import build
    g = Golden("explicit_d64_k4")
    model, tm = build(g, lazy_adam=True)
    tm.stat_envs()
    B = 500
    for s in range(3):
        sl = slice(s * B, (s + 1) * B)
        tm.train_a_batch(tm.users_tensor[sl], tm.items_tensor[sl], tm.scores_tensor[sl], tm.envs[sl],
                         tm.sample_weights[sl], 1.0)
        assert not tm.engine._dirty
    model2, tm2 = build(g, lazy_adam=False)
    tm2.stat_envs()
    for s in range(3):
        sl = slice(s * B, (s + 1) * B)
        tm2.train_a_batch(tm2.users_tensor[sl], tm2.items_tensor[sl], tm2.scores_tensor[sl], tm2.envs[sl],
                          tm2.sample_weights[sl], 1.0)
    a, b = model.state_dict(), model2.state_dict()
    for k in a:
        assert torch.equal(a[k], b[k]), k
