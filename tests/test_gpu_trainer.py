"""GPU: the reference-facing API (models.py / train.py mirror) against the golden fixtures: same
constructor arguments, same call sequence as the drivers' main() (Coat_InvPref_explicit.py:68-109)."""
import numpy as np
import pytest
import torch

from _golden import CASES, Golden
from oracle import invpref_numpy as on

pytestmark = pytest.mark.gpu


def nerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


class NullEvaluator:
    def evaluate(self):
        return {"mse": 0.0}


def build(g: Golden, epochs=1, cluster_interval=1, **kw):
    from invpref_kdd_2022_b200.models import InvPrefExplicit, InvPrefImplicit
    from invpref_kdd_2022_b200.train import ExplicitTrainManager, ImplicitTrainManager
    dev = torch.device("cuda:0")
    torch.manual_seed(g.seed)
    np.random.seed(g.seed)
    M = InvPrefImplicit if g.implicit else InvPrefExplicit
    T = ImplicitTrainManager if g.implicit else ExplicitTrainManager
    model = M(g.U, g.I, g.K, g.D, g.roe, g.ree).to(dev)
    tm = T(model=model, evaluator=NullEvaluator(), device=dev, training_data=torch.LongTensor(g.data).to(dev),
           batch_size=g.B, epochs=epochs, cluster_interval=cluster_interval, evaluate_interval=1, lr=g.lr,
           invariant_coe=g.coef["c_inv"], env_aware_coe=g.coef["c_ea"], env_coe=g.coef["c_env"],
           L2_coe=g.coef["c_L2"], L1_coe=g.coef["c_L1"], alpha=g.alpha, use_class_re_weight=g.crw,
           use_recommend_re_weight=g.rrw, **kw)
    return model, tm


@pytest.mark.parametrize("case", CASES)
def test_init_matches_reference_rng_stream(case):
    """Same seed -> same initial parameters, initial envs and eps table as the reference (the model is
    initialised on the CPU generator in the reference's order, then moved)."""
    g = Golden(case)
    model, tm = build(g)
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    init = g.group("init")
    assert sorted(sd) == sorted(init)
    for k in init:
        assert np.array_equal(sd[k], init[k]), k
    assert np.array_equal(tm.envs.cpu().numpy(), g["envs0"])
    assert np.array_equal(tm.eps_random_tensor.cpu().numpy(), g["eps_table"])
    cnt = tm.stat_envs()
    assert np.array_equal(tm.sample_weights.cpu().numpy(), g["sample_weights0"])
    assert np.array_equal(tm.class_weights.cpu().numpy(), g["class_weights0"])
    assert sum(cnt.values()) == g.N


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("cache_plans", [True, False])
def test_epoch_cluster_stat_like_the_driver(case, cache_plans):
    g = Golden(case)
    model, tm = build(g, cache_plans=cache_plans)
    tm.stat_envs()
    mean_ld = tm.train_a_epoch()
    assert list(mean_ld) == list(on.LOSS_KEYS)
    ref = g["epoch_mean_loss"]
    for j, k in enumerate(on.LOSS_KEYS):
        assert abs(mean_ld[k] - ref[j]) <= 1e-4 * abs(ref[j]), k
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    for k, v in g.group("epoch1").items():
        assert nerr(sd[k], v) <= 5e-4, k
    assert tm.epoch_cnt == 1 and tm.engine.step == tm.batch_num
    # cluster(): the numpy stream is consumed exactly as in the reference (one randint per batch)
    np.random.seed(g.seed + 1)
    diff = tm.cluster()
    chk = np.random.randint(0, 1 << 30)
    np.random.seed(g.seed + 1)
    for lo, hi in on.mini_batch_bounds(g.N, g.B):
        np.random.randint(0, tm.eps_random_tensor.shape[0], hi - lo)
    assert chk == np.random.randint(0, 1 << 30)
    new = tm.envs.cpu().numpy()
    assert diff == int((new != g["envs0"]).sum())
    cnt = tm.stat_envs()
    assert [cnt[k] for k in range(g.K)] == np.bincount(new, minlength=g.K).tolist()
    assert np.array_equal(tm.sample_weights.cpu().numpy(), on.stat_envs(new, g.K, g.N)[2])


def test_device_tie_break_rng_leaves_the_numpy_stream_alone():
    """tie_break_rng="device" (an opt-in, NOT the reference's stream): cluster() draws the tie-break rows on the device;
    the assignments can differ from the numpy-stream run only where two distances agree to ~1e-10."""
    g = Golden("coat_explicit")
    _, tm_ref = build(g)
    tm_ref.stat_envs(); tm_ref.train_a_epoch()
    np.random.seed(5)
    tm_ref.cluster()
    _, tm = build(g, tie_break_rng="device")
    tm.stat_envs(); tm.train_a_epoch()
    np.random.seed(5)
    before = np.random.get_state()[1].copy()
    np.random.seed(5)
    diff = tm.cluster()
    assert np.array_equal(np.random.get_state()[1], before)            # no host draw at all
    a, b = tm.envs.cpu().numpy(), tm_ref.envs.cpu().numpy()
    assert diff == int((a != g["envs0"]).sum())
    assert (a != b).mean() <= 0.02                                     # fp32 near-ties only
    with pytest.raises(ValueError):
        build(g, tie_break_rng="philox")


def test_train_loop_returns_reference_triple():
    g = Golden("coat_explicit")
    model, tm = build(g, epochs=3, cluster_interval=2)
    (losses, loss_epochs), (tests, test_epochs), (diffs, env_cnts, cl_epochs) = tm.train(silent=True, auto=True)
    assert loss_epochs == [1, 2, 3] and len(losses) == 3
    assert test_epochs == [0, 1, 2, 3] and len(tests) == 4          # evaluate_interval = 1 (+ epoch 0)
    assert cl_epochs == [2] and len(diffs) == 1 and sum(env_cnts[0].values()) == g.N
    assert all(np.isfinite(list(d.values())).all() for d in losses)
    assert losses[2]["loss"] < losses[0]["loss"]


def test_train_a_batch_signature_and_state_dict_roundtrip():
    g = Golden("implicit_k2")
    model, tm = build(g)
    tm.stat_envs()
    B = g.B
    ld = tm.train_a_batch(tm.users_tensor[:B], tm.items_tensor[:B], tm.scores_tensor[:B], tm.envs[:B],
                          tm.sample_weights[:B], float(g["alpha0"]))
    ref = g["epoch_losses"][0]
    for j, k in enumerate(on.LOSS_KEYS):
        assert abs(ld[k] - ref[j]) <= 1e-5 * abs(ref[j]), k
    # parameters stay visible through the nn.Module after the buffer swap
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    p64 = on.params_from_state_dict(g.group("init"), np.float64)
    st = on.new_adam_state(p64, np.float64)
    d = g.data
    hyp = on.Hyper(alpha=float(g["alpha0"]), lr=g.lr, use_class_rw=g.crw, use_rec_rw=g.rrw, **g.coef)
    on.train_step(p64, st, d[:B, 0], d[:B, 1], d[:B, 2], g["envs0"][:B].astype(np.int64), g["sample_weights0"][:B],
                  hyp, on.Flags(g.implicit, g.roe, g.ree), np.float64)
    s1 = g.group("step1")
    for k, sk in on.STATE_KEYS.items():
        assert nerr(sd[sk], p64[k]) <= max(1e-5, 5 * nerr(s1[sk], p64[k])), k
    # a second model loaded from the state dict predicts the same
    from invpref_kdd_2022_b200.models import InvPrefImplicit
    m2 = InvPrefImplicit(g.U, g.I, g.K, g.D, g.roe, g.ree).to("cuda:0")
    m2.load_state_dict(model.state_dict())
    u = tm.users_tensor[:64]
    assert torch.equal(m2.predict(u), model.predict(u))
    assert m2.predict(u).shape == (64, g.I)


@pytest.mark.parametrize("implicit", [False, True])
def test_model_forward_backward_with_torch_autograd(implicit):
    """model(u, i, e, alpha) stays autograd-compatible: loss assembled in torch (as third-party code
    would), gradients through the fused backward, compared with a float64 torch evaluation."""
    from invpref_kdd_2022_b200.models import InvPrefExplicit, InvPrefImplicit
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    M = InvPrefImplicit if implicit else InvPrefExplicit
    model = M(300, 50, 4, 40, False, True)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(20.0)
    model = model.to(dev)
    gen = torch.Generator().manual_seed(0)
    u = torch.randint(0, 300, (3000,), generator=gen).to(dev)
    i = torch.randint(0, 50, (3000,), generator=gen).to(dev)
    e = torch.randint(0, 4, (3000,), generator=gen).to(dev)
    y = (torch.randint(0, 2, (3000,), generator=gen) if implicit else torch.randint(1, 6, (3000,), generator=gen))
    y = y.float().to(dev)
    alpha = 0.8

    def loss_of(s_inv, s_env, logp, mdl):
        rec = torch.nn.functional.binary_cross_entropy if implicit else torch.nn.functional.mse_loss
        return rec(s_inv, y.to(s_inv.dtype)) + 2.0 * rec(s_env, y.to(s_inv.dtype)) \
            + 1.5 * torch.nn.functional.nll_loss(logp, e) + 0.3 * mdl.get_L2_reg(u, i, e) + 0.01 * mdl.get_L1_reg(u, i, e)

    loss = loss_of(*model(u, i, e, alpha), model)
    loss.backward()
    # float64 torch twin
    sd = {k: v.double() for k, v in model.state_dict().items()}
    P = {k: sd[v].clone().requires_grad_(True) for k, v in on.STATE_KEYS.items()}
    a, c = P["Uinv"][u], P["Iinv"][i]
    pref = a * c
    z1, z2 = pref.sum(1), (P["Uenv"][u] * P["Ienv"][i] * P["E"][e]).sum(1)
    if implicit:
        s_inv = torch.sigmoid(z1); s_env = s_inv * torch.sigmoid(z2)
    else:
        s_inv = z1; s_env = z1 + z2
    rev = pref.detach() + (-alpha) * (pref - pref.detach())
    logp = torch.log_softmax(rev @ P["W"].T + P["b"], 1)

    class Twin:
        def reg(self, n):
            f = (lambda x: x.norm(2).pow(2)) if n == 2 else (lambda x: x.norm(1))
            r = (f(P["Uenv"][u]) + f(P["Uinv"][u])) / (3000 * 40 * 2) + (f(P["Ienv"][i]) + f(P["Iinv"][i])) / (3000 * 40 * 2)
            r = f(P["W"]) / 160 + f(P["b"]) / 4 + r
            return r + f(P["E"][e]) / (3000 * 40)

        def get_L2_reg(self, *a):
            return self.reg(2)

        def get_L1_reg(self, *a):
            return self.reg(1)

    ref = loss_of(s_inv, s_env, logp, Twin())
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    for k, path in on.STATE_KEYS.items():
        got = dict(model.named_parameters())[path].grad.cpu().numpy()
        assert nerr(got, P[k].grad.cpu().numpy()) <= 1e-5, k
    # cluster_predict == env-aware score
    with torch.no_grad():
        assert nerr(model.cluster_predict(u, i, e).cpu().numpy(), s_env.detach().cpu().numpy()) <= 1e-5


def test_bounded_plan_cache_streams_the_remaining_plans():
    """plan_cache_bytes: batches whose plan does not fit the cache get it rebuilt every epoch on a loader stream, one
    step ahead (two rotating buffers).  Same epochs, bit for bit, as with every plan cached -- and as with none."""
    g = Golden("explicit_d64_k4")          # 7000 interactions, B = 3000 -> 3 batches
    out = {}
    for name, kw in (("all", {}), ("one", {"plan_cache_bytes": None}), ("none", {"plan_cache_bytes": 0}),
                     ("nocache", {"cache_plans": False})):
        if name == "one":
            model, tm = build(g, epochs=3)
            kw = {"plan_cache_bytes": tm.engine.plan_bytes(g.B) + 1}      # room for exactly one plan
        model, tm = build(g, epochs=3, use_graph=False, **kw)
        tm.stat_envs()
        lds = [tm.train_a_epoch() for _ in range(3)]
        if name == "one":
            assert len(tm._plans) == 1 and tm._ring is not None
        if name == "none":
            assert len(tm._plans) == 0 and tm._ring is not None
        if name == "all":
            assert len(tm._plans) == 3 and tm._ring is None
        out[name] = (lds, {k: v.clone() for k, v in model.state_dict().items()})
    for name in ("one", "none", "nocache"):
        assert out[name][0] == out["all"][0], name
        for k in out["all"][1]:
            assert torch.equal(out[name][1][k], out["all"][1][k]), (name, k)


def test_model_move_and_optimizer_surface():
    """A no-op .to() keeps the engine (and flushes lazily updated rows first); a real re-allocation drops it and a
    trainer built before refuses to step; trainer.optimizer offers zero_grad / state_dict / load_state_dict."""
    g = Golden("coat_explicit")
    model, tm = build(g)
    tm.stat_envs()
    tm.train_a_epoch()
    eng = tm.engine
    model.to("cuda:0")                                   # no-op: same storages
    assert model._hot is eng
    tm.train_a_epoch()                                   # still steps
    tm.optimizer.zero_grad()
    sd = tm.optimizer.state_dict()
    assert len(sd["state"]) == 7 and int(sd["state"][0]["step"]) == eng.step == 2 * tm.batch_num
    assert torch.equal(sd["state"][1]["exp_avg"], eng.m["Iinv"])
    m0 = {k: v.clone() for k, v in eng.m.items()}
    tm.optimizer.load_state_dict(sd)
    assert all(torch.equal(eng.m[k], m0[k]) for k in m0)
    model.double()                                       # re-allocates every storage
    assert model._hot is None
    with pytest.raises(RuntimeError, match="re-allocated"):
        tm.train_a_epoch()


@pytest.mark.parametrize("K,D", [(8, 256), (6, 128)])
def test_trainer_default_lazy_falls_back_where_the_fused_pass_is_unavailable(K, D):
    """Shapes whose per-CTA dE/dW slices do not fit (K * D > 682) have no fused user pass and hence no lazy Adam: the
    trainer's default lazy_adam=True must fall back to plain dense Adam by itself and match the oracle."""
    from invpref_kdd_2022_b200.models import InvPrefExplicit
    from invpref_kdd_2022_b200.train import ExplicitTrainManager
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    U, I, N = 60, 40, 3000
    data = np.stack([rng.integers(0, U, N), rng.integers(0, I, N), rng.integers(1, 6, N)], axis=1).astype(np.int64)
    data[0, :2] = (U - 1, I - 1)
    torch.manual_seed(5)
    np.random.seed(5)
    model = InvPrefExplicit(U, I, K, D, True, False)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(10.0)
    model = model.to(dev)
    p64 = {k: prm.detach().cpu().numpy().astype(np.float64) for k, prm in model.named_hot_params().items()}
    tm = ExplicitTrainManager(model=model, evaluator=NullEvaluator(), device=dev, training_data=torch.LongTensor(data).to(dev),
                              batch_size=N, epochs=1, cluster_interval=1, evaluate_interval=1, lr=1e-2, invariant_coe=0.8,
                              env_aware_coe=1.7, env_coe=1.1, L2_coe=0.6, L1_coe=0.03, alpha=1.3,
                              use_class_re_weight=True, use_recommend_re_weight=True)
    assert tm.engine.lazy_requested and not tm.engine.lazy_supported and not tm.engine.lazy
    tm.stat_envs()
    envs, sw = tm.envs.cpu().numpy(), tm.sample_weights.cpu().numpy()
    ld = tm.train_a_epoch()
    hyp = on.Hyper(0.8, 1.7, 1.1, 0.6, 0.03, alpha=1.3, lr=1e-2, use_class_rw=True, use_rec_rw=True)
    st = on.new_adam_state(p64, np.float64)
    lo, _ = on.train_step(p64, st, data[:, 0], data[:, 1], data[:, 2].astype(np.float64), envs, sw.astype(np.float64), hyp,
                          on.Flags(False, True, False), np.float64)
    for k in on.LOSS_KEYS:
        assert abs(ld[k] - float(lo[k])) <= 1e-5 * abs(float(lo[k])), k
    for k, prm in model.named_hot_params().items():
        assert nerr(prm.detach().cpu().numpy(), p64[k]) <= 1e-4, k
