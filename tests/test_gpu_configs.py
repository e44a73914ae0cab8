"""GPU, BASELINE.json configs 2-4 at their NATIVE shapes and flag sets (the small fixtures are 5 000-9 000-interaction
miniatures and D = 40 takes the generic, 10-of-16-lanes kernels, not the tuned D = 64 ones):

  C2  Yahoo!R3 explicit   U 15 400  I 1 000   K 5 D 40  B 131 072  MSE           (Yahoo_InvPref_explicit.py:17-41)
  C3  MovieLens implicit  U 6 040   I 3 706   K 2 D 40  B 65 536   BCE, reg_env_embed, recommend re-weight
                                                                                 (MovieLens_InvPref.py:17-42)
  C4  MIND implicit       U 50 000  I 51 283  K 6 D 40  B 262 144  BCE, class re-weight (MIND_InvPref.py:17-42)

Each shape: one teacher-forced step against the reference's op sequence in torch-eager ON THE GPU (fp32, and fp64
for the tolerance rule: err(ours, fp64) <= max(1e-5, 2 err(ref fp32, fp64)) norm-wise per tensor), a multi-step
lazy == dense bit-identity run, the CUDA-graph epoch against the plain loop (bit-identical), and the EM
re-assignment against the K-forward argmin (fp32 near-ties counted).  Plus the REAL Yahoo!R3 explicit file: one
full epoch + cluster() of the trainer against what the live reference produced (tests/golden/make_golden_yahoo.py).
"""
import os

import numpy as np
import pytest
import torch

import bench
from oracle import invpref_numpy as on
from oracle import invpref_torch_cpu as ot

pytestmark = pytest.mark.gpu
SHAPES = ("c2", "c3", "c4")


def _nerr(a, b):
    den = float(b.abs().max())
    return float((a - b).abs().max()) / (den if den > 0 else 1.0)


def _setup(name, nb=1, dev=None):
    dev = dev or torch.device("cuda:0")
    w = bench.WORKLOADS[name]
    U, I, B, batches = bench.synth_batches(w, nb)
    dbs = [tuple(torch.from_numpy(a).to(dev) for a in b) for b in batches]
    kw = dict(alpha=0.8, use_class_rw=w["crw"], use_rec_rw=w["rrw"], **w["coef"])
    return w, dev, dbs, kw


@pytest.mark.parametrize("name", SHAPES)
def test_native_shape_step_matches_torch_eager_on_the_gpu(name):
    from invpref_kdd_2022_b200.engine import HotPath
    w, dev, dbs, kw = _setup(name)
    u, i, y, e = dbs[0]
    init = bench.make_tables(w, dev)
    hp = HotPath({k: v.clone() for k, v in init.items()}, w["implicit"], w["roe"], w["ree"], lr=w["lr"], lazy=True)
    _, sw = hp.stat_envs(e, hp.env_hist(e))
    grads = {k: torch.zeros_like(v) for k, v in init.items()}
    loss = hp.train_step(u, i, y, e, sw, grads_out=grads, **kw)
    hp.flush()
    loss = loss.cpu().numpy().astype(np.float64)
    flags = on.Flags(w["implicit"], w["roe"], w["ree"])
    hyp = on.Hyper(alpha=0.8, lr=w["lr"], use_class_rw=w["crw"], use_rec_rw=w["rrw"], **w["coef"])
    P = {k: v.clone().requires_grad_(True) for k, v in init.items()}
    ref = ot.CpuTrainer(P, flags, hyp).train_a_batch(u, i, y, e, sw, 0.8)
    for j, k in enumerate(on.LOSS_KEYS):
        assert abs(loss[j] - ref[k]) <= 1e-5 * abs(ref[k]), (name, k, loss[j], ref[k])
    P64 = {k: v.double().requires_grad_(True) for k, v in init.items()}
    tr64 = ot.CpuTrainer(P64, flags, hyp)
    tr64.opt.step = lambda *a, **k_: None
    tr64.train_a_batch(u, i, y.double(), e, sw.double(), 0.8)
    for k in on.PARAM_ORDER:
        g64 = P64[k].grad
        err_ours, err_ref = _nerr(grads[k].double(), g64), _nerr(P[k].grad.double(), g64)
        assert err_ours <= max(1e-5, 2 * err_ref), (name, k, err_ours, err_ref)
        ok = P[k].grad.abs() > 1e-7                   # Adam's first step is -lr g / (|g| + eps): compare where |g| >> eps
        if bool(ok.any()):
            d_ours, d_ref = (hp.params[k] - init[k])[ok], (P[k].detach() - init[k])[ok]
            assert float((d_ours - d_ref).abs().max()) <= 1e-3 * w["lr"], (name, k)


@pytest.mark.parametrize("name", SHAPES)
def test_native_shape_lazy_is_bitwise_dense(name):
    """Three batches, five steps (rows skipped for a step, then hit again): lazy Adam == dense Adam, bit for bit."""
    from invpref_kdd_2022_b200.engine import HotPath
    w, dev, dbs, kw = _setup(name, nb=3)
    init = bench.make_tables(w, dev, seed=5)
    res = {}
    for mode in ("dense", "lazy"):
        hp = HotPath({k: v.clone() for k, v in init.items()}, w["implicit"], w["roe"], w["ree"], lr=w["lr"],
                     lazy=(mode == "lazy"))
        losses = []
        for s in (0, 1, 2, 0, 1):
            u, i, y, e = dbs[s]
            _, sw = hp.stat_envs(e, hp.env_hist(e))
            losses.append(hp.train_step(u, i, y, e, sw, **kw).clone())
        hp.flush()
        res[mode] = (torch.stack(losses), {k: hp.params[k].clone() for k in on.PARAM_ORDER},
                     {k: hp.m[k].clone() for k in on.PARAM_ORDER}, {k: hp.v[k].clone() for k in on.PARAM_ORDER})
    assert torch.isfinite(res["dense"][0]).all()
    assert torch.equal(res["dense"][0], res["lazy"][0])
    for grp in (1, 2, 3):
        for k in on.PARAM_ORDER:
            assert torch.equal(res["dense"][grp][k], res["lazy"][grp][k]), (name, grp, k)


class _NullEvaluator0:
    def evaluate(self):
        return {"mse": 0.0}


@pytest.mark.parametrize("name", SHAPES)
@pytest.mark.parametrize("lazy", [True, False])
def test_native_shape_graph_epochs_are_bitwise_the_plain_loop(name, lazy):
    """The trainer replays every epoch after the first as ONE CUDA graph (invpref_graph_*, step-dependent scalars in
    device records): four epochs with a cluster() + stat_envs() in between (they re-bind envs / sample_weights) and
    an odd number of batches (the double buffers end an epoch swapped) must equal the launch-by-launch loop bit for
    bit -- losses, every table, every Adam moment."""
    from invpref_kdd_2022_b200.models import InvPrefExplicit, InvPrefImplicit
    from invpref_kdd_2022_b200.train import ExplicitTrainManager, ImplicitTrainManager
    w = bench.WORKLOADS[name]
    dev = torch.device("cuda:0")
    U, I, B, batches = bench.synth_batches(w, 3)
    data = np.concatenate([np.stack([u, i, y.astype(np.int64)], axis=1) for (u, i, y, e) in batches])
    data = data[:2 * B + B // 3]                              # 3 batches, the last one short
    data[0, 0], data[0, 1] = U - 1, I - 1
    out = {}
    for use_graph in (False, True):
        torch.manual_seed(11)
        np.random.seed(11)
        M, T = (InvPrefImplicit, ImplicitTrainManager) if w["implicit"] else (InvPrefExplicit, ExplicitTrainManager)
        model = M(U, I, w["K"], w["D"], w["roe"], w["ree"]).to(dev)
        c = w["coef"]
        tm = T(model=model, evaluator=_NullEvaluator0(), device=dev, training_data=torch.LongTensor(data).to(dev),
               batch_size=B, epochs=4, cluster_interval=2, evaluate_interval=100, lr=w["lr"],
               invariant_coe=c["c_inv"], env_aware_coe=c["c_ea"], env_coe=c["c_env"], L2_coe=c["c_L2"],
               L1_coe=c["c_L1"], alpha=None, use_class_re_weight=w["crw"], use_recommend_re_weight=w["rrw"],
               lazy_adam=lazy, use_graph=use_graph)
        assert tm.batch_num == 3
        (losses, _), _, (diffs, cnts, _) = tm.train(silent=True, auto=True)
        if use_graph:
            assert tm._graph is not None and len(tm._graph.handles) == 2      # both start parities were captured
            assert tm._graph.launches(0) >= 3 * 5
        eng = tm.engine
        eng.flush()
        out[use_graph] = (losses, diffs, cnts, {k: v.clone() for k, v in eng.params.items()},
                          {k: v.clone() for k, v in eng.m.items()}, {k: v.clone() for k, v in eng.v.items()},
                          eng.step, tm.alpha)
    a, b = out[False], out[True]
    assert a[0] == b[0] and a[1] == b[1] and a[2] == b[2] and a[6] == b[6] == 12 and a[7] == b[7]
    for grp in (3, 4, 5):
        for k in on.PARAM_ORDER:
            assert torch.equal(a[grp][k], b[grp][k]), (name, lazy, grp, k)
    assert all(np.isfinite(list(d.values())).all() for d in a[0])


@pytest.mark.parametrize("name", SHAPES)
def test_native_shape_cluster_matches_k_forward_argmin(name):
    from invpref_kdd_2022_b200.engine import HotPath
    w, dev, dbs, kw = _setup(name)
    u, i, y, e = dbs[0]
    K, B = w["K"], u.numel()
    init = bench.make_tables(w, dev, seed=9)
    # env-aware tables scaled up so the K distances are separated (at init scale most samples are fp32 near-ties)
    big = {k: (v * (30.0 if k in ("Uenv", "Ienv") else (50.0 if k == "E" else (10.0 if k in ("Uinv", "Iinv") else 1.0))))
           for k, v in init.items()}
    hp = HotPath(big, w["implicit"], w["roe"], w["ree"], lr=w["lr"])
    import itertools
    base = torch.Tensor([1e-10 * (1e-1 ** k) for k in range(K)])
    eps = torch.Tensor(list(itertools.permutations(base))).to(dev)                    # train.py:763-769
    pidx = torch.from_numpy(np.random.default_rng(3).integers(0, eps.shape[0], B)).to(dev)
    for perm in (None, pidx):
        new, hist, diff = hp.cluster(u, i, y, perm, eps if perm is not None else None, e)
        assert int(hist.sum()) == B and int(diff) == int((new != e).sum())
        assert torch.equal(hist, torch.bincount(new, minlength=K))
        with torch.no_grad():
            cols = []
            for k in range(K):
                ek = torch.full((B,), k, dtype=torch.int64, device=dev)
                _, s_env, _ = ot.forward(big, u, i, ek, 0.0, w["implicit"])
                cols.append(torch.nn.functional.binary_cross_entropy(s_env, y, reduction="none") if w["implicit"]
                            else (s_env - y) ** 2)
            dist = torch.stack(cols, dim=1)
            if perm is not None:
                dist = dist + eps[perm]
        mism = (new != torch.argmin(dist, dim=1)).cpu().numpy()
        ties = on.near_tie_mask(dist.cpu().numpy())
        assert not (mism & ~ties).any(), (name, int((mism & ~ties).sum()))
        assert mism.sum() <= 2e-3 * B, (name, int(mism.sum()))


# ------------------------------------------------------------------------------------------------------------
# the REAL Yahoo!R3 explicit training file at the driver's own configuration, against the live reference's output
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "yahoo_explicit_full.npz")


class _NullEvaluator:
    def evaluate(self):
        return {"mse": 0.0}


@pytest.mark.skipif(not os.path.isfile(GOLD), reason="tests/golden/yahoo_explicit_full.npz not generated")
def test_yahoo_explicit_real_file_epoch_and_cluster_match_the_live_reference():
    from invpref_kdd_2022_b200.drivers import Yahoo_InvPref_explicit as drv
    from invpref_kdd_2022_b200.models import InvPrefExplicit
    from invpref_kdd_2022_b200.train import ExplicitTrainManager
    z = np.load(GOLD)
    meta = dict(zip(z["meta_keys"].tolist(), z["meta_vals"].tolist()))
    seed, stride = int(meta["seed"]), int(z["user_stride"])
    data = np.stack([z["users"], z["items"], z["scores"]], axis=1).astype(np.int64)
    U, I, N = int(data[:, 0].max()) + 1, int(data[:, 1].max()) + 1, len(data)
    assert (U, I, N) == drv.SHAPE
    mc, tc = drv.MODEL_CONFIG, drv.TRAIN_CONFIG
    assert (mc["env_num"], mc["factor_num"], tc["batch_size"]) == (int(meta["K"]), int(meta["D"]), int(meta["B"]))
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    np.random.seed(seed)
    model = InvPrefExplicit(U, I, mc["env_num"], mc["factor_num"], mc["reg_only_embed"], mc["reg_env_embed"])
    chk = sum(int(v.detach().contiguous().view(torch.int32).to(torch.int64).sum()) for v in model.state_dict().values())
    assert chk == int(z["init_checksum"])                  # same RNG stream as the reference's constructor
    model = model.to(dev)
    tm = ExplicitTrainManager(
        model=model, evaluator=_NullEvaluator(), device=dev, training_data=torch.LongTensor(data).to(dev),
        batch_size=tc["batch_size"], epochs=1, cluster_interval=1, evaluate_interval=1, lr=tc["lr"],
        invariant_coe=tc["invariant_coe"], env_aware_coe=tc["env_aware_coe"], env_coe=tc["env_coe"],
        L2_coe=tc["L2_coe"], L1_coe=tc["L1_coe"], alpha=tc["alpha"], use_class_re_weight=tc["use_class_re_weight"],
        use_recommend_re_weight=tc["use_recommend_re_weight"])
    assert np.array_equal(tm.envs.cpu().numpy(), z["envs0"])
    tm.stat_envs()
    assert np.array_equal(tm.class_weights.cpu().numpy(), z["class_weights0"])

    # step 0, teacher forced, with the gradients exported: the tolerance rule against the reference's fp64 twin
    B = tc["batch_size"]
    hp = model.hot_path()
    snap = {k: v.clone() for k, v in hp.params.items()}
    grads = {k: torch.zeros_like(v) for k, v in hp.params.items()}
    hp2 = type(hp)(snap, False, mc["reg_only_embed"], mc["reg_env_embed"], lr=tc["lr"])
    loss0 = hp2.train_step(tm.users_tensor[:B], tm.items_tensor[:B], tm.scores_tensor[:B], tm.envs[:B],
                           tm.sample_weights[:B], c_inv=tc["invariant_coe"], c_ea=tc["env_aware_coe"],
                           c_env=tc["env_coe"], c_L2=tc["L2_coe"], c_L1=tc["L1_coe"], alpha=float(z["alpha0"]),
                           use_class_rw=False, use_rec_rw=False, grads_out=grads).cpu().numpy()
    # losses by the same rule as the gradients: at this size the reference's own fp32 L2 / L1 values are ~2e-4 off its
    # fp64 twin (norm() over 5.2 M gathered elements accumulated in fp32 on the CPU)
    ref_losses, l64 = z["epoch_losses"], z["loss0_f64"]
    for j, k in enumerate(on.LOSS_KEYS):
        err_ours, err_ref = abs(loss0[j] - l64[j]) / abs(l64[j]), abs(ref_losses[0][j] - l64[j]) / abs(l64[j])
        assert err_ours <= max(1e-5, 2 * err_ref), (k, loss0[j], ref_losses[0][j], l64[j])
    ref_err = dict(zip(z["grad0_ref_err_keys"].tolist(), z["grad0_ref_err"].tolist()))
    for k, sk in on.STATE_KEYS.items():
        g = grads[k].cpu().numpy().astype(np.float64)
        if sk.startswith("embed_user"):
            g = g[::stride]
        g64 = z["grad0_f64/" + sk]
        err = float(np.abs(g - g64).max() / np.abs(g64).max())
        assert err <= max(1e-5, 2 * ref_err[sk]), (k, err, ref_err[sk])

    # the full epoch through the trainer (3 steps, alpha schedule, cached plans, CUDA-graph replay allowed)
    mean_ld = tm.train_a_epoch()
    m64 = z["epoch_mean_loss_f64"]
    for j, k in enumerate(on.LOSS_KEYS):
        err_ours = abs(mean_ld[k] - m64[j]) / abs(m64[j])
        err_ref = abs(z["epoch_mean_loss"][j] - m64[j]) / abs(m64[j])
        assert err_ours <= max(1e-4, 2 * err_ref), (k, mean_ld[k], z["epoch_mean_loss"][j], m64[j])
    sd = {k: v.cpu().numpy() for k, v in model.state_dict().items()}
    for k in sd:
        ours = sd[k][::stride] if k.startswith("embed_user") else sd[k]
        want, w64 = z["epoch1/" + k], z["epoch1_f64/" + k]
        den = np.abs(w64).max()
        err_ours = float(np.abs(ours.astype(np.float64) - w64).max() / den)
        err_ref = float(np.abs(want.astype(np.float64) - w64).max() / den)
        assert err_ours <= max(5e-4, 2 * err_ref), (k, err_ours, err_ref)

    # cluster(): same host-drawn tie-break stream; assignments equal except fp32 near-ties (bounded by the count
    # the reference's own distances have)
    np.random.seed(seed + 1)
    diff = tm.cluster()
    new = tm.envs.cpu().numpy()
    mism = int((new != z["cluster_envs"]).sum())
    assert mism <= max(int(z["cluster_near_ties"]), int(0.05 * N)), (mism, int(z["cluster_near_ties"]))
    assert abs(diff - int(z["cluster_diff"])) <= mism
    cnt = tm.stat_envs()
    assert sum(cnt.values()) == N
    # at the trained scale 93 % of the samples are fp32 near-ties; with the env-aware tables scaled
    # by 20 the K distances are separated: assignments must agree except where the two smallest distances are
    # within 1e-3 (what the 5e-4 parameter drift of three untethered steps can flip)
    with torch.no_grad():
        for t in (model.embed_user_env_aware, model.embed_item_env_aware, model.embed_env):
            t.weight.mul_(20.0)
    np.random.seed(seed + 2)
    tm.cluster()
    mism = int((tm.envs.cpu().numpy() != z["sep_envs"]).sum())
    assert mism <= int(z["sep_loose_ties"]), (mism, int(z["sep_loose_ties"]), int(z["sep_near_ties"]))
