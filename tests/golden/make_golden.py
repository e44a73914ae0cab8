"""Generate the golden fixtures by running the LIVE reference (build container only).

    python tests/golden/make_golden.py

Each ``<case>.npz`` holds the inputs (ids, scores, initial envs, initial parameters, the
host-drawn tie-break indices) and what the reference's own ``train.py`` / ``models.py``
produced from them on CPU in fp32 (and, for the step-0 gradients, in fp64):

  step-0 forward (s_inv, s_env, logp), step-0 gradients, parameters + Adam state after
  step 0, the loss dict of every step of one epoch, parameters + Adam state after that
  epoch, then ``cluster()`` (new envs, diff_num) and ``stat_envs()`` after it, and a second
  ``cluster()`` after scaling the env-aware tables by 20 so that the K distances are well
  separated (at N(0, 0.01) init scale nearly every sample is an fp32 tie).

The reference ships no tests/golden vectors of its own, so these files are
what pins the oracle and the CUDA path.  Re-running this script must reproduce them
bit-for-bit on the same torch build (torch 2.11.0 CPU, 8 threads).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from oracle import invpref_numpy as on  # noqa: E402

CASES = {
    # the reference's own CPU-runnable case: Coat explicit, real file, driver hyper-parameters
    # (Coat_InvPref_explicit.py:17-42)
    "coat_explicit": dict(
        implicit=False, data="Coat_explicit_all_data", K=4, D=30, B=1024, roe=True, ree=False,
        lr=0.01, c_inv=2.050646960185343, c_ea=8.632289952059462, c_env=5.100067503854663,
        c_L2=7.731619515414727, c_L1=0.0015415961377493945, alpha=1.7379692382330174,
        crw=True, rrw=True, seed=17373331),
    # Yahoo explicit shape, scaled down, alpha schedule (Yahoo_InvPref_explicit.py:17-41)
    "explicit_sched_k5": dict(
        implicit=False, data=(160, 50, 9000), K=5, D=40, B=4096, roe=True, ree=False,
        lr=1e-3, c_inv=0.007375309563638757, c_ea=7.207790368836971, c_env=7.30272189219841,
        c_L2=5.105587170019545, c_L1=0.004098813161410509, alpha=None,
        crw=False, rrw=False, seed=17373511),
    # MovieLens flags (MovieLens_InvPref.py:17-42): BCE, K=2, reg_env_embed, recommend re-weight
    "implicit_k2": dict(
        implicit=True, data=(120, 80, 5000), K=2, D=40, B=2048, roe=True, ree=True,
        lr=1e-2, c_inv=8.909348155983732, c_ea=1.233057369609993, c_env=8.064376793624795,
        c_L2=3.4987474005653665, c_L1=0.9355983539586914, alpha=None,
        crw=False, rrw=True, seed=17373423),
    # MIND flags (MIND_InvPref.py:17-42): BCE, K=6, class re-weight; classifier regularised too
    "implicit_k6": dict(
        implicit=True, data=(100, 140, 6000), K=6, D=40, B=2500, roe=False, ree=False,
        lr=1e-3, c_inv=0.41343891722673093, c_ea=9.833594297680568, c_env=7.521558049068597,
        c_L2=4.324061954456766, c_L1=0.33322012936680223, alpha=1.5359474241627789,
        crw=True, rrw=False, seed=17373331),
    # synthetic-scale shape: D=64, K=4, heavy duplicates (hot items), short last batch
    "explicit_d64_k4": dict(
        implicit=False, data=(90, 30, 7000), K=4, D=64, B=3000, roe=True, ree=True,
        lr=5e-3, c_inv=1.0, c_ea=2.0, c_env=1.5, c_L2=0.8, c_L1=0.05, alpha=0.9,
        crw=True, rrw=True, seed=20220814),
}


def _load_data(spec, implicit, seed):
    if isinstance(spec, str):
        import pandas as pd
        df = pd.read_csv(os.path.join(ref_shim.REF_ROOT, "dataset", spec, "train.csv"))
        return df.values.astype(np.int64)
    U, I, N = spec
    rng = np.random.default_rng(seed)
    u = np.floor(U * rng.random(N) ** 1.5).astype(np.int64)          # SURVEY.md §8d generators
    i = np.floor(I * rng.random(N) ** 3).astype(np.int64)
    u[0], i[0] = U - 1, I - 1                                        # pin table sizes (max id + 1)
    y = rng.integers(0, 2, N) if implicit else rng.integers(1, 6, N)
    return np.stack([u, i, y], axis=1).astype(np.int64)


def _sd(model):
    return {k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def _adam(tm, model, which):
    out = {}
    for name, prm in model.named_parameters():
        out[name] = tm.optimizer.state[prm][which].detach().numpy().copy()
    return out


def make(name, cfg, ref):
    data = _load_data(cfg["data"], cfg["implicit"], cfg["seed"])
    U, I = int(data[:, 0].max()) + 1, int(data[:, 1].max()) + 1
    N, K, D, B = len(data), cfg["K"], cfg["D"], cfg["B"]
    M = ref.models.InvPrefImplicit if cfg["implicit"] else ref.models.InvPrefExplicit
    T = ref.train.ImplicitTrainManager if cfg["implicit"] else ref.train.ExplicitTrainManager

    def build(double=False):
        torch.manual_seed(cfg["seed"])
        np.random.seed(cfg["seed"])
        model = M(U, I, K, D, cfg["roe"], cfg["ree"])
        if double:
            model = model.double()
        tm = T(model=model, evaluator=ref_shim.NullEvaluator(), device=torch.device("cpu"),
               training_data=torch.LongTensor(data), batch_size=B, epochs=1, cluster_interval=1,
               evaluate_interval=1, lr=cfg["lr"], invariant_coe=cfg["c_inv"], env_aware_coe=cfg["c_ea"],
               env_coe=cfg["c_env"], L2_coe=cfg["c_L2"], L1_coe=cfg["c_L1"], alpha=cfg["alpha"],
               use_class_re_weight=cfg["crw"], use_recommend_re_weight=cfg["rrw"])
        tm.stat_envs()
        return model, tm

    out = {"data": data.astype(np.int32), "meta_keys": np.array(sorted(k for k in cfg if k != "data")),
           "meta_vals": np.array([str(cfg[k]) for k in sorted(k for k in cfg if k != "data")])}
    model, tm = build()
    for k, v in _sd(model).items():
        out["init/" + k] = v
    out["envs0"] = tm.envs.numpy().astype(np.int8)
    out["sample_weights0"] = tm.sample_weights.numpy().copy()
    out["class_weights0"] = tm.class_weights.numpy().copy()
    out["eps_table"] = tm.eps_random_tensor.numpy().copy()

    # ---- step 0, teacher forced: forward, grads (fp32 and fp64), post-step state ----
    alpha0 = cfg["alpha"] if cfg["alpha"] is not None else on.alpha_schedule(0, 0, tm.batch_num)
    out["alpha0"] = np.float64(alpha0)
    sl = slice(0, min(B, N))
    u, i = tm.users_tensor[sl], tm.items_tensor[sl]
    y, e, w = tm.scores_tensor[sl], tm.envs[sl], tm.sample_weights[sl]
    with torch.no_grad():
        s_inv, s_env, logp = model(u, i, e, alpha0)
    out["fwd0/s_inv"], out["fwd0/s_env"], out["fwd0/logp"] = s_inv.numpy(), s_env.numpy(), logp.numpy()

    md, tmd = build(double=True)
    tmd.train_a_batch(u, i, y.double(), e, w.double(), alpha0)
    for k, prm in md.named_parameters():
        out["grad0_f64/" + k] = prm.grad.numpy().copy()

    # ---- one full epoch in fp32 (train.py:881-910) ----
    losses = []
    orig = tm.train_a_batch

    def spy(**kw):
        ld = orig(**kw)
        losses.append([ld[k] for k in on.LOSS_KEYS])
        if len(losses) == 1:
            for k, prm in model.named_parameters():
                out["grad0/" + k] = prm.grad.numpy().copy()
            for k, v in _sd(model).items():
                out["step1/" + k] = v
            for k, v in _adam(tm, model, "exp_avg").items():
                out["step1_m/" + k] = v
            for k, v in _adam(tm, model, "exp_avg_sq").items():
                out["step1_v/" + k] = v
        return ld

    tm.train_a_batch = spy
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mean_ld = tm.train_a_epoch()
    out["epoch_losses"] = np.asarray(losses, dtype=np.float64)            # [steps, 6] python floats
    out["epoch_mean_loss"] = np.asarray([mean_ld[k] for k in on.LOSS_KEYS], dtype=np.float64)
    for k, v in _sd(model).items():
        out["epoch1/" + k] = v
    for k, v in _adam(tm, model, "exp_avg").items():
        out["epoch1_m/" + k] = v
    for k, v in _adam(tm, model, "exp_avg_sq").items():
        out["epoch1_v/" + k] = v

    # ---- cluster() + stat_envs() (train.py:912-957); the tie-break draws come from the numpy
    # global stream, one randint per cluster batch (train.py:870-871) ----
    np.random.seed(cfg["seed"] + 1)
    idx = [np.random.randint(0, tm.eps_random_tensor.shape[0], hi - lo) for lo, hi in on.mini_batch_bounds(N, B)]
    out["cluster_perm_idx"] = np.concatenate(idx).astype(np.int16)
    np.random.seed(cfg["seed"] + 1)
    diff = tm.cluster()
    out["cluster_envs"] = tm.envs.numpy().astype(np.int8)
    out["cluster_diff"] = np.int64(diff)
    cnt = tm.stat_envs()
    out["stat_counts"] = np.asarray([cnt[k] for k in range(K)], dtype=np.int64)
    out["stat_class_weights"] = tm.class_weights.numpy().copy()
    # distances the reference argmin'd, for tie accounting (fp32)
    p = on.params_from_state_dict({k[len("epoch1/"):]: v for k, v in out.items() if k.startswith("epoch1/")})
    dist = on.cluster_distances(p, data[:, 0], data[:, 1], data[:, 2], on.Flags(cfg["implicit"], cfg["roe"], cfg["ree"]))
    out["cluster_near_ties"] = np.int64(on.near_tie_mask(dist).sum())
    # ---- second cluster() with well-separated distances ----
    with torch.no_grad():
        for t in (model.embed_user_env_aware, model.embed_item_env_aware, model.embed_env):
            t.weight.mul_(20.0)
    for k, v in _sd(model).items():
        if "env" in k and "classifier" not in k:
            out["sep/" + k] = v
    np.random.seed(cfg["seed"] + 2)
    idx = [np.random.randint(0, tm.eps_random_tensor.shape[0], hi - lo) for lo, hi in on.mini_batch_bounds(N, B)]
    out["sep_perm_idx"] = np.concatenate(idx).astype(np.int16)
    np.random.seed(cfg["seed"] + 2)
    out["sep_diff"] = np.int64(tm.cluster())
    out["sep_envs"] = tm.envs.numpy().astype(np.int8)
    p = on.params_from_state_dict(_sd(model))
    dist = on.cluster_distances(p, data[:, 0], data[:, 1], data[:, 2], on.Flags(cfg["implicit"], cfg["roe"], cfg["ree"]))
    out["sep_near_ties"] = np.int64(on.near_tie_mask(dist).sum())
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: U={U} I={I} N={N} steps={len(losses)} diff={diff} near_ties={int(out['cluster_near_ties'])} sep_diff={int(out['sep_diff'])} sep_ties={int(out['sep_near_ties'])} "
          f"-> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    ref = ref_shim.load()
    only = sys.argv[1:]
    for name, cfg in CASES.items():
        if not only or name in only:
            make(name, cfg, ref)
