"""Closed-form numpy restatement of the InvPref hot path.  TEST INFRASTRUCTURE ONLY.

Every function restates one piece of the reference (AIflowerQ/InvPref_KDD_2022, paths
relative to the reference checkout) without calling torch, so it is an independent
check of both the reference's autograd and the CUDA kernels.  ``dt`` selects the
arithmetic type (np.float32 to mimic the reference, np.float64 for the "truth" twin
used by the tolerance rule in tests).

Parity pinning: checked against the live reference by ``tests/golden/make_golden.py``
(build container only) and against the committed fixtures ``tests/golden/*.npz`` by
``tests/test_oracle_golden.py`` (runs anywhere).
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass, field

import numpy as np

# our short names -> reference ``state_dict`` keys (models.py:424-432, 200)
STATE_KEYS = {
    "Uinv": "embed_user_invariant.weight",
    "Iinv": "embed_item_invariant.weight",
    "Uenv": "embed_user_env_aware.weight",
    "Ienv": "embed_item_env_aware.weight",
    "E": "embed_env.weight",
    "W": "env_classifier.linear_map.weight",
    "b": "env_classifier.linear_map.bias",
}
PARAM_ORDER = ("Uinv", "Iinv", "Uenv", "Ienv", "E", "W", "b")  # == model.parameters() order
LOSS_KEYS = ("invariant_loss", "env_aware_loss", "envs_loss", "L2_reg", "L1_reg", "loss")  # train.py:836-843


@dataclass
class Hyper:
    """Loss coefficients and optimiser settings (train.py:694-703, 718)."""

    c_inv: float
    c_ea: float
    c_env: float
    c_L2: float
    c_L1: float
    alpha: float = 1.0
    lr: float = 1e-3
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    use_class_rw: bool = False
    use_rec_rw: bool = True


@dataclass
class Flags:
    """Model variant flags (models.py:415-418 / 273-276)."""

    implicit: bool = False
    reg_only_embed: bool = False
    reg_env_embed: bool = True


def params_from_state_dict(sd, dt=np.float32) -> dict:
    return {k: np.ascontiguousarray(np.asarray(sd[v]), dtype=dt) for k, v in STATE_KEYS.items()}


def params_to_state_dict(p) -> dict:
    return {v: np.asarray(p[k]) for k, v in STATE_KEYS.items()}


def _sigmoid(x):
    one = x.dtype.type(1)
    return one / (one + np.exp(-x))


def forward(p, u, i, e, flags: Flags, dt=np.float32):
    """models.py:448-467 (explicit) / 307-326 (implicit).  Returns extra intermediates."""
    a = p["Uinv"].astype(dt)[u]
    c = p["Iinv"].astype(dt)[i]
    ue = p["Uenv"].astype(dt)[u]
    ie = p["Ienv"].astype(dt)[i]
    ee = p["E"].astype(dt)[e]
    pref = a * c                                    # models.py:457 / 316
    q = ue * ie * ee                                # models.py:458 / 317
    z1 = pref.sum(axis=1, dtype=dt)
    z2 = q.sum(axis=1, dtype=dt)
    if flags.implicit:                              # models.py:319-321
        s_inv = _sigmoid(z1)
        s2 = _sigmoid(z2)
        s_env = s_inv * s2
    else:                                           # models.py:460-462
        s_inv = z1
        s2 = z2
        s_env = z1 + z2
    logits = pref @ p["W"].astype(dt).T + p["b"].astype(dt)   # models.py:207
    m = logits.max(axis=1, keepdims=True)
    logp = logits - m - np.log(np.exp(logits - m).sum(axis=1, keepdims=True))  # models.py:208
    inter = dict(a=a, c=c, ue=ue, ie=ie, ee=ee, pref=pref, z1=z1, z2=z2, s2=s2, logits=logits)
    return s_inv.astype(dt), s_env.astype(dt), logp.astype(dt), inter


def _rec_loss(x, y, implicit, dt):
    """nn.MSELoss / nn.BCELoss per element (train.py:719 / 42); BCE log clamp at -100."""
    if implicit:
        lo = dt(-100.0)
        with np.errstate(divide="ignore"):
            lx = np.maximum(np.log(x), lo)
            l1x = np.maximum(np.log(dt(1) - x), lo)
        return -(y * lx + (dt(1) - y) * l1x)
    return (x - y) ** 2


def _rec_loss_grad(x, y, implicit, dt):
    """d loss / d x per element; BCE backward clamps x(1-x) at 1e-12 (ATen binary_cross_entropy_backward)."""
    if implicit:
        return (x - y) / np.maximum(x * (dt(1) - x), dt(1e-12))
    return dt(2) * (x - y)


def loss_and_grads(p, u, i, y, e, w, hp: Hyper, flags: Flags, dt=np.float32):
    """Loss assembly of train.py:771-830 and the closed-form backward.

    Returns (loss dict with the six keys of train.py:836-843, dense grads dict).
    """
    u = np.asarray(u, dtype=np.int64)
    i = np.asarray(i, dtype=np.int64)
    e = np.asarray(e, dtype=np.int64)
    y = np.asarray(y, dtype=dt)
    w = np.asarray(w, dtype=dt)
    B = len(u)
    K, D = p["W"].shape
    s_inv, s_env, logp, t = forward(p, u, i, e, flags, dt)
    a, c, ue, ie, ee, pref, s2 = t["a"], t["c"], t["ue"], t["ie"], t["ee"], t["pref"], t["s2"]
    W = p["W"].astype(dt)
    b = p["b"].astype(dt)
    ones = np.ones(B, dtype=dt)
    w_r = w if hp.use_rec_rw else ones               # train.py:817-819
    w_c = w if hp.use_class_rw else ones             # train.py:814-815

    l_inv = _rec_loss(s_inv, y, flags.implicit, dt)
    l_ea = _rec_loss(s_env, y, flags.implicit, dt)
    nll = -logp[np.arange(B), e]                     # train.py:812
    inv_loss = (l_inv * w_r).mean(dtype=dt)
    ea_loss = (l_ea * w_r).mean(dtype=dt)
    envs_loss = (nll * w_c).mean(dtype=dt)

    # models.py:469-532: norms over the GATHERED rows, duplicates count
    bd2 = dt(B) * dt(D) * dt(2)
    bd = dt(B) * dt(D)
    L2 = (np.sum(ue * ue, dtype=dt) + np.sum(a * a, dtype=dt)) / bd2 \
        + (np.sum(ie * ie, dtype=dt) + np.sum(c * c, dtype=dt)) / bd2
    L1 = (np.abs(ue).sum(dtype=dt) + np.abs(a).sum(dtype=dt)) / bd2 \
        + (np.abs(ie).sum(dtype=dt) + np.abs(c).sum(dtype=dt)) / bd2
    if not flags.reg_only_embed:                     # models.py:210-217
        L2 = L2 + np.sum(W * W, dtype=dt) / dt(D * K) + np.sum(b * b, dtype=dt) / dt(K)
        L1 = L1 + np.abs(W).sum(dtype=dt) / dt(D * K) + np.abs(b).sum(dtype=dt) / dt(K)
    if flags.reg_env_embed:                          # models.py:499-504
        L2 = L2 + np.sum(ee * ee, dtype=dt) / bd
        L1 = L1 + np.abs(ee).sum(dtype=dt) / bd
    loss = inv_loss * dt(hp.c_inv) + ea_loss * dt(hp.c_ea) + envs_loss * dt(hp.c_env) \
        + L2 * dt(hp.c_L2) + L1 * dt(hp.c_L1)       # train.py:829-830
    losses = dict(zip(LOSS_KEYS, (inv_loss, ea_loss, envs_loss, L2, L1, loss)))

    # ---- backward ----
    invB = dt(1) / dt(B)
    r_inv = _rec_loss_grad(s_inv, y, flags.implicit, dt)
    r_ea = _rec_loss_grad(s_env, y, flags.implicit, dt)
    if flags.implicit:
        g_s1 = w_r * invB * (dt(hp.c_inv) * r_inv + dt(hp.c_ea) * r_ea * s2)
        g_s2 = w_r * invB * dt(hp.c_ea) * r_ea * s_inv
        g_z1 = g_s1 * s_inv * (dt(1) - s_inv)
        g_z2 = g_s2 * s2 * (dt(1) - s2)
    else:
        g_z1 = w_r * invB * (dt(hp.c_inv) * r_inv + dt(hp.c_ea) * r_ea)
        g_z2 = w_r * invB * dt(hp.c_ea) * r_ea
    soft = np.exp(logp)
    onehot = np.zeros_like(soft)
    onehot[np.arange(B), e] = 1
    g_logits = (dt(hp.c_env) * w_c * invB)[:, None] * (soft - onehot)            # [B,K]
    # functions.py:13-16: reversal (-alpha) only on the classifier branch
    g_p = g_z1[:, None] + dt(-hp.alpha) * (g_logits @ W)                          # [B,D]

    def R(x, div):
        return (dt(2) * dt(hp.c_L2) * x + dt(hp.c_L1) * np.sign(x)) / div

    g = {k: np.zeros_like(p[k], dtype=dt) for k in PARAM_ORDER}
    np.add.at(g["Uinv"], u, g_p * c + R(a, bd2))
    np.add.at(g["Iinv"], i, g_p * a + R(c, bd2))
    np.add.at(g["Uenv"], u, g_z2[:, None] * ie * ee + R(ue, bd2))
    np.add.at(g["Ienv"], i, g_z2[:, None] * ue * ee + R(ie, bd2))
    ge = g_z2[:, None] * ue * ie
    if flags.reg_env_embed:
        ge = ge + R(ee, bd)
    np.add.at(g["E"], e, ge)
    g["W"] = g_logits.T @ pref                       # NOT reversed
    g["b"] = g_logits.sum(axis=0, dtype=dt)
    if not flags.reg_only_embed:
        g["W"] = g["W"] + R(W, dt(D * K))
        g["b"] = g["b"] + R(b, dt(K))
    g = {k: v.astype(dt) for k, v in g.items()}
    return losses, g


def new_adam_state(p, dt=np.float32):
    return {"step": 0, "m": {k: np.zeros_like(p[k], dtype=dt) for k in PARAM_ORDER},
            "v": {k: np.zeros_like(p[k], dtype=dt) for k in PARAM_ORDER}}


def adam_step(p, g, st, hp: Hyper, dt=np.float32):
    """torch.optim.Adam single-tensor path as configured at train.py:718 (amsgrad off, wd 0).

    Bias-correction scalars are python float64 and enter the tensor ops as scalars of the
    tensor dtype, as in torch/optim/adam.py (_single_tensor_adam).
    """
    st["step"] += 1
    t = st["step"]
    bc1 = 1.0 - hp.beta1 ** t
    bc2 = 1.0 - hp.beta2 ** t
    step_size = dt(hp.lr / bc1)
    bc2_sqrt = dt(math.sqrt(bc2))
    for k in PARAM_ORDER:
        m, v, gg = st["m"][k], st["v"][k], g[k].astype(dt)
        m += (gg - m) * dt(1.0 - hp.beta1)                          # lerp_
        v *= dt(hp.beta2)
        v += dt(1.0 - hp.beta2) * gg * gg                           # addcmul_
        denom = np.sqrt(v) / bc2_sqrt + dt(hp.eps)
        p[k] = (p[k].astype(dt) - step_size * (m / denom)).astype(dt)   # addcdiv_
    return p, st


def train_step(p, st, u, i, y, e, w, hp: Hyper, flags: Flags, dt=np.float32):
    """One train_a_batch (train.py:771-844): returns loss dict, grads; updates p/st in place."""
    losses, g = loss_and_grads(p, u, i, y, e, w, hp, flags, dt)
    adam_step(p, g, st, hp, dt)
    return losses, g


def init_eps(K: int) -> np.ndarray:
    """train.py:763-769: [K!, K] fp32 permutations of 1e-10 * 0.1**k (values rounded to fp32 first)."""
    base = [1e-10 * (1e-1 ** k) for k in range(K)]
    temp = np.asarray(base, dtype=np.float32)
    return np.asarray(list(itertools.permutations(temp.tolist())), dtype=np.float32)


def cluster_distances(p, u, i, y, flags: Flags, dt=np.float32):
    """train.py:853-866: distance of every sample under each of the K environments -> [b, K]."""
    K = p["E"].shape[0]
    y = np.asarray(y, dtype=dt)
    cols = []
    for k in range(K):
        e = np.full(len(u), k, dtype=np.int64)
        _, s_env, _, _ = forward(p, u, i, e, flags, dt)
        cols.append(_rec_loss(s_env, y, flags.implicit, dt))
    return np.stack(cols, axis=1).astype(dt)


def cluster_batch(p, u, i, y, flags: Flags, perm_idx=None, eps_table=None, dt=np.float32):
    """train.py:846-879: argmin over envs (first-min on ties), optional eps perturbation."""
    dist = cluster_distances(p, u, i, y, flags, dt)
    if perm_idx is not None:
        dist = dist + eps_table.astype(dt)[perm_idx]                # train.py:872-873
    return np.argmin(dist, axis=1).astype(np.int64), dist


def near_tie_mask(dist, rel=1e-5, abs_=1e-12):
    """Samples whose two smallest distances are closer than the fp32 summation noise
    (the "argmin ties" the north star exempts and counts)."""
    s = np.sort(dist.astype(np.float64), axis=1)
    gap = s[:, 1] - s[:, 0]
    return gap <= np.maximum(rel * np.abs(s[:, 0]), abs_)


def stat_envs(envs, K: int, N: int):
    """train.py:945-957: counts, class_weights (float64 -> fp32), sample_weights."""
    cnt = np.bincount(envs, minlength=K)[:K].astype(np.int64)
    rate = np.minimum(cnt + 1, N - 1).astype(np.float64) / N
    cw = rate.astype(np.float32)
    return cnt, cw, cw[envs]


def alpha_schedule(batch_index: int, epoch_cnt: int, batch_num: int) -> float:
    """train.py:891-894."""
    pp = float(batch_index + (epoch_cnt + 1) * batch_num) / float((epoch_cnt + 1) * batch_num)
    return float(2.0 / (1.0 + np.exp(-10.0 * pp)) - 1.0)


def mini_batch_bounds(N: int, B: int):
    """utils.py:12-19: sequential, unshuffled, last batch short."""
    return [(s, min(s + B, N)) for s in range(0, N, B)]


def stable_segments(ids):
    """Reference for the segment builder: stable sort permutation, unique rows, offsets."""
    ids = np.asarray(ids)
    perm = np.argsort(ids, kind="stable").astype(np.int64)
    srt = ids[perm]
    if len(ids) == 0:
        return perm, srt[:0], np.zeros(1, dtype=np.int64)
    starts = np.flatnonzero(np.r_[True, srt[1:] != srt[:-1]])
    return perm, srt[starts], np.r_[starts, len(ids)].astype(np.int64)
