"""Torch-eager CPU restatement of the reference trainer's hot path.  TEST INFRASTRUCTURE ONLY.

This is the "port" that ``bench.py`` times as the CPU baseline (``cpu_baseline.kind == "port"`` and
``--impl reference``): it issues the same ATen op sequence per step as the reference does --
five embedding gathers for the forward (models.py:449-455), 4-5 more for each of the L2 and L1
regularisers (models.py:469-532), autograd through a gradient-reversal node (functions.py:4-16),
``embedding_dense_backward`` for every gather, and a dense ``torch.optim.Adam`` step (train.py:718,
832-834) -- so its CPU cost is representative of the reference's own.  The python reference itself
cannot travel to the GPU box (``/root/reference`` is not there and its sources may not be copied).

Pinned against the golden fixtures by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import itertools

import numpy as np
import torch
import torch.nn.functional as F

from .invpref_numpy import LOSS_KEYS, PARAM_ORDER, STATE_KEYS, Flags, Hyper


class _Reverse(torch.autograd.Function):
    """functions.py:4-16."""

    @staticmethod
    def forward(ctx, x, alpha):
        ctx.alpha = alpha
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.neg() * ctx.alpha, None


def params_from_state_dict(sd, dtype=torch.float32):
    return {k: torch.as_tensor(np.asarray(sd[v])).to(dtype).clone().requires_grad_(True) for k, v in STATE_KEYS.items()}


def random_params(U, I, K, D, seed=17373331, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    shapes = {"Uinv": (U, D), "Iinv": (I, D), "Uenv": (U, D), "Ienv": (I, D), "E": (K, D), "W": (K, D), "b": (K,)}
    out = {}
    for k in PARAM_ORDER:
        std = 0.01 if k not in ("W", "b") else 0.1
        out[k] = (torch.randn(shapes[k], generator=g, dtype=dtype) * std).requires_grad_(True)
    return out


def forward(P, u, i, e, alpha, implicit: bool):
    """models.py:448-467 / 307-326."""
    a, c = F.embedding(u, P["Uinv"]), F.embedding(i, P["Iinv"])
    ue, ie, ee = F.embedding(u, P["Uenv"]), F.embedding(i, P["Ienv"]), F.embedding(e, P["E"])
    pref = a * c
    q = ue * ie * ee
    s_inv, mid = pref.sum(dim=1), q.sum(dim=1)
    if implicit:
        s_inv, mid = torch.sigmoid(s_inv), torch.sigmoid(mid)
        s_env = s_inv * mid
    else:
        s_env = s_inv + mid
    logp = F.log_softmax(F.linear(_Reverse.apply(pref, alpha), P["W"], P["b"]), dim=1)
    return s_inv.reshape(-1), s_env.reshape(-1), logp.reshape(-1, P["W"].shape[0])


def _reg(P, u, i, e, flags: Flags, norm: int):
    """models.py:469-532: the rows are gathered AGAIN for each norm, as the reference does."""
    D = P["Uinv"].shape[1]
    K = P["W"].shape[0]

    def n(x):
        return x.norm(2).pow(2) if norm == 2 else x.norm(1)

    r = (n(F.embedding(u, P["Uenv"])) + n(F.embedding(u, P["Uinv"]))) / (float(len(u)) * float(D) * 2) \
        + (n(F.embedding(i, P["Ienv"])) + n(F.embedding(i, P["Iinv"]))) / (float(len(i)) * float(D) * 2)
    if not flags.reg_only_embed:
        r = n(P["W"]) / float(D * K) + n(P["b"]) / float(K) + r
    if flags.reg_env_embed:
        r = r + n(F.embedding(e, P["E"])) / (float(len(e)) * float(D))
    return r


class CpuTrainer:
    """train.py:693-957 reduced to the hot path: train_a_batch, cluster_a_batch, stat_envs."""

    def __init__(self, P, flags: Flags, hp: Hyper):
        self.P, self.flags, self.hp = P, flags, hp
        self.opt = torch.optim.Adam([P[k] for k in PARAM_ORDER], lr=hp.lr, betas=(hp.beta1, hp.beta2), eps=hp.eps)
        K = P["E"].shape[0]
        base = torch.Tensor([1e-10 * (1e-1 ** k) for k in range(K)])
        self.eps_table = torch.Tensor(list(itertools.permutations(base)))           # train.py:763-769

    def train_a_batch(self, u, i, y, e, w, alpha) -> dict:
        hp, fl = self.hp, self.flags
        s_inv, s_env, logp = forward(self.P, u, i, e, alpha, fl.implicit)
        rec = F.binary_cross_entropy if fl.implicit else F.mse_loss
        if hp.use_rec_rw:                                                            # train.py:817-819
            inv_loss = torch.mean(rec(s_inv, y, reduction="none") * w)
            ea_loss = torch.mean(rec(s_env, y, reduction="none") * w)
        else:
            inv_loss, ea_loss = rec(s_inv, y), rec(s_env, y)
        if hp.use_class_rw:                                                          # train.py:814-815
            envs_loss = torch.mean(F.nll_loss(logp, e, reduction="none") * w)
        else:
            envs_loss = F.nll_loss(logp, e)
        L2 = _reg(self.P, u, i, e, fl, 2)
        L1 = _reg(self.P, u, i, e, fl, 1)
        loss = inv_loss * hp.c_inv + ea_loss * hp.c_ea + envs_loss * hp.c_env + L2 * hp.c_L2 + L1 * hp.c_L1
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        vals = (inv_loss, ea_loss, envs_loss, L2, L1, loss)
        return {k: float(v.detach()) for k, v in zip(LOSS_KEYS, vals)}

    def cluster_a_batch(self, u, i, y, perm_idx=None):
        """train.py:846-879: K full forwards (no no_grad, like the reference), cat, eps, argmin."""
        K = self.P["E"].shape[0]
        rec = F.binary_cross_entropy if self.flags.implicit else F.mse_loss
        cols = []
        for k in range(K):
            ek = torch.full((len(u),), k, dtype=torch.int64)
            _, s_env, _ = forward(self.P, u, i, ek, 0.0, self.flags.implicit)
            cols.append(rec(s_env, y, reduction="none").reshape(-1, 1))
        dist = torch.cat(cols, dim=1)
        if perm_idx is not None:
            dist = dist + self.eps_table[perm_idx]
        return torch.argmin(dist, dim=1)

    @staticmethod
    def stat_envs(envs, K, N):
        """train.py:945-957."""
        rate = np.zeros(K)
        cnts = {}
        for k in range(K):
            c = int(torch.sum(envs == k))
            cnts[k] = c
            rate[k] = min(c + 1, N - 1)
        cw = torch.Tensor(rate / N)
        return cnts, cw, cw[envs]
