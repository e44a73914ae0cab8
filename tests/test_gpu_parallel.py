"""GPU (one device): the multi-GPU step code run as G simulated ranks in one process (SimDriver executes the
collectives the step generators yield) must reproduce the single-GPU result on the same global batches:
losses, every table row, env assignments.  This is the "fake collective" test of SURVEY.md §4; the real NCCL
path runs the same generators under DistDriver (tests/test_dist_gloo.py covers the drivers on CPU)."""
import numpy as np
import pytest
import torch

from oracle import invpref_numpy as on

pytestmark = pytest.mark.gpu
TOL = 1e-5


def nerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


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


def single_gpu(p, batches, implicit, roe, ree, lr):
    from invpref_kdd_2022_b200.engine import HotPath
    dev = torch.device("cuda:0")
    hp = HotPath({k: torch.tensor(v, device=dev) for k, v in p.items()}, implicit, roe, ree, lr=lr)
    t = lambda a: torch.tensor(a, device=dev)
    losses = [hp.train_step(t(u), t(i), t(y), t(e), t(w), **KW).cpu().numpy().copy() for (u, i, y, e, w) in batches]
    return hp, np.asarray(losses)


@pytest.mark.parametrize("p2p", [False, True])
@pytest.mark.parametrize("lazy", [True, False])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("implicit,roe,ree", [(False, True, False), (True, False, True)])
def test_sharded_matches_single_gpu(world, implicit, roe, ree, lazy, p2p):
    from invpref_kdd_2022_b200.parallel import ShardedTrainer, SimDriver
    dev = torch.device("cuda:0")
    U, I, K, D, B = 1000, 203, 4, 64, 20000
    u, i, y, e, w, p = synth(U, I, 2 * B + 777, K, D, implicit, 5)
    bounds = on.mini_batch_bounds(len(u), B)                               # 3 batches, the last one short
    batches = [(u[a:b], i[a:b], y[a:b], e[a:b], w[a:b]) for a, b in bounds]
    ref, ref_losses = single_gpu(p, batches, implicit, roe, ree, 1e-2)
    init = {k: torch.tensor(v, device=dev) for k, v in p.items()}
    ranks = [ShardedTrainer(U, I, K, D, implicit, roe, ree, 1e-2, r, world, dev, cache_rows=I, init=init, lazy=lazy)
             for r in range(world)]
    if p2p:   # peer-memory exchange: the simulated ranks' buffers are plain tensors on the same device
        for rk in ranks:
            rk.enable_p2p([x.Iinv.data_ptr() for x in ranks], [x.Ienv.data_ptr() for x in ranks],
                          [x.gcache[0].data_ptr() for x in ranks], [x.gcache[1].data_ptr() for x in ranks])
    sim = SimDriver(world)
    t = lambda a: torch.tensor(a, device=dev)
    losses = []
    for (bu, bi, by, be, bw) in batches:
        sbs = sim.run_all([rk.prepare_gen(t(bu), t(bi), t(by)) for rk in ranks])
        assert sum(sb.sel.numel() for sb in sbs) == len(bu)
        outs = sim.run_all([rk.step_gen(sb, t(be)[sb.sel].contiguous(), t(bw)[sb.sel].contiguous(), **KW)
                            for rk, sb in zip(ranks, sbs)])
        for o in outs[1:]:
            assert torch.equal(o, outs[0])                                  # replicated results identical
        losses.append(outs[0].cpu().numpy().copy())
    assert np.abs(np.asarray(losses) - ref_losses).max() <= TOL * np.abs(ref_losses).max()
    full = {k: ref.params[k].cpu().numpy() for k in on.PARAM_ORDER}
    for r, rk in enumerate(ranks):
        loc = {k: v.cpu().numpy() for k, v in rk.local_tables().items()}
        # three Adam steps with a different (but fixed) summation order of the item partials: Adam's
        # g/(|g|+eps) amplification of tiny gradients allows ~1e-4 (multi-step drift)
        for k in ("Uinv", "Uenv", "Iinv", "Ienv"):
            assert nerr(loc[k], full[k][r::world]) <= 2e-4, (k, r)
        for k in ("E", "W", "b"):
            assert nerr(loc[k], full[k]) <= 2e-4, (k, r)
    # EM re-assignment on the last batch: the union of the ranks' results equals the single-GPU result
    bu, bi, by, be, bw = batches[-1]
    eps = torch.tensor(on.init_eps(K), device=dev)
    pidx = torch.tensor(np.random.default_rng(0).integers(0, eps.shape[0], len(bu)), device=dev)
    new_ref, hist_ref, diff_ref = ref.cluster(t(bu), t(bi), t(by), pidx, eps, t(be))
    res = sim.run_all([rk.cluster_gen(sb, pidx[sb.sel].contiguous(), eps, t(be)[sb.sel].contiguous())
                       for rk, sb in zip(ranks, sbs)])
    new = torch.empty_like(new_ref)
    for sb, (nv, hist, diff) in zip(sbs, res):
        new[sb.sel] = nv
    mism = (new != new_ref).cpu().numpy()
    dist = on.cluster_distances(full, bu, bi, by, on.Flags(implicit, roe, ree))
    assert not (mism & ~on.near_tie_mask(dist)).any()
    assert sum(int(h.sum()) for _, h, _ in res) == len(bu)


@pytest.mark.parametrize("world", [2, 4])
def test_replicated_matches_single_gpu(world):
    from invpref_kdd_2022_b200.parallel import ReplicatedTrainer, SimDriver
    dev = torch.device("cuda:0")
    U, I, K, D, B = 300, 120, 6, 40, 9000
    u, i, y, e, w, p = synth(U, I, 2 * B + 100, K, D, True, 9)
    bounds = on.mini_batch_bounds(len(u), B)
    batches = [(u[a:b], i[a:b], y[a:b], e[a:b], w[a:b]) for a, b in bounds]
    ref, ref_losses = single_gpu(p, batches, True, False, True, 1e-2)
    ranks = [ReplicatedTrainer({k: torch.tensor(v, device=dev) for k, v in p.items()}, True, False, True, 1e-2, r,
                               world) for r in range(world)]
    sim = SimDriver(world)
    t = lambda a: torch.tensor(a, device=dev)
    losses = []
    for (bu, bi, by, be, bw) in batches:
        gens = []
        for rk in ranks:
            a, b = rk.chunk(0, len(bu))
            gens.append(rk.step_gen(t(bu[a:b]), t(bi[a:b]), t(by[a:b]), t(be[a:b]), t(bw[a:b]), len(bu), **KW))
        outs = sim.run_all(gens)
        losses.append(outs[0].cpu().numpy().copy())
    assert np.abs(np.asarray(losses) - ref_losses).max() <= TOL * np.abs(ref_losses).max()
    for rk in ranks:
        assert torch.equal(rk.flat, ranks[0].flat)
        for k in on.PARAM_ORDER:
            assert nerr(rk.params[k].cpu().numpy(), ref.params[k].cpu().numpy()) <= 2e-4, k


def test_sharded_peer_memory_is_bitwise_the_collective_path():
    """The peer-memory exchange (invpref_fetch_rows_p2p + invpref_owner_adam_p2p) sums the partial item
    gradients in the same rank order with the same arithmetic as all-to-all + scatter-add + adam_dense."""
    from invpref_kdd_2022_b200.parallel import ShardedTrainer, SimDriver
    dev = torch.device("cuda:0")
    world, U, I, K, D, B = 3, 700, 151, 4, 64, 12000
    u, i, y, e, w, p = synth(U, I, 3 * B, K, D, False, 11)
    init = {k: torch.tensor(v, device=dev) for k, v in p.items()}
    t = lambda a: torch.tensor(a, device=dev)
    out = []
    for p2p in (False, True):
        ranks = [ShardedTrainer(U, I, K, D, False, True, False, 1e-2, r, world, dev, cache_rows=I, init=init)
                 for r in range(world)]
        if p2p:
            for rk in ranks:
                rk.enable_p2p([x.Iinv.data_ptr() for x in ranks], [x.Ienv.data_ptr() for x in ranks],
                              [x.gcache[0].data_ptr() for x in ranks], [x.gcache[1].data_ptr() for x in ranks])
        sim = SimDriver(world)
        losses = []
        sbs_all = []
        for s in range(3):
            sl = slice(s * B, (s + 1) * B)
            sbs_all.append(sim.run_all([rk.prepare_gen(t(u[sl]), t(i[sl]), t(y[sl])) for rk in ranks]))
        for s in range(3):
            sl = slice(s * B, (s + 1) * B)
            sbs = sbs_all[s]
            nxt = sbs_all[s + 1] if s + 1 < 3 else [None] * world
            res = sim.run_all([rk.step_gen(sb, t(e[sl])[sb.sel].contiguous(), t(w[sl])[sb.sel].contiguous(),
                                           next_sb=nx, **KW) for rk, sb, nx in zip(ranks, sbs, nxt)])
            losses.append(res[0].clone())
        out.append((losses, [{k: v.clone() for k, v in rk.local_tables().items()} for rk in ranks]))
    for a, b in zip(out[0][0], out[1][0]):
        assert torch.equal(a, b)
    for ta, tb in zip(out[0][1], out[1][1]):
        for k in ta:
            assert torch.equal(ta[k], tb[k]), k
