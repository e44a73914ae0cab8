"""GPU: the driver mains run end to end (model + evaluator + trainer as in the reference's main()) and the
evaluators agree with straightforward restatements of the reference's metric definitions."""
import numpy as np
import pytest
import torch

from _golden import Golden

pytestmark = pytest.mark.gpu


def test_coat_explicit_driver_main_on_real_interactions():
    """Coat explicit (the reference's own CPU-runnable config), real train interactions from the fixture."""
    from invpref_kdd_2022_b200.dataloader import ExplicitDataLoader
    from invpref_kdd_2022_b200.drivers import Coat_InvPref_explicit as drv
    g = Golden("coat_explicit")
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    test = g.data[rng.permutation(g.N)[:1500]]
    loader = ExplicitDataLoader("", dev, train=g.data, test=test)
    assert (loader.user_num, loader.item_num) == (290, 300)
    tc = dict(drv.TRAIN_CONFIG, epochs=40, evaluate_interval=10, cluster_interval=20)
    best, idx, res = drv.main(dev, drv.MODEL_CONFIG, tc, drv.EVALUATE_CONFIG, loader, 17373331, silent=True,
                              auto=True, query=False)
    assert set(res) == {"mse", "rmse", "mae"} and np.isfinite(best)
    assert abs(res["rmse"] ** 2 - res["mse"]) < 1e-6 * res["mse"]
    assert idx[0] > 0                      # training lowers the test MSE below the epoch-0 value


def _reference_style_metrics(rating, mask_sets, gt_sets, ks):
    """evaluate.py:11-56 restated with python sets / numpy on the host."""
    rating = rating.copy()
    for r, s in enumerate(mask_sets):
        rating[r, list(s)] = -(1 << 10)
    top = np.argsort(-rating, axis=1, kind="stable")[:, :max(ks)]
    out = {"ndcg": {}, "recall": {}, "precision": {}}
    for k in ks:
        pre = rec = nd = 0.0
        for r, gt in enumerate(gt_sets):
            hit = np.array([1.0 if it in gt else 0.0 for it in top[r, :k]])
            pre += hit.sum() / k
            rec += hit.sum() / len(gt)
            idcg = sum(1.0 / np.log2(j + 2) for j in range(min(k, len(gt))))
            nd += (hit / np.log2(np.arange(2, k + 2))).sum() / (idcg if idcg > 0 else 1.0)
        n = len(gt_sets)
        out["ndcg"][k], out["recall"][k], out["precision"][k] = nd / n, rec / n, pre / n
    return out


def test_implicit_evaluator_matches_metric_definitions():
    from invpref_kdd_2022_b200.dataloader import YahooImplicitBCELossDataLoader, synthetic_interactions
    from invpref_kdd_2022_b200.evaluate import ImplicitTestManager
    from invpref_kdd_2022_b200.models import InvPrefImplicit
    dev = torch.device("cuda:0")
    tr = synthetic_interactions(120, 90, 4000, True, seed=1)
    te = synthetic_interactions(120, 90, 600, True, seed=2)
    te = te[te[:, 2] > 0]
    loader = YahooImplicitBCELossDataLoader("", dev, train=tr, test=te)
    torch.manual_seed(0)
    model = InvPrefImplicit(loader.user_num, loader.item_num, 2, 16).to(dev)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(30.0)
    ev = ImplicitTestManager(model, loader, test_batch_size=50, top_k_list=[7, 3, 5])
    got = ev.evaluate()
    users = loader.all_test_users_by_sorted_list
    assert users == sorted(set(te[:, 0].tolist()))
    rating = model.predict(loader.all_test_users_by_sorted_tensor).detach().cpu().numpy()
    ref = _reference_style_metrics(rating, [loader.user_mask_items(u) for u in users],
                                   loader.get_sorted_all_test_users_ground_truth, [3, 5, 7])
    for m in ref:
        assert list(got[m]) == [3, 5, 7]
        for k in ref[m]:
            assert abs(got[m][k] - ref[m][k]) < 1e-9, (m, k)


def test_implicit_driver_main_synthetic():
    from invpref_kdd_2022_b200.dataloader import YahooImplicitBCELossDataLoader, synthetic_interactions
    from invpref_kdd_2022_b200.drivers import Yahoo_InvPref_Implicit as drv
    dev = torch.device("cuda:0")
    tr = synthetic_interactions(500, 200, 30000, True, seed=3)
    te = synthetic_interactions(500, 200, 3000, True, seed=4)
    te = te[te[:, 2] > 0]
    loader = YahooImplicitBCELossDataLoader("", dev, train=tr, test=te)
    # the reference's Yahoo implicit main ranks only the test item pool (Yahoo_InvPref_Implicit.py:87, :207)
    assert drv.USE_ITEM_POOL and drv.HAS_ITEM_POOL_FILE
    from invpref_kdd_2022_b200.dataloader import synthetic_item_pool
    loader.set_item_pool(synthetic_item_pool(te, 500, 200))
    tc = dict(drv.TRAIN_CONFIG, epochs=6, evaluate_interval=3, cluster_interval=2)
    best, idx, res = drv.main(dev, drv.MODEL_CONFIG, tc, drv.EVALUATE_CONFIG, loader, 17373331, silent=True,
                              auto=True, query=False)                 # the reference's 3-tuple
    assert 0.0 <= best <= 1.0 and len(idx) >= 1
    ks = drv.EVALUATE_CONFIG['top_k_list']
    assert set(res) == {f"{m}@{k}" for m in ("ndcg", "recall", "precision") for k in ks}
    assert res[f"{drv.EVALUATE_CONFIG['eval_metric']}@{drv.EVALUATE_CONFIG['eval_k']}"] == best


def test_implicit_evaluator_device_path_equals_host_path_with_item_pool():
    """invpref_mask_scores / invpref_hits_from_csr (CSR lists resident on the device) against the dense-mask
    path that tests/test_ref_loaders_evaluators.py pins to the live reference, with the item pool on."""
    from invpref_kdd_2022_b200.dataloader import YahooImplicitBCELossDataLoader, _csr, synthetic_interactions
    from invpref_kdd_2022_b200.evaluate import ImplicitTestManager
    from invpref_kdd_2022_b200.models import InvPrefImplicit
    dev = torch.device("cuda:0")
    tr = synthetic_interactions(300, 211, 9000, True, seed=5)
    te = synthetic_interactions(300, 211, 1500, True, seed=6)
    te = te[te[:, 2] > 0]
    loader = YahooImplicitBCELossDataLoader("", dev, train=tr, test=te)
    pool = synthetic_interactions(300, 211, 6000, True, seed=7)
    loader.has_item_pool = True
    loader.pool_off, loader.pool_items = _csr(pool[:, 0], pool[:, 1], loader.user_num)
    torch.manual_seed(1)
    model = InvPrefImplicit(loader.user_num, loader.item_num, 2, 24).to(dev)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(30.0)
    for use_pool in (False, True):
        ev = ImplicitTestManager(model, loader, test_batch_size=64, top_k_list=[10, 3, 5], use_item_pool=use_pool)
        users_t, users_l = loader.all_test_users_by_sorted_tensor, loader.all_test_users_by_sorted_list
        a = ev.evaluate_batch(users_t, users_l)
        b = ev.evaluate_batch(users_t, users_l, force_host=True)
        for m in ("ndcg", "recall", "precision"):
            assert np.allclose(a[m], b[m], rtol=0, atol=1e-9), (use_pool, m, a[m], b[m])


@pytest.mark.parametrize("n_items,dim", [(211, 24), (3706, 40), (60000, 40), (997, 30)])
@pytest.mark.parametrize("use_pool", [False, True])
def test_fused_evaluator_kernel_matches_the_step_by_step_path(n_items, dim, use_pool):
    """invpref_eval_topk (scores + mask + pool + exact top-k + hits in one kernel, no rating matrix) against
    model.predict -> mask kernels -> torch.topk -> hit kernel, which test_ref_loaders_evaluators.py pins to the live
    reference.  Item sets that fit the shared-memory score cache (211, 997, 3706) and one that does not (60000:
    bitmap + recompute path); D = 30 takes the scalar dot product.  The selected ITEMS must be the same wherever the
    k-th and (k+1)-th ratings are distinguishable (the fused kernel sums the dot product in a different order than
    the GEMM: ratings within 1e-6 may swap ranks), and the metric dictionaries must agree."""
    from invpref_kdd_2022_b200.dataloader import YahooImplicitBCELossDataLoader, synthetic_interactions
    from invpref_kdd_2022_b200.dataloader import synthetic_item_pool
    from invpref_kdd_2022_b200.evaluate import ImplicitTestManager
    from invpref_kdd_2022_b200.models import InvPrefImplicit
    dev = torch.device("cuda:0")
    U = 300
    tr = synthetic_interactions(U, n_items, 9000, True, seed=5)
    te = synthetic_interactions(U, n_items, 1500, True, seed=6)
    te = te[te[:, 2] > 0]
    loader = YahooImplicitBCELossDataLoader("", dev, train=tr, test=te)
    if use_pool:
        loader.set_item_pool(synthetic_item_pool(te, U, n_items, extra=min(50, n_items // 3)))
    torch.manual_seed(1)
    model = InvPrefImplicit(loader.user_num, loader.item_num, 2, dim).to(dev)
    with torch.no_grad():
        for p in model.parameters():
            p.mul_(30.0)
    ks = [40, 3, 5]
    users_t, users_l = loader.all_test_users_by_sorted_tensor, loader.all_test_users_by_sorted_list
    ev_f = ImplicitTestManager(model, loader, test_batch_size=64, top_k_list=ks, use_item_pool=use_pool, fused=True)
    ev_s = ImplicitTestManager(model, loader, test_batch_size=64, top_k_list=ks, use_item_pool=use_pool, fused=False)
    hits, n_gt, top = ev_f._hits_fused(users_t, 40)
    # the step-by-step ratings, adjusted the same way
    rating = model.predict(users_t).clone().contiguous()
    h2, n2 = ev_s._hits_device(rating, users_t.to(dev), 40)          # masks `rating` in place, then top-k + hits
    vals, top_ref = torch.topk(rating, k=40)
    got_vals = torch.gather(rating, 1, top)
    assert torch.equal(n_gt, n2)
    assert float((got_vals - vals).abs().max()) <= 2e-6               # same rating at every rank ...
    same = (top == top_ref)
    gap_ok = (vals[:, :-1] - vals[:, 1:]).abs() > 4e-6                # ... and the same item wherever ranks are distinct
    clear = torch.ones_like(same)
    clear[:, :-1] &= gap_ok
    clear[:, 1:] &= gap_ok
    kth = torch.topk(rating, k=41).values[:, 40:41] if n_items > 40 else None
    if kth is not None:
        clear[:, -1:] &= (vals[:, -1:] - kth).abs() > 4e-6
    assert bool(same[clear].all())
    assert float(clear.double().mean()) > 0.5
    # descending order with ties broken by the lower item id
    assert bool((got_vals[:, :-1] >= got_vals[:, 1:]).all())
    a, b = ev_f.evaluate(), ev_s.evaluate()
    for m in ("ndcg", "recall", "precision"):
        for k in ks:
            assert abs(a[m][k] - b[m][k]) <= 2e-3, (m, k, a[m][k], b[m][k])


def test_fused_evaluator_ties_go_to_the_lowest_item_id():
    """All-equal ratings (zero tables): torch.topk leaves the order of ties unspecified, the fused kernel takes the
    lowest ids in ascending order; masked items drop out, pool items come first."""
    from invpref_kdd_2022_b200.dataloader import YahooImplicitBCELossDataLoader
    from invpref_kdd_2022_b200.evaluate import ImplicitTestManager
    from invpref_kdd_2022_b200.models import InvPrefImplicit
    dev = torch.device("cuda:0")
    tr = np.array([[0, 0, 1], [0, 2, 1], [1, 1, 1], [2, 699, 0]], dtype=np.int64)
    te = np.array([[0, 5, 1], [1, 0, 1], [2, 3, 1]], dtype=np.int64)
    loader = YahooImplicitBCELossDataLoader("", dev, train=tr, test=te)
    loader.set_item_pool(np.array([[0, 650], [0, 9], [1, 1]], dtype=np.int64))
    model = InvPrefImplicit(loader.user_num, loader.item_num, 2, 16).to(dev)
    with torch.no_grad():
        for p in model.parameters():
            p.zero_()
    ev = ImplicitTestManager(model, loader, test_batch_size=8, top_k_list=[6], use_item_pool=True)
    hits, n_gt, top = ev._hits_fused(loader.all_test_users_by_sorted_tensor, 6)
    top = top.cpu().tolist()
    assert top[0] == [9, 650, 1, 3, 4, 5]          # pool first (ascending), then the unmasked ids 1, 3, 4, 5 (0, 2 masked)
    assert top[1] == [0, 2, 3, 4, 5, 6]            # item 1 is masked AND in the pool: -1024 + 1024 = 0 < 0.5
    assert top[2] == [0, 1, 2, 3, 4, 5]
    assert hits.cpu().tolist() == [[0, 0, 0, 0, 0, 1], [1, 0, 0, 0, 0, 0], [0, 0, 0, 1, 0, 0]]
    assert n_gt.cpu().tolist() == [1, 1, 1]


@pytest.mark.parametrize("name", ["MovieLens_InvPref", "MIND_InvPref", "Yahoo_InvPref_explicit"])
def test_remaining_driver_mains_run_on_synthetic_data(name):
    """The MovieLens / MIND / Yahoo-explicit mains (reference MovieLens_InvPref.py, MIND_InvPref.py,
    Yahoo_InvPref_explicit.py main()) with their own MODEL / TRAIN / EVALUATE configs, on synthetic interactions of a
    reduced shape (their train.csv files are not in the reference checkout / too large for a unit test): epochs run
    as CUDA graphs from the second one on, cluster() + stat_envs() in between, evaluator as the driver configures it
    (MIND: test item pool; MovieLens: all items; Yahoo explicit: MSE)."""
    import importlib
    from invpref_kdd_2022_b200 import dataloader as dl
    drv = importlib.import_module("invpref_kdd_2022_b200.drivers." + name)
    dev = torch.device("cuda:0")
    implicit = name != "Yahoo_InvPref_explicit"
    U, I, N = 800, 300, 40000
    tr = dl.synthetic_interactions(U, I, N, implicit, seed=21)
    te = dl.synthetic_interactions(U, I, 4000, implicit, seed=22)
    if implicit:
        te = te[te[:, 2] > 0]
        loader = dl.YahooImplicitBCELossDataLoader("", dev, train=tr, test=te)
        if drv.HAS_ITEM_POOL_FILE:
            loader.set_item_pool(dl.synthetic_item_pool(te, U, I))
    else:
        loader = dl.ExplicitDataLoader("", dev, train=tr, test=te)
    tc = dict(drv.TRAIN_CONFIG, epochs=6, evaluate_interval=3, cluster_interval=2, batch_size=8192)
    best, idx, res = drv.main(dev, drv.MODEL_CONFIG, tc, drv.EVALUATE_CONFIG, loader, 17373331, silent=True, auto=True,
                              query=False)
    assert np.isfinite(best) and len(idx) >= 1
    if implicit:
        ks = drv.EVALUATE_CONFIG["top_k_list"]
        assert set(res) == {f"{m}@{k}" for m in ("ndcg", "recall", "precision") for k in ks}
        assert 0.0 <= best <= 1.0
    else:
        assert set(res) == {"mse", "rmse", "mae"}
