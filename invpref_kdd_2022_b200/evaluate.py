"""Evaluators with the reference's result formats (reference ``evaluate.py:59-212``), off the hot path
(SURVEY.md §8f ranks 1 and 4).

``ExplicitTestManager``: MSE / RMSE / MAE of ``model.predict(test_users, test_items)`` (the fused
predict kernel).  ``ImplicitTestManager``: recall / precision / NDCG @k per test user; the reference builds
python index lists per user and counts hits with python sets (evaluate.py:94-135).  On a CUDA model the per-user
item lists (train positives, item pool, ground truth) are uploaded ONCE as CSR and two library kernels do the
masking and the hit look-ups (``invpref_mask_scores`` / ``invpref_hits_from_csr``): no per-batch host work at all;
otherwise (CPU stub models in the reference-parity tests) dense boolean matrices are built per batch.  The
metrics are the same tensor expressions on both paths.
"""
from __future__ import annotations

import numpy as np
import torch

from .utils import mini_batch


class ExplicitTestManager:
    """reference evaluate.py:178-212."""

    def __init__(self, model, data_loader):
        self.model, self.data_loader = model, data_loader

    def evaluate(self) -> dict:
        self.model.eval()
        pairs = self.data_loader.all_test_pairs_tensor
        scores = self.data_loader.all_test_scores_tensor
        with torch.no_grad():
            pred = self.model.predict(pairs[:, 0].contiguous(), pairs[:, 1].contiguous())
            err = scores - pred
            mse = torch.mean(err * err)
            out = torch.stack([mse, torch.sqrt(mse), torch.mean(err.abs())]).cpu().tolist()
        return {'mse': out[0], 'rmse': out[1], 'mae': out[2]}


class ImplicitTestManager:
    """reference evaluate.py:59-175; result = {'ndcg': {k: v}, 'recall': {k: v}, 'precision': {k: v}}."""

    def __init__(self, model, data_loader, test_batch_size: int, top_k_list: list, use_item_pool: bool = False,
                 fused: bool = True):
        """``fused``: on an InvPref model that lives on the GPU the whole batch -- scores, masking, pool highlight,
        top-k, hit look-up -- is ONE library kernel (``invpref_eval_topk``) and the [b, item_num] rating matrix is
        never materialised; ``fused=False`` keeps the step-by-step device path (``model.predict`` -> mask kernels ->
        ``torch.topk`` -> hit kernel)."""
        self.model, self.data_loader = model, data_loader
        self.batch_size = test_batch_size
        self.top_k_list = sorted(top_k_list)
        self.use_item_pool = use_item_pool
        self.fused = fused

    def _hits_fused(self, users_t, kmax):
        """One kernel per batch: ``invpref_eval_topk``.  Returns (hits [b, kmax] float64, n_gt [b] float64, top)."""
        import ctypes as C
        from . import _lib
        hp = self.model.hot_path()
        hp.flush()
        dev = hp.device
        users_t = users_t.to(dev).contiguous()
        hp.check_ids(users_t, None, None)
        b = users_t.numel()
        m_off, m_items = self._csr_on("mask", dev)
        g_off, g_items = self._csr_on("gt", dev)
        p_off = p_items = None
        if self.use_item_pool:
            p_off, p_items = self._csr_on("pool", dev)
        top = torch.empty((b, kmax), dtype=torch.int64, device=dev)
        hits = torch.zeros((b, kmax), dtype=torch.uint8, device=dev)
        n_gt = torch.zeros(b, dtype=torch.int64, device=dev)
        pad = lambda t: t if t.numel() else torch.zeros(1, dtype=torch.int64, device=dev)   # empty lists: valid pointer
        p = _lib.make_params(hp.params)
        _lib.check(hp.lib.invpref_eval_topk(
            C.byref(hp.desc), C.byref(p), _lib.ptr(users_t, torch.int64), b, _lib.ptr(m_off, torch.int64),
            _lib.ptr(pad(m_items), torch.int64), _lib.ptr(p_off, torch.int64) if p_off is not None else None,
            _lib.ptr(pad(p_items), torch.int64) if p_items is not None else None, _lib.ptr(g_off, torch.int64),
            _lib.ptr(pad(g_items), torch.int64), int(kmax), _lib.ptr(top), None, _lib.ptr(hits), _lib.ptr(n_gt),
            _lib.stream_ptr()), "eval_topk")
        return hits.double(), n_gt.double(), top

    def _dense(self, off, items, users: np.ndarray, n_items: int, device) -> torch.Tensor:
        lens = off[users + 1] - off[users]
        rows = np.repeat(np.arange(len(users)), lens)
        cols = np.concatenate([items[off[u]:off[u + 1]] for u in users]) if len(users) else np.zeros(0, np.int64)
        m = torch.zeros((len(users), n_items), dtype=torch.bool, device=device)
        if len(rows):
            m[torch.from_numpy(rows).to(device), torch.from_numpy(cols).to(device)] = True
        return m

    def _csr_on(self, name: str, device):
        """(off, items) of the loader's CSR `name` as int64 tensors on `device`, uploaded once."""
        cache = self.__dict__.setdefault("_csr_cache", {})
        key = (name, str(device))
        if key not in cache:
            off = np.ascontiguousarray(getattr(self.data_loader, name + "_off"), dtype=np.int64)
            items = np.ascontiguousarray(getattr(self.data_loader, name + "_items"), dtype=np.int64)
            cache[key] = (torch.from_numpy(off).to(device), torch.from_numpy(items).to(device))
        return cache[key]

    def _hits_device(self, rating, users_t, kmax):
        """CUDA path: masks, top-k and hit look-ups without leaving the device."""
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        b, n_items = rating.shape
        users_t = users_t.contiguous()

        def mask(name, value, add):
            off, items = self._csr_on(name, rating.device)
            _lib.check(lib.invpref_mask_scores(_lib.ptr(rating, torch.float32), b, n_items,
                                               _lib.ptr(users_t, torch.int64), _lib.ptr(off, torch.int64),
                                               _lib.ptr(items, torch.int64) if items.numel() else None,
                                               C.c_float(value), add, _lib.stream_ptr()), "mask_scores")

        mask("mask", -float(1 << 10), 0)                                                        # evaluate.py:98
        if self.use_item_pool:
            mask("pool", float(1 << 10), 1)                                                     # evaluate.py:110
        _, top = torch.topk(rating, k=kmax)
        top = top.contiguous()
        off, items = self._csr_on("gt", rating.device)
        hits = torch.empty((b, kmax), dtype=torch.uint8, device=rating.device)
        n_gt = torch.empty(b, dtype=torch.int64, device=rating.device)
        if items.numel():
            _lib.check(lib.invpref_hits_from_csr(_lib.ptr(top, torch.int64), b, kmax, _lib.ptr(users_t, torch.int64),
                                                 _lib.ptr(off, torch.int64), _lib.ptr(items, torch.int64),
                                                 _lib.ptr(hits), _lib.ptr(n_gt), _lib.stream_ptr()), "hits_from_csr")
        else:
            hits.zero_()
            n_gt.zero_()
        return hits.double(), n_gt.double()

    def _hits_host(self, rating, users: np.ndarray, kmax):
        """Any device: dense boolean masks built per batch from the loader's CSR arrays."""
        dl = self.data_loader
        dev, n_items = rating.device, rating.shape[1]
        rating[self._dense(dl.mask_off, dl.mask_items, users, n_items, dev)] = -(1 << 10)      # evaluate.py:98
        if self.use_item_pool:
            rating += self._dense(dl.pool_off, dl.pool_items, users, n_items, dev).float() * (1 << 10)
        _, top = torch.topk(rating, k=kmax)
        gt = self._dense(dl.gt_off, dl.gt_items, users, n_items, dev)
        return torch.gather(gt, 1, top).double(), gt.sum(dim=1).double()

    def evaluate_batch(self, batch_users_tensor: torch.Tensor, batch_users_list: list, batch_users_ground_truth=None,
                       force_host: bool = False):
        users = np.asarray(batch_users_list, dtype=np.int64)
        kmax = max(self.top_k_list)
        use_fused = self.fused and not force_host and hasattr(self.model, "hot_path") and kmax <= 256 \
            and next(self.model.parameters()).is_cuda
        if use_fused:
            hits, n_gt, _ = self._hits_fused(batch_users_tensor, kmax)
            dev = hits.device
        else:
            with torch.no_grad():
                rating = self.model.predict(batch_users_tensor).clone()
            dev = rating.device
            if rating.is_cuda and not force_host:
                hits, n_gt = self._hits_device(rating.contiguous(), batch_users_tensor.to(dev), kmax)
            else:
                hits, n_gt = self._hits_host(rating, users, kmax)                               # [b, kmax] 0/1
        disc = 1.0 / torch.log2(torch.arange(2, kmax + 2, device=dev, dtype=torch.float64))
        pre, rec, ndcg = [], [], []
        for k in self.top_k_list:
            right = hits[:, :k].sum(dim=1)
            pre.append(float((right / k).sum()))
            rec.append(float((right / n_gt).sum()))
            ideal = (torch.arange(k, device=dev)[None, :] < n_gt[:, None]).double()
            idcg = (ideal * disc[:k]).sum(dim=1)
            idcg = torch.where(idcg == 0, torch.ones_like(idcg), idcg)
            ndcg.append(float(((hits[:, :k] * disc[:k]).sum(dim=1) / idcg).sum()))
        return {'ndcg': ndcg, 'recall': rec, 'precision': pre}

    def evaluate(self) -> dict:
        self.model.eval()
        users_t = self.data_loader.all_test_users_by_sorted_tensor
        users_l = self.data_loader.all_test_users_by_sorted_list
        parts = [self.evaluate_batch(ut, ul) for ut, ul in mini_batch(self.batch_size, users_t, users_l)]
        n = float(len(users_l))
        out = {}
        for metric in ('ndcg', 'recall', 'precision'):
            tot = np.sum(np.array([p[metric] for p in parts]), axis=0) / n
            out[metric] = {k: float(v) for k, v in zip(self.top_k_list, tot.tolist())}
        return out
