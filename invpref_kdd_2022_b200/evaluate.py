"""Evaluators with the reference's result formats (reference ``evaluate.py:59-212``), off the hot path
(SURVEY.md §8f ranks 1 and 4).

``ExplicitTestManager``: MSE / RMSE / MAE of ``model.predict(test_users, test_items)`` (the fused
predict kernel).  ``ImplicitTestManager``: recall / precision / NDCG @k per test user; the reference builds
python index lists per user and counts hits with python sets (evaluate.py:94-135); here masks and ground truth
are dense boolean matrices built once per batch on the device and the metrics are tensor expressions.
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

    def __init__(self, model, data_loader, test_batch_size: int, top_k_list: list, use_item_pool: bool = False):
        self.model, self.data_loader = model, data_loader
        self.batch_size = test_batch_size
        self.top_k_list = sorted(top_k_list)
        self.use_item_pool = use_item_pool

    def _dense(self, off, items, users: np.ndarray, n_items: int, device) -> torch.Tensor:
        lens = off[users + 1] - off[users]
        rows = np.repeat(np.arange(len(users)), lens)
        cols = np.concatenate([items[off[u]:off[u + 1]] for u in users]) if len(users) else np.zeros(0, np.int64)
        m = torch.zeros((len(users), n_items), dtype=torch.bool, device=device)
        if len(rows):
            m[torch.from_numpy(rows).to(device), torch.from_numpy(cols).to(device)] = True
        return m

    def evaluate_batch(self, batch_users_tensor: torch.Tensor, batch_users_list: list, batch_users_ground_truth=None):
        dl = self.data_loader
        users = np.asarray(batch_users_list, dtype=np.int64)
        with torch.no_grad():
            rating = self.model.predict(batch_users_tensor).clone()
        dev, n_items = rating.device, rating.shape[1]
        rating[self._dense(dl.mask_off, dl.mask_items, users, n_items, dev)] = -(1 << 10)      # evaluate.py:98
        if self.use_item_pool:
            rating = rating + self._dense(dl.pool_off, dl.pool_items, users, n_items, dev).float() * (1 << 10)
        kmax = max(self.top_k_list)
        _, top = torch.topk(rating, k=kmax)
        gt = self._dense(dl.gt_off, dl.gt_items, users, n_items, dev)
        hits = torch.gather(gt, 1, top).double()                                                # [b, kmax] 0/1
        n_gt = gt.sum(dim=1).double()
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
