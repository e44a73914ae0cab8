"""Worker of tests/test_gpu_multi.py: one of WORLD_SIZE processes, one GPU each, NCCL.  Trains a small config through
the distributed trainer (dist_train.py) with the requested exchange and compares, on rank 0, with the single-GPU
trainer (train.py) on the same data, seeds and initial tables."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def synth(U, I, N, K, D, implicit, seed):
    """Interactions and tables at the scale of tests/test_gpu_parallel.py (gradients well above Adam's eps, so that a
    different summation order of the item partials moves the tables by ~1e-5, not by +-lr)."""
    rng = np.random.default_rng(seed)
    u = np.floor(U * rng.random(N) ** 1.5).astype(np.int64)
    i = np.floor(I * rng.random(N) ** 3).astype(np.int64)
    u[0], i[0] = U - 1, I - 1
    y = (rng.integers(0, 2, N) if implicit else rng.integers(1, 6, N)).astype(np.int64)
    p = {"Uinv": rng.normal(0, 0.1, (U, D)), "Iinv": rng.normal(0, 0.1, (I, D)), "Uenv": rng.normal(0, 0.3, (U, D)),
         "Ienv": rng.normal(0, 0.3, (I, D)), "E": rng.normal(0, 0.5, (K, D)), "W": rng.normal(0, 0.3, (K, D)),
         "b": rng.normal(0, 0.1, (K,))}
    return np.stack([u, i, y], axis=1), {k: v.astype(np.float32) for k, v in p.items()}


def main():
    from invpref_kdd_2022_b200.dist_train import ShardedExplicitTrainManager, ShardedImplicitTrainManager
    from invpref_kdd_2022_b200.models import InvPrefExplicit, InvPrefImplicit
    from invpref_kdd_2022_b200.train import ExplicitTrainManager, ImplicitTrainManager
    from oracle import invpref_numpy as on
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    case, exchange = sys.argv[1], sys.argv[2]
    implicit = case == "implicit"
    U, I, K, D, B, N = (3000, 407, 4, 64, 20000, 47777) if not implicit else (2500, 1203, 6, 40, 16384, 40000)
    roe, ree = (True, False) if not implicit else (False, True)
    data, p = synth(U, I, N, K, D, implicit, 5)
    M = InvPrefImplicit if implicit else InvPrefExplicit
    model = M(U, I, K, D, roe, ree).to(dev)
    with torch.no_grad():
        for k, prm in model.named_hot_params().items():
            prm.copy_(torch.tensor(p[k], device=dev))
    init = {k: prm.data.clone() for k, prm in model.named_hot_params().items()}
    common = dict(batch_size=B, epochs=3, cluster_interval=100, evaluate_interval=100, lr=1e-2, invariant_coe=0.8,
                  env_aware_coe=1.7, env_coe=1.1, L2_coe=0.6, L1_coe=0.03, alpha=None, use_class_re_weight=True,
                  use_recommend_re_weight=True)
    np.random.seed(7)
    T = ShardedImplicitTrainManager if implicit else ShardedExplicitTrainManager
    tm = T(U, I, K, D, torch.LongTensor(data), dev, reg_only_embed=roe, reg_env_embed=ree, init=init,
           exchange=exchange, **common)
    (losses, _), _, _ = tm.train(silent=True, auto=True)          # 3 epochs x 3 global batches, no re-assignment yet
    sd = tm.gather_state_dict(0)
    np.random.seed(11)
    diff = tm.cluster()                                           # then ONE re-assignment on the trained tables
    cnts = tm.stat_envs()
    envs = [None] * world if rank == 0 else None
    dist.gather_object((tm.rows.cpu().numpy(), tm.envs.cpu().numpy()), envs, dst=0)
    out = {"rank": rank, "exchange": tm.exchange, "ok": True,
           "graph_epochs": bool(tm._graph is not None and len(tm._graph.handles) > 0),
           "peer_sync": tm.trainer.sync is not None}
    if rank == 0:
        class Null:
            def evaluate(self):
                return {"mse": 0.0}
        np.random.seed(7)
        T1 = ImplicitTrainManager if implicit else ExplicitTrainManager
        ref = T1(model=model, evaluator=Null(), device=dev, training_data=torch.LongTensor(data).to(dev), **common)
        (rlosses, _), _, _ = ref.train(silent=True, auto=True)
        rsd = {k: v.clone() for k, v in model.state_dict().items()}
        np.random.seed(11)
        rdiff = ref.cluster()
        rcnts = ref.stat_envs()
        lerr = max(abs(a[k] - b[k]) / abs(b[k]) for a, b in zip(losses, rlosses) for k in on.LOSS_KEYS)
        terr = {k: float((sd[k] - rsd[k]).abs().max() / rsd[k].abs().max()) for k in rsd}
        full = np.zeros(N, dtype=np.int64)
        for rows, e in envs:
            full[rows] = e
        mism = int((full != ref.envs.cpu().numpy()).sum())
        out.update({"max_rel_loss_err": lerr, "max_table_err": max(terr.values()), "table_err": terr,
                    "env_mismatch": mism, "N": N, "diff": diff, "ref_diff": rdiff, "counts_sum": sum(cnts.values()),
                    "counts": cnts, "ref_counts": rcnts,
                    "ok": bool(lerr <= 1e-4 and max(terr.values()) <= 5e-4 and sum(cnts.values()) == N)})
    print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
