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


def main():
    from _golden import Golden
    from invpref_kdd_2022_b200.dist_train import ShardedExplicitTrainManager, ShardedImplicitTrainManager
    from invpref_kdd_2022_b200.models import InvPrefExplicit, InvPrefImplicit
    from invpref_kdd_2022_b200.train import ExplicitTrainManager, ImplicitTrainManager
    from oracle import invpref_numpy as on
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    case, exchange = sys.argv[1], sys.argv[2]
    g = Golden(case)
    torch.manual_seed(g.seed)
    M = InvPrefImplicit if g.implicit else InvPrefExplicit
    model = M(g.U, g.I, g.K, g.D, g.roe, g.ree).to(dev)
    init = {k: p.data.clone() for k, p in model.named_hot_params().items()}
    common = dict(batch_size=g.B, epochs=3, cluster_interval=2, evaluate_interval=100, lr=g.lr,
                  invariant_coe=g.coef["c_inv"], env_aware_coe=g.coef["c_ea"], env_coe=g.coef["c_env"],
                  L2_coe=g.coef["c_L2"], L1_coe=g.coef["c_L1"], alpha=g.alpha, use_class_re_weight=g.crw,
                  use_recommend_re_weight=g.rrw)
    np.random.seed(g.seed)
    T = ShardedImplicitTrainManager if g.implicit else ShardedExplicitTrainManager
    tm = T(g.U, g.I, g.K, g.D, torch.LongTensor(g.data), dev, reg_only_embed=g.roe, reg_env_embed=g.ree, init=init,
           exchange=exchange, **common)
    (losses, _), _, (diffs, cnts, _) = tm.train(silent=True, auto=True)
    sd = tm.gather_state_dict(0)
    envs = [None] * world if rank == 0 else None
    dist.gather_object((tm.rows.cpu().numpy(), tm.envs.cpu().numpy()), envs, dst=0)
    out = {"rank": rank, "exchange": tm.exchange, "ok": True}
    if rank == 0:
        class Null:
            def evaluate(self):
                return {"mse": 0.0}
        np.random.seed(g.seed)
        T1 = ImplicitTrainManager if g.implicit else ExplicitTrainManager
        ref = T1(model=model, evaluator=Null(), device=dev, training_data=torch.LongTensor(g.data).to(dev), **common)
        (rlosses, _), _, (rdiffs, rcnts, _) = ref.train(silent=True, auto=True)
        rsd = model.state_dict()
        lerr = max(abs(a[k] - b[k]) / abs(b[k]) for a, b in zip(losses, rlosses) for k in on.LOSS_KEYS)
        terr = {k: float((sd[k] - rsd[k]).abs().max() / rsd[k].abs().max()) for k in rsd}
        full = np.zeros(g.N, dtype=np.int64)
        for rows, e in envs:
            full[rows] = e
        mism = int((full != ref.envs.cpu().numpy()).sum())
        out.update({"max_rel_loss_err": lerr, "max_table_err": max(terr.values()), "env_mismatch": mism, "N": g.N,
                    "diffs": diffs, "ref_diffs": rdiffs, "counts_sum": sum(cnts[0].values()),
                    "ok": bool(lerr <= 1e-4 and max(terr.values()) <= 1e-3 and sum(cnts[0].values()) == g.N)})
    print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
