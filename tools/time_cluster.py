"""CUDA-event timing of the re-assignment kernels (invpref_cluster / env_hist / stat_envs) on the dataset-scale
BASELINE shapes, one batch per launch as the trainer issues them.  GPU only.   python tools/time_cluster.py"""
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B  # noqa: E402
from invpref_kdd_2022_b200 import engine  # noqa: E402


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3      # us


def main():
    dev = torch.device("cuda:0")
    out = {}
    for name in sys.argv[1:] or ["c2", "c3", "c4"]:
        w = B.WORKLOADS[name]
        U, I, Bn, batches = B.synth_batches(w, 1)
        u, i, y, e = (torch.from_numpy(a).to(dev) for a in batches[0])
        tabs = B.make_tables(w, dev)
        hot = engine.HotPath(tabs, w["implicit"], w["roe"], w["ree"], lr=w["lr"])
        K = w["K"]
        base = torch.Tensor([1e-10 * (1e-1 ** k) for k in range(K)])
        eps = torch.Tensor(list(itertools.permutations(base))).to(dev)
        pidx = torch.randint(0, eps.shape[0], (Bn,), device=dev)
        new = torch.empty(Bn, dtype=torch.int64, device=dev)
        sw = torch.empty(Bn, dtype=torch.float32, device=dev)
        hist = hot.env_hist(e)
        r = {"B": Bn,
             "cluster_us": timed(lambda: hot.cluster(u, i, y, pidx, eps, e, trusted=True, out=new)),
             "cluster_no_tiebreak_us": timed(lambda: hot.cluster(u, i, y, None, None, e, trusted=True, out=new)),
             "env_hist_us": timed(lambda: hot.env_hist(e)),
             "stat_envs_us": timed(lambda: hot.stat_envs(e, hist, out=sw))}
        r["cluster_Gsamples_s"] = Bn / r["cluster_us"] / 1e3
        out[name] = {k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items()}
    print(json.dumps({"reassignment_kernels": out}))


if __name__ == "__main__":
    main()
