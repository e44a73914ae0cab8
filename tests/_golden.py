"""Loader for the fixtures written by tests/golden/make_golden.py."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ("coat_explicit", "explicit_sched_k5", "implicit_k2", "implicit_k6", "explicit_d64_k4")


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(HERE, "golden", name + ".npz"))
        meta = dict(zip(self.z["meta_keys"].tolist(), self.z["meta_vals"].tolist()))
        self.implicit = meta["implicit"] == "True"
        self.K, self.D, self.B = int(meta["K"]), int(meta["D"]), int(meta["B"])
        self.roe, self.ree = meta["roe"] == "True", meta["ree"] == "True"
        self.crw, self.rrw = meta["crw"] == "True", meta["rrw"] == "True"
        self.lr = float(meta["lr"])
        self.alpha = None if meta["alpha"] == "None" else float(meta["alpha"])
        self.coef = {k: float(meta[k]) for k in ("c_inv", "c_ea", "c_env", "c_L2", "c_L1")}
        self.seed = int(meta["seed"])
        self.data = self.z["data"].astype(np.int64)
        self.N = len(self.data)
        self.U = int(self.data[:, 0].max()) + 1
        self.I = int(self.data[:, 1].max()) + 1

    def group(self, prefix):
        """state-dict style sub-dictionary, e.g. group('init') -> {'embed_env.weight': ...}."""
        pre = prefix + "/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def __getitem__(self, k):
        return self.z[k]
