"""Times the sort-segment plan build (invpref_build_plan: hand-written radix sort + scans + plan kernels) on the
BASELINE shapes and checks the permutation against torch.sort(stable=True).  GPU only.

    python tools/time_plan.py            -> one JSON line
"""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from invpref_kdd_2022_b200 import engine  # noqa: E402

SHAPES = {
    "c5": (10_000_000, 1_000_000, 1 << 22),
    "c4": (50_000, 51_283, 262_144),
    "c2": (15_400, 1_000, 131_072),
    "tiny": (300, 290, 9_000),
}


def main():
    dev = torch.device("cuda:0")
    out = {}
    names = sys.argv[1:] or list(SHAPES)
    for name in names:
        U, I, B = SHAPES[name]
        g = torch.Generator(device=dev).manual_seed(7)
        users = torch.randint(0, U, (B,), device=dev, generator=g)
        items = (torch.rand(B, device=dev, generator=g) ** 2 * I).long().clamp_(max=I - 1)   # skewed: long segments
        for ids, rows, side in ((users, U, "users"), (items, I, "items")):
            perm, seg_row, seg_off = engine.build_segments(ids, rows)
            ref = torch.sort(ids, stable=True)
            assert torch.equal(perm, ref.indices), (name, side, "perm")
            uq, cnt = torch.unique_consecutive(ref.values, return_counts=True)
            assert torch.equal(seg_row, uq), (name, side, "seg_row")
            assert torch.equal(seg_off[1:] - seg_off[:-1], cnt), (name, side, "seg_off")
        params = {k: torch.zeros(r, 8, device=dev) for k, r in (("Uinv", U), ("Uenv", U), ("Iinv", I), ("Ienv", I))}
        params.update(E=torch.zeros(2, 8, device=dev), W=torch.zeros(2, 8, device=dev), b=torch.zeros(2, device=dev))
        hot = engine.HotPath(params, False, False, False, lr=1e-3)
        plan = hot.new_plan(users, items)
        for _ in range(3):
            hot.new_plan(users, items, out=plan)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            hot.new_plan(users, items, out=plan)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            torch.sort(users, stable=True)
            torch.sort(items, stable=True)
        t1.record()
        torch.cuda.synchronize()
        out[name] = {"B": B, "plan_build_ms": round(ms, 4), "two_torch_sorts_int64_ms": round(t0.elapsed_time(t1) / n, 4),
                     "perm_bit_equal_to_torch_stable_sort": True}
    print(json.dumps({"plan_build": out}))


if __name__ == "__main__":
    main()
