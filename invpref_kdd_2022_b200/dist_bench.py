"""bench.py's multi-GPU leg: one process per GPU under torchrun (NCCL), see parallel.py.

Fixed global workload (strong scaling): the same global batches as the 1-GPU run are split over the ranks.
Timed region: K steps between a barrier + synchronize on both sides, CUDA events on every rank, MAX over
ranks; value = K * global_batch / that time.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch
import torch.distributed as dist


def run(args, w):
    from invpref_kdd_2022_b200 import _lib
    from invpref_kdd_2022_b200.parallel import DistDriver, ReplicatedTrainer, ShardedTrainer
    import bench as B

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    # stdout carries exactly one JSON line: NCCL's version banner / debug log (NCCL_DEBUG from the environment or
    # nccl.conf) goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    drv = DistDriver()
    K, D = w["K"], w["D"]
    nb = max(1, min(args.nbatch, args.steps + args.warmup))
    U, I, Bg, batches = B.synth_batches(w, nb)
    P = 2 * (U + I) * D + 2 * K * D + K
    sharded = P * 4 > 2.5e8                       # dataset-scale tables are replicated
    kw = dict(alpha=1.0, use_class_rw=w["crw"], use_rec_rw=w["rrw"], **w["coef"])
    t = lambda a: torch.from_numpy(a).to(dev)

    p2p_note = "off"
    if sharded:
        cache_rows = min(I, Bg // world * 2 + 1024)
        # item exchange over peer memory (NVLink loads from torch symmetric memory) unless unavailable or
        # INVPREF_P2P=0; every rank must take the same path, so the outcome is agreed on with an all-reduce
        store, why = None, "disabled (INVPREF_P2P=0)"
        if os.environ.get("INVPREF_P2P", "1") != "0":
            try:
                from invpref_kdd_2022_b200.parallel import SymmetricItemStorage
                store = SymmetricItemStorage(I, D, world, cache_rows, dev, dist.group.WORLD)
                why = ""
            except Exception as ex:      # noqa: BLE001 -- any failure means "use NCCL"
                store, why = None, f"{type(ex).__name__}: {ex}"[:200]
        ok = torch.tensor([1 if store is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            store = None
        tr = ShardedTrainer(U, I, K, D, w["implicit"], w["roe"], w["ree"], w["lr"], rank, world, dev,
                            cache_rows=cache_rows, alloc=store.alloc if store is not None else None)
        if store is not None:
            tr.enable_p2p(store.ptrs("Iinv"), store.ptrs("Ienv"), store.ptrs("gcache0"), store.ptrs("gcache1"))
            p2p_note = "peer memory (torch symmetric memory, NVLink loads)"
        else:
            p2p_note = "NCCL all-to-all (" + (why or "a peer could not map symmetric memory") + ")"
        torch.cuda.synchronize()
        dist.barrier()                   # every shard initialised before anyone reads a peer's rows
        prepared = []
        for (u, i, y, e) in batches:
            sb = drv.run(tr.prepare_gen(t(u), t(i), t(y)))
            le = t(e)[sb.sel].contiguous()
            cw, sw = tr.hot.stat_envs(le, tr.hot.env_hist(le)) if le.numel() else (None, torch.zeros(0, device=dev))
            prepared.append((sb, le, sw))

        def step(s):
            sb, le, sw = prepared[s % nb]
            return drv.run(tr.step_gen(sb, le, sw, next_sb=prepared[(s + 1) % nb][0], **kw))
        mode = f"users+items row-sharded (mod {world}), interactions routed to the user's owner, " \
               f"item rows/grads exchanged over {p2p_note}, E/W/b all-reduce"
    else:
        g = torch.Generator(device=dev).manual_seed(17373331)
        tr = ReplicatedTrainer(B.make_tables(w, dev), w["implicit"], w["roe"], w["ree"], w["lr"], rank, world)
        prepared = []
        for (u, i, y, e) in batches:
            a, b = tr.chunk(0, Bg)
            le = t(e[a:b])
            cw, sw = tr.hot.stat_envs(le, tr.hot.env_hist(le)) if le.numel() else (None, torch.zeros(0, device=dev))
            lu, li = t(u[a:b]), t(i[a:b])
            plan = tr.hot.new_plan(lu, li) if lu.numel() else None
            prepared.append((lu, li, t(y[a:b]), le, sw, plan))

        def step(s):
            lu, li, ly, le, sw, plan = prepared[s % nb]
            return drv.run(tr.step_gen(lu, li, ly, le, sw, Bg, plan=plan, **kw))
        mode = f"tables replicated, batch chunked over {world} ranks, flat gradient all-reduce"

    for s in range(args.warmup):
        step(s)
    if sharded:
        tr.flush()
    torch.cuda.synchronize()
    if sharded:
        tr.phase_events = []
    clocks = B.ClockSampler(dev.index)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for s in range(args.warmup, args.warmup + args.steps):
        loss = step(s)
    if sharded:
        tr.flush()        # lazy Adam: every local user row brought up to date inside the timed region
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = _lib.launch_count() - l0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item()) / args.steps
    assert torch.isfinite(loss).all()
    clk = clocks.stop() if rank == 0 else None
    phases = {}
    if sharded and tr.phase_events:
        evs = tr.phase_events
        for (n0, e0), (n1, e1) in zip(evs[:-1], evs[1:]):
            if n1 != "start":
                phases[n1] = phases.get(n1, 0.0) + e0.elapsed_time(e1) / args.steps
        info = [sb.route.n_cache for sb, _, _ in prepared], [int(sb.users.numel()) for sb, _, _ in prepared]
        phases["cache_rows"], phases["local_batch"] = info[0][0], info[1][0]
    # ---- e2e at N GPUs: the per-step mutable inputs of this rank's share (scores, envs, sample weights) come
    # from pinned host memory every step, the six losses go back to the host, host-synchronised per step.
    # (ids and their routing / sort-segment plans are static per batch and stay resident, as in the trainer.)
    if sharded:
        tr.phase_events = None
        host = [(sb.scores.cpu().pin_memory(), le.cpu().pin_memory(), sw.cpu().pin_memory()) for sb, le, sw in prepared]
    else:
        host = [(p[2].cpu().pin_memory(), p[3].cpu().pin_memory(), p[4].cpu().pin_memory()) for p in prepared]
    h_loss = torch.empty(6).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step(s):
        j = s % nb
        hy, he, hw = host[j]
        if sharded:
            sb, le, sw = prepared[j]
            sb.scores.copy_(hy, non_blocking=True); le.copy_(he, non_blocking=True); sw.copy_(hw, non_blocking=True)
        else:
            prepared[j][2].copy_(hy, non_blocking=True); prepared[j][3].copy_(he, non_blocking=True)
            prepared[j][4].copy_(hw, non_blocking=True)
        out = step(s)
        h_loss.copy_(out, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step(0)
    dist.barrier()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        e2e_step(s)
    if sharded:
        tr.flush()
        torch.cuda.synchronize()
    dist.barrier()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) / e2e_steps * 1e3], device=dev)
    dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    h2d = torch.tensor([sum(t_.numel() * t_.element_size() for t_ in host[0])], device=dev, dtype=torch.float64)
    dist.all_reduce(h2d)
    peak, peak_src = B.measured_peaks()
    sbytes = B.step_bytes(Bg, D, K, P)
    if rank == 0:
        line = {"metric": "train interactions/sec (fwd+bwd+Adam)", "value": Bg / (ms * 1e-3),
                "unit": "interactions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": w["name"], "global_batch": Bg, "params": P, "distinct_batches": nb,
                           "l2": "inputs larger than L2" if sharded else "tables fit in L2 (no flush)",
                           "parallelism": mode},
                "roofline": {"bound": "hbm", "kernel": "fused train step (all kernels, all ranks)",
                             "achieved": sbytes / (ms * 1e-3) / 1e9, "peak": peak * world,
                             "peak_source": peak_src + f" x {world} GPUs", "unit": "GB/s",
                             "frac": sbytes / (ms * 1e-3) / 1e9 / (peak * world), "traffic": None},
                "e2e": {"value": Bg / (e2e_ms * 1e-3), "unit": "interactions/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": 24 * world, "steps": e2e_steps,
                        "note": "per step every rank copies its share's scores / envs / sample weights from pinned "
                                "host memory, runs the step, reads the six losses back, host-synchronised; ids, "
                                "routing and sort-segment plans are static per batch and stay resident"},
                "gpu_launches": int(launches), "clocks": clk, "final_loss": float(loss[5]),
                "rank0_phase_ms": phases}
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()
