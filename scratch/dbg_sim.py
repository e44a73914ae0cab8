import faulthandler, sys, os
faulthandler.dump_traceback_later(45, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from test_gpu_parallel import synth, KW
from invpref_kdd_2022_b200.parallel import ShardedTrainer, SimDriver
dev = torch.device("cuda:0")
def P(*a):
    torch.cuda.synchronize(); print(*a, flush=True)
world=2
U, I, K, D, B = 1000, 203, 4, 64, 20000
u, i, y, e, w, p = synth(U, I, B, K, D, False, 5)
init = {k: torch.tensor(v, device=dev) for k, v in p.items()}
ranks = [ShardedTrainer(U, I, K, D, False, True, False, 1e-2, r, world, dev, cache_rows=I, init=init) for r in range(world)]
P("built")
sim = SimDriver(world)
t = lambda a: torch.tensor(a, device=dev)
sbs = sim.run_all([rk.prepare_gen(t(u), t(i), t(y)) for rk in ranks])
P("prepared", [sb.sel.numel() for sb in sbs], [sb.route.n_cache for sb in sbs])
# manual stepping of rank generators with prints
gens = [rk.step_gen(sb, t(e)[sb.sel].contiguous(), t(w)[sb.sel].contiguous(), **KW) for rk, sb in zip(ranks, sbs)]
class Dbg(SimDriver):
    pass
reqs = [next(g) for g in gens]; P("first yield", reqs[0][0])
n=0
done=[False]*world
while True:
    op = reqs[0][0]
    if op == "all_reduce":
        tot = reqs[0][1].clone()
        for r in reqs[1:]: tot += r[1]
        for r in reqs: r[1].copy_(tot)
    else:
        for dst in range(world):
            _, out, _, out_splits, _ = reqs[dst]
            o = 0
            for src in range(world):
                _, _, inp, _, in_splits = reqs[src]
                start = sum(in_splits[:dst]); nn = in_splits[dst]
                out[o:o + nn] = inp[start:start + nn]; o += nn
    n+=1; P("collective", n, op, "done")
    for k in range(world):
        try:
            reqs[k] = gens[k].send(None); P(" rank", k, "advanced to", reqs[k][0])
        except StopIteration as s:
            done[k]=True; P(" rank", k, "finished", s.value.tolist())
    if all(done): break
P("OK")
