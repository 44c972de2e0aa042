"""Where does the time go at N>1?  Times, per rank: the local step, the union kernel, the MAX
all-reduce, the cardinality of the result.  torchrun --nproc-per-node N tools/dist_probe.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dandd_b200 import dist as dd_dist
from dandd_b200.engine import Engine
rank, world = dd_dist.init("nccl")
local = int(os.environ.get("LOCAL_RANK", "0"))
eng = Engine(local)
m, nk = 1 << 20, 23
regs = torch.randint(0, 30, (12, nk, m), dtype=torch.uint8, device=eng.device)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n
full = eng.union([regs[g] for g in range(12)])
print(rank, "union", t(lambda: eng.union([regs[g] for g in range(12)])))
print(rank, "allreduce", t(lambda: dist.all_reduce(full, op=dist.ReduceOp.MAX)))
print(rank, "cards", t(lambda: eng.cards(full, 20)))
orders = np.stack([np.random.default_rng(i).permutation(12) for i in range(30)]).astype(np.int32)
print(rank, "prefix", t(lambda: eng.prefix_union_cards(regs, orders, 20), 5))
dist.barrier(); dist.destroy_process_group()
