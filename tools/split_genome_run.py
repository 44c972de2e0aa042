#!/usr/bin/env python
"""One genome over N GPUs ("genomes < GPUs"): every rank sketches its part of the same
FASTA (dandd_b200.dist.split_fasta: whole records, overlapping pieces of large records), registers are
max-reduced over NCCL.  Rank 0 also sketches the whole file alone and checks that the merged
registers are bit-identical.   torchrun --nproc-per-node N tools/split_genome_run.py [bases] [records]"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.scale_check import synth_fasta  # noqa: E402


def main():
    bases = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
    records = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    from dandd_b200 import dist as dd_dist
    rank, world = dd_dist.init("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from dandd_b200.engine import Engine
    eng = Engine(local)
    ks, p = list(range(10, 33)), 20
    text = synth_fasta(bases, records, seed=7, device=eng.device).cpu().numpy()     # same text on every rank
    t0 = time.perf_counter()
    part = dd_dist.split_fasta(text, world, only=rank)[rank]
    t_split = time.perf_counter() - t0
    dev_part = torch.from_numpy(np.frombuffer(part, dtype=np.uint8).copy()).to(eng.device)
    for _ in range(2):      # second pass is the timed one (NCCL channels, lazy kernel loading)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        regs, _ = eng.sketch(eng.pack(dev_part, start=0), ks, p=p, floor_every=64_000_000)
        dd_dist.union_over_ranks(regs)
        cards = eng.cards(regs, p)
        torch.cuda.synchronize()
        t_par = time.perf_counter() - t0
    tt = torch.tensor([t_par], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    rep = {"bases": bases, "records": records, "n_gpus": world, "nk": len(ks), "p": p, "part_bytes": len(part),
           "split_host_s": t_split, "sketch_merge_s": float(tt.item())}
    if rank == 0:
        whole = torch.from_numpy(text).to(eng.device)
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ref, ref_cards = eng.sketch(eng.pack(whole, start=0), ks, p=p, floor_every=64_000_000)
            torch.cuda.synchronize()
            t_one = time.perf_counter() - t0
        rep.update(one_gpu_s=t_one, identical=bool(torch.equal(ref, regs)),
                   cards_equal=bool(torch.equal(ref_cards, cards)), speedup=t_one / float(tt.item()))
        print(json.dumps(rep))
    if world > 1:
        dist.barrier()


if __name__ == "__main__":
    main()
