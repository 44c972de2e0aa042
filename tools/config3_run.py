#!/usr/bin/env python
"""Config 3 (BASELINE.json): G synthetic human-scale assemblies sharded across N B200s, full k sweep
(k = 2..32), progressive unions (identity ordering) and delta / argmax-k per prefix.

    torchrun --nproc-per-node N tools/config3_run.py --bases 3.1e9 --genomes 8

Every rank generates the same ancestor on its GPU (24 chromosome-like records, 1 % N runs, 50 %
soft-masked; seed 3) and derives its own genomes from it (0.1 % substitutions, seed 3+g), then
    K1+K2  pack + all-k sketch of its genomes (chunked, min-register floor refreshed per chunk)
    C1     all_gather of the register arrays  [G][31][2^20]  (31 MiB per genome)
    K3+K4  prefix unions of the identity ordering, k range split across ranks, cards all-gathered
    host   delta_i = max_k card_i,k / k and its argmax k for every prefix i
Times are CUDA-event / wall times per phase, max over ranks.  --check recomputes everything on
rank 0 alone from the gathered registers' sources is not possible at this size; instead the
result file can be diffed between runs with different N (registers and cards must be identical)."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.synth import mutate_text, synth_fasta  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=float, default=3.1e9)
    ap.add_argument("--genomes", type=int, default=8)
    ap.add_argument("--kmin", type=int, default=2)
    ap.add_argument("--kmax", type=int, default=32)
    ap.add_argument("--p", type=int, default=20)
    ap.add_argument("--chunk", type=float, default=64e6)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    from dandd_b200 import build, dist as dd_dist
    rank, world = dd_dist.init("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from dandd_b200.engine import Engine
    eng = Engine(local)
    dev = eng.device
    ks = list(range(args.kmin, args.kmax + 1))
    nk, m, p = len(ks), 1 << args.p, args.p
    G = args.genomes
    owners = [[g for g in range(G) if g % world == r] for r in range(world)]
    mine = owners[rank]

    def gpu_wall(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, (time.perf_counter() - t0) * 1e3

    t_all0 = time.perf_counter()
    ancestor, t_gen = gpu_wall(lambda: synth_fasta(int(args.bases), 24, seed=3, device=dev))
    local_regs = torch.empty((len(mine), nk, m), dtype=torch.uint8, device=dev)
    local_cards = torch.empty((len(mine), nk), dtype=torch.float64, device=dev)
    t_pack = t_sketch = t_mut = 0.0
    for j, g in enumerate(mine):
        text, dt = gpu_wall(lambda: mutate_text(ancestor, 0.001, 3 + g) if g else ancestor)
        t_mut += dt
        seq, dt = gpu_wall(lambda: eng.pack(text, start=0))
        t_pack += dt
        (_, cards), dt = gpu_wall(lambda: eng.sketch(seq, ks, p=p, out=local_regs[j]))
        t_sketch += dt
        local_cards[j] = cards
        del text, seq
    del ancestor
    torch.cuda.empty_cache()

    allregs, t_gather = gpu_wall(lambda: dd_dist.gather_registers(local_regs, owners))
    allcards = dd_dist.gather_cards(local_cards, owners)

    # progressive unions of the identity ordering; k columns split across ranks
    cols = list(dd_dist.split_work(nk))

    def progressive():
        out = torch.zeros((G, nk), dtype=torch.float64, device=dev)
        if cols:
            sub = allregs[:, cols[0]:cols[-1] + 1].contiguous()
            c = eng.prefix_union_cards(sub, [list(range(G))], p)       # [1, G, len(cols)]
            out[:, cols[0]:cols[-1] + 1] = c[0]
        if world > 1:
            dist.all_reduce(out, op=dist.ReduceOp.SUM)
        return out
    prefix_cards, t_prog = gpu_wall(progressive)
    t_total = (time.perf_counter() - t_all0) * 1e3

    times = torch.tensor([t_gen, t_mut, t_pack, t_sketch, t_gather, t_prog, t_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    if rank == 0:
        karr = np.array(ks, dtype=np.float64)
        pc = prefix_cards.cpu().numpy()
        deltas = pc / karr
        lc = allcards.cpu().numpy()
        t = times.cpu().numpy()
        total_bases = args.bases * G
        rep = {
            "config": {"genomes": G, "bases_per_genome": args.bases, "k": [args.kmin, args.kmax], "p": p, "n_gpus": world,
                       "chunk": args.chunk},
            "ms_max_over_ranks": {"generate": t[0], "mutate": t[1], "pack": t[2], "sketch": t[3], "gather": t[4],
                                  "progressive": t[5], "total_wall": t[6]},
            "gbp_s_pack_plus_sketch": total_bases / ((t[2] + t[3]) / 1e3) / 1e9,
            "tree_wall_s_without_generation": (t[2] + t[3] + t[4] + t[5]) / 1e3,
            "leaf_delta": [float((lc[g] / karr).max()) for g in range(G)],
            "leaf_argmax_k": [int(ks[int(np.argmax(lc[g] / karr))]) for g in range(G)],
            "prefix_delta": [float(deltas[i].max()) for i in range(G)],
            "prefix_argmax_k": [int(ks[int(np.argmax(deltas[i]))]) for i in range(G)],
            "registers_sha1": hashlib.sha1(allregs.cpu().numpy().tobytes()).hexdigest(),
            "prefix_cards_sha1": hashlib.sha1(np.round(pc, 3).tobytes()).hexdigest(),
        }
        print(json.dumps(rep))
        if args.out:
            with open(args.out, "w") as fh:
                json.dump(rep, fh, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
