#!/usr/bin/env python
"""Config 5 (BASELINE.json): helpers/allpairs.py-style pairwise union / delta / KIJ / Jaccard matrix
over N synthetic 5 Mbp genomes (clusters of 10 mutated copies), k = 10..32, p = 18 (allpairs default
--nest 262144) -- sketches stay in HBM, all N(N-1)/2 pairs go through K6 in tiles.
    python tools/config5_run.py --n 1000        (torchrun: the pair list is split across ranks)"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.config3_run import mutate_text  # noqa: E402
from tools.scale_check import synth_fasta  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--bases", type=float, default=5e6)
    ap.add_argument("--p", type=int, default=18)
    ap.add_argument("--tile", type=int, default=50000)
    args = ap.parse_args()
    from dandd_b200 import build, dist as dd_dist
    rank, world = dd_dist.init("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from dandd_b200.engine import Engine
    from oracle import pyoracle as orc
    eng = Engine(local)
    ks = list(range(10, 33))
    nk, p, n = len(ks), args.p, args.n
    m = 1 << p
    regs = torch.empty((n, nk, m), dtype=torch.uint8, device=eng.device)
    hist = torch.empty((n, nk, 64), dtype=torch.int32, device=eng.device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    anc = None
    keep_text = {}
    for g in range(n):                         # every rank sketches everything here (6 GB of sketches); the
        if g % 10 == 0:                        # pair matrix is what is split across ranks
            anc = synth_fasta(int(args.bases), 1, seed=5000 + g // 10, device=eng.device)
        text = anc if g % 10 == 0 else mutate_text(anc, 0.02, 50000 + g)
        if g < 2:
            keep_text[g] = text.cpu().numpy().tobytes()
        eng.sketch(eng.pack(text, start=0), ks, p=p, out=regs[g], hist_out=hist[g])
    single = eng.mle(hist, p)
    torch.cuda.synchronize()
    t_sketch = time.perf_counter() - t0
    pairs = np.array([(a, b) for a in range(n) for b in range(a + 1, n)], dtype=np.int32)
    mine = pairs[list(dd_dist.split_work(len(pairs)))[0]:list(dd_dist.split_work(len(pairs)))[-1] + 1] if len(pairs) else pairs
    karr = torch.arange(10, 33, device=eng.device, dtype=torch.float64)
    d_single = (single / karr).max(dim=1).values
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kij_sum = torch.zeros(1, dtype=torch.float64, device=eng.device)
    first = None
    for s in range(0, len(mine), args.tile):
        pr = mine[s:s + args.tile]
        cards = eng.pairwise_cards(regs, pr, p)                       # [tile, nk]
        d_pair = (cards / karr).max(dim=1).values
        a = torch.as_tensor(pr[:, 0], device=eng.device).long()
        b = torch.as_tensor(pr[:, 1], device=eng.device).long()
        kij = (d_single[a] + d_single[b] - d_pair) / d_pair
        kij_sum += kij.sum()
        if first is None:
            first = (cards[:2].cpu().numpy(), kij[:12].cpu().numpy())
    torch.cuda.synchronize()
    t_pairs = time.perf_counter() - t0
    tt = torch.tensor([t_sketch, t_pairs], dtype=torch.float64, device=eng.device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(kij_sum, op=dist.ReduceOp.SUM)
    if rank == 0:
        sym = [orc.fasta_symbols(keep_text[g]) for g in (0, 1)]
        ok = True
        for i in (0, nk - 1):
            u = orc.union_max([orc.hll_sketch(sym[0], ks[i], p), orc.hll_sketch(sym[1], ks[i], p)])
            ok &= abs(orc.card(u, p) - float(first[0][0, i])) <= 1e-9 * orc.card(u, p)
        cells = len(pairs) * nk
        print(json.dumps({"n": n, "p": p, "nk": nk, "n_gpus": world, "pairs": len(pairs), "sketch_all_s": float(tt[0]),
                          "pairs_s": float(tt[1]), "pair_k_cells_per_s": cells / float(tt[1]),
                          "mean_kij": float(kij_sum) / len(pairs), "kij_within_cluster_first": first[1][:8].tolist(),
                          "oracle_pair01_ok": bool(ok)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
