#!/usr/bin/env python
"""Config 5 (BASELINE.json): helpers/allpairs.py-style pairwise union / delta / KIJ matrix over N
synthetic 5 Mbp genomes (clusters of 10 mutated copies), k = 10..32, p = 18 (allpairs default
--nest 262144; reference helpers/allpairs.py:327,360-380 re-sketches BOTH FASTAs for every pair and k).

    [torchrun --nproc-per-node G] python tools/config5_run.py --genomes 1000 --out profiles/r02_config5_nG.json

    sketch   genomes are sharded round-robin over the ranks; each rank packs + sketches only its own
    gather   NCCL all_gather of the register arrays: every rank then holds all N x 23 x 2^18 sketches
    planes   dd_to_planes ONCE per rank (bit planes of all sketches stay in HBM)
    pairs    the N(N-1)/2 pair list is split across ranks; each rank runs its share through K6 in tiles
             (dd_pairwise_union_card_planes), delta_AB = max_k card/k and KIJ on the device
    check    a few pairs x all k against the oracle (registers from the oracle, numpy max, oracle MLE)
Tile reuse factor = sketch bytes a tile's pairs touch / distinct sketch bytes in the tile (how often a
sketch loaded into L2 is used again inside the tile)."""
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
from tools.synth import mutate_text, synth_fasta  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", dest="n", type=int, default=1000)
    ap.add_argument("--bases", type=float, default=5e6)
    ap.add_argument("--p", type=int, default=18)
    ap.add_argument("--tile", type=int, default=1 << 16)
    ap.add_argument("--check", type=int, default=4)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from dandd_b200 import build, dist as dd_dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, world = dd_dist.init("nccl") if world > 1 else (0, 1)
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    from dandd_b200.engine import Engine
    eng = Engine(local)
    dev = eng.device
    ks = list(range(10, 33))
    nk, p, n = len(ks), args.p, args.n
    m = 1 << p

    def sync_time(fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, time.perf_counter() - t0

    owners = [[g for g in range(n) if g % world == r] for r in range(world)]
    mine = owners[rank]
    keep_text = {}

    def genome_text(g, cache={}):
        c = g // 10
        if cache.get("c") != c:
            cache["c"], cache["anc"] = c, synth_fasta(int(args.bases), 1, seed=5000 + c, device=dev)
        return cache["anc"] if g % 10 == 0 else mutate_text(cache["anc"], 0.02, 50000 + g)

    texts = []                      # this rank's FASTA texts, resident in HBM (generation is not part of the timed path)
    for g in mine:
        text = genome_text(g)
        texts.append(text)
        if g < args.check:
            keep_text[g] = text.cpu().numpy().tobytes()

    def sketch_mine():
        regs = torch.empty((len(mine), nk, m), dtype=torch.uint8, device=dev)
        hist = torch.empty((len(mine), nk, 64), dtype=torch.int32, device=dev)
        for j, text in enumerate(texts):
            eng.sketch(eng.pack(text, start=0), ks, p=p, out=regs[j], hist_out=hist[j])
        return regs, eng.mle(hist, p)
    sketch_mine()                   # warm-up (allocations, lazy module loading)
    (local_regs, local_cards), t_sketch = sync_time(sketch_mine)
    del texts
    if world > 1:                   # NCCL builds its channels on the first collective: keep that out of the timing
        dd_dist.union_over_ranks(torch.zeros(1 << 20, dtype=torch.uint8, device=dev))
    (regs, single), t_gather = sync_time(lambda: (dd_dist.gather_registers(local_regs, owners), dd_dist.gather_cards(local_cards, owners)))
    del local_regs
    planes, t_planes = sync_time(lambda: eng.to_planes(regs, p))
    iu = np.triu_indices(n, 1)
    pairs = np.stack(iu, axis=1).astype(np.int32)
    span = list(dd_dist.split_work(len(pairs))) if world > 1 else range(len(pairs))
    my_pairs = pairs[span[0]:span[-1] + 1] if len(span) else pairs[:0]
    karr = torch.tensor(ks, device=dev, dtype=torch.float64)
    d_single = (single / karr).max(dim=1).values

    reuse = []

    def all_pairs():
        kij_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        first = None
        for s in range(0, len(my_pairs), args.tile):
            pr = my_pairs[s:s + args.tile]
            cards = eng.pairwise_cards(None, pr, p, planes=planes, n_genomes=n, nk=nk)      # [tile, nk]
            d_pair = (cards / karr).max(dim=1).values
            a = torch.as_tensor(pr[:, 0], device=dev).long()
            b = torch.as_tensor(pr[:, 1], device=dev).long()
            kij_sum += ((d_single[a] + d_single[b] - d_pair) / d_pair).sum()
            reuse.append(2 * len(pr) / len(np.unique(pr)))
            if first is None:
                first = cards[:64].cpu().numpy()
        return kij_sum, first
    eng.pairwise_cards(None, pairs[:256], p, planes=planes, n_genomes=n, nk=nk)     # warm-up: lazy module loading of the pair kernel
    reuse.clear()
    (kij_sum, first), t_pairs = sync_time(all_pairs)
    tt = torch.tensor([t_sketch, t_gather, t_planes, t_pairs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(kij_sum, op=dist.ReduceOp.SUM)
    if rank == 0:
        from oracle import pyoracle as orc
        sym = {g: orc.fasta_symbols(t) for g, t in keep_text.items()}
        oreg = {g: [orc.hll_sketch(sym[g], k, p) for k in ks] for g in sym}
        worst, checked = 0.0, 0
        for j, (a, b) in enumerate(pairs[:64]):
            if a in oreg and b in oreg:
                for i in range(nk):
                    want = orc.card(np.maximum(oreg[a][i], oreg[b][i]), p)
                    worst = max(worst, abs(want - float(first[j, i])) / want)
                    checked += 1
        regs_ok = all(np.array_equal(regs[g].cpu().numpy(), np.stack(oreg[g])) for g in oreg if g % world == 0 or True)
        cells = len(pairs) * nk
        total = float(tt.sum())
        rep = {"n": n, "p": p, "nk": nk, "n_gpus": world, "pairs": len(pairs), "genome_bp": args.bases,
               "sketch_s": float(tt[0]), "gather_s": float(tt[1]), "to_planes_s": float(tt[2]), "pairs_s": float(tt[3]),
               "total_s": total, "pair_k_cells_per_s": cells / float(tt[3]), "sketch_gbp_per_s": n * args.bases / float(tt[0]) / 1e9,
               "tile_pairs": args.tile, "tile_reuse_factor_mean": float(np.mean(reuse)) if reuse else None,
               "k6_algorithmic_GB_per_s": cells * 2 * m / float(tt[3]) / 1e9,
               "mean_kij": float(kij_sum) / len(pairs), "oracle_cells_checked": checked, "oracle_max_rel_err": worst,
               "oracle_registers_equal": bool(regs_ok),
               "reference_shape": "helpers/allpairs.py would run 2 sketches + 1 union card for each of the "
                                  f"{cells} (pair, k) cells: {2 * cells} FASTA passes"}
        print(json.dumps(rep))
        if args.out:
            with open(args.out, "w") as fh:
                json.dump(rep, fh, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
