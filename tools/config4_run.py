#!/usr/bin/env python
"""Config 4 (BASELINE.json): --exact mode, KMC-equivalent distinct canonical k-mer counts for
8 synthetic 100 Mbp genomes (mutated copies of one ancestor), k in {8,12,16,20,24,28,32}: per-genome
counts and the progressive exact unions, with the CPU oracle (sort + unique) as the checker for a
subset.  python tools/config4_run.py [--bases 100e6] [--genomes 8]
Under torchrun every rank scans all genomes but keeps only its key range of the k-mer set
(dd_exact_insert_shard); the per-rank distinct counts are summed with one all-reduce."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.config3_run import mutate_text  # noqa: E402
from tools.scale_check import synth_fasta  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=float, default=100e6)
    ap.add_argument("--genomes", type=int, default=8)
    ap.add_argument("--ks", default="8,12,16,20,24,28,32")
    args = ap.parse_args()
    from dandd_b200 import build, dist as dd_dist
    rank, world = dd_dist.init("nccl")
    if rank == 0:
        build.build()
    if world > 1:
        torch.distributed.barrier()
    from dandd_b200.engine import Engine
    from oracle import pyoracle as orc
    eng = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    shard = (rank, world) if world > 1 else None

    def counts(seq_list, k):
        got = torch.tensor(eng.exact_counts(seq_list, k, shard=shard), dtype=torch.int64, device=eng.device)
        return [int(v) for v in dd_dist.sum_counts(got).cpu().tolist()]
    ks = [int(k) for k in args.ks.split(",")]
    anc = synth_fasta(int(args.bases), 8, seed=4, device=eng.device)
    texts = [anc if g == 0 else mutate_text(anc, 0.01, 4 + g) for g in range(args.genomes)]
    seqs = [eng.pack(t, start=0) for t in texts]
    total = sum(s.nsym for s in seqs)
    rep = {"genomes": args.genomes, "bases_per_genome": args.bases, "ks": ks, "n_gpus": world, "rows": []}
    for k in ks:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        singles = [counts([s], k)[0] for s in seqs]                      # `kmc` + `kmc_tools info` per genome
        prefix = counts(seqs, k)                                         # `kmc_tools complex` unions, progressive
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rep["rows"].append({"k": k, "single": singles, "progressive_union": prefix, "seconds": dt,
                            "gkmer_per_s": 2 * total / dt / 1e9})
        if rank == 0:
            print(json.dumps(rep["rows"][-1]), flush=True)
    if rank != 0:
        torch.distributed.barrier()
        return
    # oracle check on what the CPU can do in seconds: genome 0 at k = 20 and 32, union of two at k = 12
    sym0 = orc.fasta_symbols(texts[0].cpu().numpy().tobytes())
    sym1 = orc.fasta_symbols(texts[1].cpu().numpy().tobytes())
    checks = {}
    for k in (20, 32):
        if k in ks:
            t0 = time.perf_counter()
            want = orc.exact_count([sym0], k)
            checks[f"single_k{k}"] = {"equal": want == rep["rows"][ks.index(k)]["single"][0], "oracle_s": time.perf_counter() - t0}
    if 12 in ks:
        want = orc.exact_count([sym0, sym1], 12)
        checks["union2_k12"] = {"equal": want == rep["rows"][ks.index(12)]["progressive_union"][1]}
    rep["oracle_checks"] = checks
    print(json.dumps({"oracle_checks": checks}))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "config4.json"), "w") as fh:
        json.dump(rep, fh, indent=1)
    if world > 1:
        torch.distributed.barrier()


if __name__ == "__main__":
    main()
