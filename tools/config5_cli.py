#!/usr/bin/env python
"""BASELINE config 5 through the drop-in all-pairs driver, from FASTA FILES: N synthetic genomes
(clusters of 10, 2 % substitutions inside a cluster) on disk + an AFproject dataset file, then
`dandd_b200/helpers/allpairs.py --tool dashing --klist 10..32 --nest 2^p` in this process (so that the
stage times and the table are at hand), i.e. what a user of the reference's helpers/allpairs.py runs.
tools/config5_run.py times the same job on texts that are already in HBM; this one adds reading the
files and writing card.tsv / delta.tsv / the 2 x 24 PHYLIP matrices.

    python tools/config5_cli.py --genomes 1000 --out profiles/config5_cli_n1.json
    torchrun --nproc-per-node 8 tools/config5_cli.py --genomes 1000 --out profiles/config5_cli_n8.json
    python tools/config5_cli.py --genomes 6 --bases 2e4 --p 12 --oracle-store      # smoke test without a GPU

--check C compares C random (pair, k) cells and C single cardinalities with the oracle
(registers from the oracle's sketch of the file, numpy max, oracle MLE)."""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def write_genomes(directory, n, bases, seed=5):
    """<directory>/g0000.fasta ..: 80-column single-record FASTA, genome g a 2 % mutated copy of the ancestor
    of cluster g // 10."""
    os.makedirs(directory, exist_ok=True)
    rng = np.random.default_rng(seed)
    paths = []
    for g in range(n):
        if g % 10 == 0:
            anc = ACGT[rng.integers(0, 4, int(bases))]
        s = anc.copy()
        hit = rng.random(s.size) < 0.02
        s[hit] = ACGT[rng.integers(0, 4, int(hit.sum()))]
        body = s[:s.size // 80 * 80].reshape(-1, 80)
        path = os.path.join(directory, "g%04d.fasta" % g)
        with open(path, "wb") as fh:
            fh.write(b">g%d cluster%d\n" % (g, g // 10) + b"\n".join(r.tobytes() for r in body) + b"\n")
        paths.append(path)
    return paths


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=1000)
    ap.add_argument("--bases", type=float, default=5e6)
    ap.add_argument("--p", type=int, default=18, help="log2(--nest); the reference's default is 18")
    ap.add_argument("--kmin", type=int, default=10)
    ap.add_argument("--kmax", type=int, default=32)
    ap.add_argument("--workdir", default="/tmp/dandd_config5")
    ap.add_argument("--check", type=int, default=8)
    ap.add_argument("--cpu", type=int, default=-1, help="worker processes for the output files")
    ap.add_argument("--oracle-store", action="store_true", help="CPU smoke test: the oracle-backed store double, no GPU")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from dandd_b200 import dist as dd_dist, timing
    from dandd_b200.helpers import allpairs
    rank, world = dd_dist.init()
    if args.oracle_store:
        from dandd_b200 import store as ddstore
        from tests.oracle_store import OracleStore
        ddstore.set_store(OracleStore())
    data = os.path.join(args.workdir, "data")
    t0 = time.perf_counter()
    if rank == 0:
        shutil.rmtree(args.workdir, ignore_errors=True)
        inputs = write_genomes(data, args.genomes, args.bases)
        with open(os.path.join(args.workdir, "dataset.json"), "w") as fh:
            json.dump({"seqids": [f[:-len(".fasta")] for f in inputs], "treids": []}, fh)
        os.sync()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    t_generate = time.perf_counter() - t0
    out = os.path.join(args.workdir, "out")
    klist = list(range(args.kmin, args.kmax + 1))
    argv = ["--tool", "dashing", "--name", os.path.join(out, "run"), "--dataset", os.path.join(args.workdir, "dataset.json"),
            "--card-results", os.path.join(out, "card.tsv"), "--delta-results", os.path.join(out, "delta.tsv"),
            "--j-results-phylip", os.path.join(out, "sim.phylip"), "--ani-results-phylip", os.path.join(out, "ani.phylip"),
            "--nest", str(1 << args.p), "--klist", ",".join(map(str, klist)), "--cpu", str(args.cpu)]
    if rank == 0:
        os.makedirs(out, exist_ok=True)
    before = timing.snapshot()
    t0 = time.perf_counter()
    table = allpairs.go(argv)
    wall = time.perf_counter() - t0
    stages = {k: v - before.get(k, 0.0) for k, v in timing.snapshot().items() if not k.startswith("at_")}
    if rank == 0:
        from oracle import pyoracle as orc
        rng = np.random.default_rng(1)
        n = args.genomes
        worst, checked, sym = 0.0, 0, {}

        def registers(g, k):
            if g not in sym:
                with open(os.path.join(data, "g%04d.fasta" % g), "rb") as fh:
                    sym[g] = orc.fasta_symbols(fh.read())
            return orc.hll_sketch(sym[g], k, args.p)
        for _ in range(args.check if n > 1 else 0):
            a, b = sorted(rng.choice(min(n, 20), 2, replace=False).tolist())      # (a few genomes: each costs an oracle pass per k)
            c = int(rng.integers(0, len(klist)))
            want = orc.card(np.maximum(registers(a, klist[c]), registers(b, klist[c])), args.p)
            worst = max(worst, abs(table.pair[table.pair_row(a, b), c] - want) / want)
            want = orc.card(registers(a, klist[c]), args.p)
            worst = max(worst, abs(table.single[a, c] - want) / want)
            checked += 2
        kij = table.kij_values()[0]
        same = table.pairs[:, 0] // 10 == table.pairs[:, 1] // 10
        files = sorted(os.listdir(out))
        rep = {"genomes": n, "genome_bp": args.bases, "p": args.p, "nk": len(klist), "n_gpus": world, "pairs": int(len(table.pairs)),
               "store": "oracle double (CPU smoke test)" if args.oracle_store else "GpuSketchStore",
               "generate_s": t_generate, "allpairs_wall_s": wall, "stages_rank0": stages,
               "output_files": len(files) - 1, "output_bytes": sum(os.path.getsize(os.path.join(out, f)) for f in files if f != "run"),
               "kij_within_clusters": float(kij[same].mean()) if same.any() else None,
               "kij_between_clusters": float(kij[~same].mean()) if (~same).any() else None,
               "oracle_cells_checked": checked, "oracle_max_rel_err": worst,
               "host_cores": os.cpu_count(),
               "reference_shape": "helpers/allpairs.py would start %d `dashing hll` processes (%d FASTA passes)" % (
                   len(klist) * n * (n + 1) // 2, len(klist) * n * n)}
        print(json.dumps(rep))
        if args.out:
            with open(args.out, "w") as fh:
                json.dump(rep, fh, indent=1)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
