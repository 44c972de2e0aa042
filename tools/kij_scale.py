#!/usr/bin/env python
"""`dandd kij` at scale through the drop-in command line: N synthetic genomes (clusters of 10, 2 %
substitutions inside a cluster), `dandd tree -k 14` then `dandd kij --mink 10 --maxk 32` (the
reference builds one SubSpider -- ~5 dashing union + card processes -- per pair, lib/huffman_dandd.py:666-695;
here the N(N-1)/2 x nk union cardinalities come from ONE batched K6 job and each pair is replayed on
that table).  Reports wall times and the stage split; --check compares a few pairs with the oracle.

    python tools/kij_scale.py --genomes 200 --bases 1e6 --out profiles/r02_kij_200.json"""
import argparse
import csv
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DANDD = os.path.join(ROOT, "dandd_b200", "lib", "dandd")
ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def write_genomes(directory, n, bases, seed=5):
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
        text = b">g%d cluster%d\n" % (g, g // 10) + b"\n".join(r.tobytes() for r in body) + b"\n"
        path = os.path.join(directory, f"g{g:04d}.fa")
        with open(path, "wb") as fh:
            fh.write(text)
        paths.append(path)
    return paths


def run(cmd, env):
    t0 = time.perf_counter()
    p = subprocess.run([sys.executable] + cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode:
        raise RuntimeError(p.stdout[-3000:])
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=200)
    ap.add_argument("--bases", type=float, default=1e6)
    ap.add_argument("--registers", type=int, default=18)
    ap.add_argument("--workdir", default="/tmp/dandd_kij")
    ap.add_argument("--check", type=int, default=5, help="pairs to compare with the oracle")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from dandd_b200 import build
    build.build()
    shutil.rmtree(args.workdir, ignore_errors=True)
    data, out = os.path.join(args.workdir, "fa"), os.path.join(args.workdir, "out")
    t0 = time.perf_counter()
    paths = write_genomes(data, args.genomes, args.bases)
    rep = {"genomes": args.genomes, "bases": args.bases, "registers": args.registers, "pairs": args.genomes * (args.genomes - 1) // 2,
           "write_fastas_s": round(time.perf_counter() - t0, 2)}
    tfile = os.path.join(args.workdir, "timing.jsonl")
    env = dict(os.environ, DANDD_B200_TIMING=tfile)
    rep["tree_wall_s"] = round(run([DANDD, "tree", "-d", data, "-o", out, "-s", "kij", "-k", "14", "-r", str(args.registers)], env), 2)
    pk = os.path.join(out, f"kij_{args.genomes}_dashing_dtree.pickle")
    rep["kij_wall_s"] = round(run([DANDD, "kij", "-d", pk, "-o", out, "--mink", "10", "--maxk", "32"], env), 2)
    recs = [json.loads(ln) for ln in open(tfile)]
    rep["kij_stages"] = recs[-1]["stages"]
    rows = list(csv.DictReader(open(os.path.join(out, f"kij_{args.genomes}_dashing.kij.csv"))))
    rep["kij_rows"] = len(rows)
    kij = np.array([float(r["KIJ"]) for r in rows])
    same = np.array([int(os.path.basename(r["A"])[1:5]) // 10 == int(os.path.basename(r["B"])[1:5]) // 10 for r in rows])
    rep["kij_mean_within_cluster"], rep["kij_mean_between"] = float(kij[same].mean()), float(kij[~same].mean())
    rep["union_files"] = sum(len(fs) for d, _, fs in os.walk(os.path.join(out, "sketchdb", "ngen2")))
    if args.check:
        from oracle import pyoracle as orc
        rng = np.random.default_rng(1)
        worst = 0.0
        for r in [rows[i] for i in rng.integers(0, len(rows), args.check)]:
            sa, sb = (orc.fasta_symbols(open(r[x], "rb").read()) for x in ("A", "B"))
            k = int(r["ABk"])
            u = np.maximum(orc.hll_sketch(sa, k, args.registers), orc.hll_sketch(sb, k, args.registers))
            want = orc.card(u, args.registers) / k
            worst = max(worst, abs(want - float(r["ABdelta"])) / want)
        rep["oracle_pairs_checked"], rep["oracle_max_rel_err_ABdelta"] = args.check, worst
    print(json.dumps(rep))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(rep, fh, indent=1)
    shutil.rmtree(args.workdir, ignore_errors=True)


if __name__ == "__main__":
    main()
