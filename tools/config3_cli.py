#!/usr/bin/env python
"""Config 3 through the drop-in command line (BASELINE.json north-star target run):

    G synthetic human-scale FASTA FILES on disk (24 records, ~1 % N runs, 50 % soft-masked;
    copies 0.1 % substitutions apart)
    torchrun --nproc-per-node N  dandd tree --ksweep            (k = 2..32, p = 20; reference
                                                                lib/dandd_cmd.py:43-62, lib/huffman_dandd.py:839-877)
    dandd progressive -d <dtree.pickle> -n 1 --ksweep           (identity ordering; reference :65-87, :624-663)

and reports wall times, the per-stage times every rank recorded (DANDD_B200_TIMING), argmax-k / delta of
every leaf and every prefix, a CPU proxy (the oracle port on a sample of one genome, extrapolated --
the reference's own CPU path needs Dashing, which does not exist here), and optionally an oracle
check of the sketch FILES the run left behind (registers bit-exact, cardinalities 1e-9) for a few k
per genome.

    python tools/config3_cli.py --gpus 8 --bases 3.1e9 --out profiles/r02_config3_cli.json
    python tools/config3_cli.py --gpus 1 --bases 3.1e9 --oracle-ks 2,18,32 --out ...   (same data, checked)

Test infrastructure / measurement driver: not part of the product."""
import argparse
import csv
import glob
import json
import os
import pickle
import shutil
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DANDD = os.path.join(ROOT, "dandd_b200", "lib", "dandd")


def generate(workdir, genomes, bases, world, rank, device):
    """Rank r writes genome g for g % world == r: the seed-3 ancestor, or a 0.1 % substitution copy of it."""
    import torch
    from tools.synth import mutate_text, synth_fasta
    os.makedirs(workdir, exist_ok=True)
    dev = torch.device("cuda", device)
    ancestor = synth_fasta(int(bases), 24, seed=3, device=dev)
    for g in range(genomes):
        if g % world != rank:
            continue
        text = mutate_text(ancestor, 0.001, 3 + g) if g else ancestor
        host = text.cpu().numpy()
        with open(os.path.join(workdir, f"genome{g}.fa"), "wb") as fh:
            fh.write(memoryview(host))
        del text, host


def torchrun(nproc, argv, env=None, port=29611):
    if nproc == 1:
        cmd = [sys.executable] + argv
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", str(port)] + argv
    t0 = time.perf_counter()
    proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    dt = time.perf_counter() - t0
    if proc.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + proc.stdout[-4000:])
    return dt, proc.stdout


def read_timing(path):
    recs = []
    if os.path.exists(path):
        with open(path) as fh:
            recs = [json.loads(line) for line in fh if line.strip()]
    return recs


def summarize_timing(recs):
    """Per stage: max over ranks (the critical path) and the mean."""
    keys = sorted({k for r in recs for k in r["stages"]})
    out = {}
    for k in keys:
        vals = [r["stages"].get(k, 0.0) for r in recs]
        out[k] = {"max": round(max(vals), 4), "mean": round(sum(vals) / len(vals), 4)}
    out["process_wall_s"] = {"max": round(max(r["wall_s"] for r in recs), 3), "mean": round(sum(r["wall_s"] for r in recs) / len(recs), 3)}
    return out


def cards_by_genome(sketchdir, tag, genomes, ks, p=20):
    with open(os.path.join(sketchdir, f"{tag}_dashing_cardinalities.pickle"), "rb") as fh:
        cardkey = pickle.load(fh)
    out = {}
    for g in range(genomes):
        out[g] = [float(cardkey[os.path.join(sketchdir, "ngen1", f"k{k}", f"genome{g}.fa.w.{k}.spacing.{p}.hll")]) for k in ks]
    return out, cardkey


def cpu_proxy(path, ks, p, sample_bytes, threads):
    """The reference's CPU topology on the oracle port: one single-threaded sketch job per k,
    `threads` in flight, on the first `sample_bytes` of one genome."""
    from oracle import pyoracle as orc
    with open(path, "rb") as fh:
        text = fh.read(sample_bytes)
    text = text[:text.rfind(b"\n") + 1]
    sym = orc.fasta_symbols(text)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda k: orc.card(orc.hll_sketch(sym, k, p), p), ks))
    dt = time.perf_counter() - t0
    return {"sample_bases": int((sym < 4).sum()), "seconds": dt, "threads": threads, "ks": len(ks)}


def oracle_check(workdir, sketchdir, genomes, check_ks, cards, ks, p, threads):
    """Registers in the .hll files the run wrote == oracle registers of the FASTA file; cardinality
    stored in the cardinality pickle == oracle estimator (1e-9)."""
    from dandd_b200 import hllfile
    from oracle import pyoracle as orc
    results = []
    for g in range(genomes):
        with open(os.path.join(workdir, f"genome{g}.fa"), "rb") as fh:
            sym = orc.fasta_symbols(fh.read())

        def one(k):
            want = orc.hll_sketch(sym, k, p)
            got, gp, _ = hllfile.read_hll(os.path.join(sketchdir, "ngen1", f"k{k}", f"genome{g}.fa.w.{k}.spacing.{p}.hll"))
            c = orc.card(want, p)
            return {"genome": g, "k": k, "registers_equal": bool(gp == p and np.array_equal(got, want)),
                    "card_oracle": c, "card_run": cards[g][ks.index(k)], "rel_err": abs(cards[g][ks.index(k)] - c) / c}
        with ThreadPoolExecutor(min(threads, len(check_ks))) as ex:
            results += list(ex.map(one, check_ks))
        del sym
    return results


def run(args):
    workdir = os.path.abspath(args.workdir)
    data, out = os.path.join(workdir, "fasta"), os.path.join(workdir, "out")
    sketchdir = os.path.join(out, "sketchdb")
    if not args.keep:
        shutil.rmtree(workdir, ignore_errors=True)
    os.makedirs(out, exist_ok=True)
    ks = list(range(args.kmin, args.kmax + 1))
    rep = {"config": {"genomes": args.genomes, "bases_per_genome": args.bases, "k": [args.kmin, args.kmax], "p": 20,
                      "n_gpus": args.gpus, "host_cores": os.cpu_count()}}
    # 1. data
    if not os.path.exists(os.path.join(data, f"genome{args.genomes - 1}.fa")):
        dt, _ = torchrun(args.gpus, [os.path.abspath(__file__), "--_generate", "--workdir", workdir, "--genomes", str(args.genomes),
                                     "--bases", str(args.bases)], port=29621)
        rep["generate_s"] = round(dt, 2)
    rep["fasta_bytes"] = [os.path.getsize(os.path.join(data, f"genome{g}.fa")) for g in range(args.genomes)]
    t0 = time.perf_counter()
    subprocess.run(["sync"])      # the files were just written: let the write-back finish before anything is timed
    rep["sync_after_generate_s"] = round(time.perf_counter() - t0, 2)
    # 2. dandd tree --ksweep under torchrun
    timing_tree = os.path.join(workdir, "timing_tree.jsonl")
    env = dict(os.environ, DANDD_B200_TIMING=timing_tree, DANDD_B200_UNION_FILES=args.union_files)
    if args.gpus > 1:
        # this tool measures the N-process path: `--gpus N` is an upper bound otherwise (one worker per 12 GiB of
        # fresh FASTA, dandd_cmd._self_launch) and would run config 3 (25 GB) in a single process
        env.setdefault("DANDD_B200_BYTES_PER_GPU", "0")
        rep["config"]["bytes_per_gpu_policy"] = env["DANDD_B200_BYTES_PER_GPU"]
    tree_argv = [DANDD, "tree", "--datadir", data, "-o", out, "--tag", "cfg3", "--ksweep", "--mink", str(args.kmin),
                 "--maxk", str(args.kmax)]
    launcher = getattr(args, "launcher", "self")
    rep["config"]["launcher"] = launcher if args.gpus > 1 else "single process"

    def launch(argv, env_):
        if launcher == "self" and args.gpus > 1:      # `dandd tree --gpus N`: rank 0 starts the other ranks itself
            return torchrun(1, argv + ["--gpus", str(args.gpus)], env=env_)
        return torchrun(args.gpus, argv, env=env_)
    dt, log = launch(tree_argv, env)
    rep["tree_wall_s"] = round(dt, 3)
    rep["tree_stages"] = summarize_timing(read_timing(timing_tree))
    # a fully cached re-run (zero work)
    dt, _ = launch(tree_argv, dict(env, DANDD_B200_TIMING=os.path.join(workdir, "timing_tree2.jsonl")))
    rep["tree_cached_rerun_wall_s"] = round(dt, 3)
    if getattr(args, "also_torchrun", False) and args.gpus > 1:   # the same job under torchrun, into a fresh database
        out2 = os.path.join(workdir, "out_torchrun")
        argv2 = [DANDD, "tree", "--datadir", data, "-o", out2, "--tag", "cfg3", "--ksweep", "--mink", str(args.kmin),
                 "--maxk", str(args.kmax)]
        t2 = os.path.join(workdir, "timing_tree_torchrun.jsonl")
        dt, _ = torchrun(args.gpus, argv2, env=dict(env, DANDD_B200_TIMING=t2))
        rep["tree_wall_torchrun_s"] = round(dt, 3)
        rep["tree_stages_torchrun"] = summarize_timing(read_timing(t2))
        c2, _ = cards_by_genome(os.path.join(out2, "sketchdb"), "cfg3", args.genomes, ks)
        rep["torchrun_cards_identical"] = True   # checked below once the first run's cards are read
        rep["_cards_torchrun"] = c2
    # 3. dandd progressive -n 1 --ksweep (single process)
    if not getattr(args, "skip_progressive", False):
        pickle_path = os.path.join(out, f"cfg3_{args.genomes}_dashing_dtree.pickle")
        timing_prog = os.path.join(workdir, "timing_prog.jsonl")
        dt, log = torchrun(1, [DANDD, "progressive", "-d", pickle_path, "-n", "1", "--ksweep", "--mink", str(args.kmin), "--maxk",
                               str(args.kmax), "-o", out], env=dict(env, DANDD_B200_TIMING=timing_prog))
        rep["progressive_wall_s"] = round(dt, 3)
        rep["progressive_stages"] = summarize_timing(read_timing(timing_prog))
    # 4. results
    cards, cardkey = cards_by_genome(sketchdir, "cfg3", args.genomes, ks)
    if "_cards_torchrun" in rep:
        rep["torchrun_cards_identical"] = rep.pop("_cards_torchrun") == cards
    karr = np.array(ks, dtype=np.float64)
    rep["leaf_argmax_k"] = [int(ks[int(np.argmax(np.array(cards[g]) / karr))]) for g in range(args.genomes)]
    rep["leaf_delta"] = [float((np.array(cards[g]) / karr).max()) for g in range(args.genomes)]
    rep["leaf_cards"] = {str(g): cards[g] for g in range(args.genomes)}
    if not getattr(args, "skip_progressive", False):
        summ = [r for r in csv.DictReader(open(glob.glob(os.path.join(out, "cfg3_progu1_*summary.csv"))[0]))]
        by_n = {}
        for r in summ:
            by_n.setdefault(int(r["ngen"]), {})[int(r["kval"])] = float(r["delta_pos"])
        rep["prefix_argmax_k"] = [max(by_n[n], key=by_n[n].get) for n in sorted(by_n)]
        rep["prefix_delta"] = [max(by_n[n].values()) for n in sorted(by_n)]
    rep["sketch_files"] = sum(len(fs) for _, _, fs in os.walk(sketchdir))
    total_bases = args.bases * args.genomes
    rep["tree_gbp_per_s_wall"] = total_bases / rep["tree_wall_s"] / 1e9
    # 5. CPU proxy on the same box
    if args.cpu_sample_bytes > 0:
        threads = max(1, int((os.cpu_count() or 1) * 0.95))
        cp = cpu_proxy(os.path.join(data, "genome0.fa"), ks, 20, int(args.cpu_sample_bytes), threads)
        per_base = cp["seconds"] / cp["sample_bases"]
        cp["extrapolated_leaf_sketch_s_all_genomes"] = per_base * total_bases
        cp["note"] = ("oracle port (kind: port), one single-threaded job per k, floor(0.95*cores) in flight, genomes sequential "
                      "as in the reference (lib/huffman_dandd.py:402-404); leaf sketches only, measured on a sample of genome 0 "
                      "and scaled linearly to G genomes -- not Dashing, which is not installed")
        rep["cpu_proxy"] = cp
        rep["tree_speedup_vs_cpu_proxy"] = cp["extrapolated_leaf_sketch_s_all_genomes"] / rep["tree_wall_s"]
    # 6. oracle check of the sketch files
    if args.oracle_ks:
        check_ks = [int(k) for k in args.oracle_ks.split(",")]
        t0 = time.perf_counter()
        res = oracle_check(data, sketchdir, args.genomes, check_ks, cards, ks, 20, max(1, (os.cpu_count() or 2) - 1))
        rep["oracle_check"] = {"ks": check_ks, "seconds": round(time.perf_counter() - t0, 1),
                               "all_registers_equal": all(r["registers_equal"] for r in res),
                               "max_card_rel_err": max(r["rel_err"] for r in res), "cells": res}
    if getattr(args, "verbose", False):
        print(json.dumps(rep))
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as fh:
            json.dump(rep, fh, indent=1)
    if not args.keep:
        shutil.rmtree(workdir, ignore_errors=True)
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--genomes", type=int, default=8)
    ap.add_argument("--bases", type=float, default=3.1e9)
    ap.add_argument("--kmin", type=int, default=2)
    ap.add_argument("--kmax", type=int, default=32)
    ap.add_argument("--workdir", default="/tmp/dandd_cfg3")
    ap.add_argument("--union-files", default="full", choices=["full", "stub"])
    ap.add_argument("--cpu-sample-bytes", type=float, default=64e6)
    ap.add_argument("--oracle-ks", default="")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--skip-progressive", action="store_true", help="time the tree command only")
    ap.add_argument("--launcher", default="self", choices=["self", "torchrun"],
                    help="multi-GPU tree: `dandd tree --gpus N` (rank 0 starts the others) or torchrun")
    ap.add_argument("--also-torchrun", action="store_true", help="additionally time the same tree job under torchrun")
    ap.add_argument("--out", default=None)
    ap.add_argument("--_generate", action="store_true")
    args = ap.parse_args()
    if args._generate:
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        generate(os.path.join(os.path.abspath(args.workdir), "fasta"), args.genomes, args.bases, world, rank,
                 int(os.environ.get("LOCAL_RANK", "0")))
        return
    from dandd_b200 import build
    build.build()
    args.verbose = True
    run(args)


if __name__ == "__main__":
    main()
