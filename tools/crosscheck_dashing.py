#!/usr/bin/env python
"""Pin the oracle to a REAL Dashing v1 when one is available.

Parity is otherwise unpinned: the reference ships no arithmetic and neither Dashing's binary nor its
source exists in this environment, so oracle/dandd_oracle.c restates the published algorithms
(several items marked UNVERIFIED).  Run this wherever a `dashing` binary is on
PATH (or pass --dashing /path/to/dashing): it drives the binary with exactly the command lines the
reference builds (lib/sketch_classes.py:351-373, 306-321; helpers/allpairs.py:32-35) on small
adversarial inputs and diffs against the oracle, one verdict per assumption:

    A.1  FASTA rules    records, '@' headers, '+' lines, CRLF, lower case, N / IUPAC breaks, preamble
    A.2-A.5 registers   bit-identical registers for k in {1, 4, 15, 16, 17, 21, 31, 32}, p in {10, 14, 20}
    A.3  --no-canon
    A.6  poly-T >= 32   which of the two behaviours (plain / all-ones sentinel) the binary shows
    A.7  .hll layout    header width (28 or 32 bytes), gzip or raw
    A.8  union -z -o    register-wise max
    A.9  card           Ertl MLE within 1e-6 relative of the oracle estimator (printed with 6 decimals)
    A.10 hll            one sketch over several files; last token of stdout

Without the binary it prints {"status": "skipped"} and exits 0.  Test infrastructure, not product."""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(cmd):
    return subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dashing", default=shutil.which("dashing"))
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if not args.dashing or not os.path.exists(args.dashing):
        print(json.dumps({"status": "skipped", "why": "no `dashing` binary on PATH (parity stays unpinned)"}))
        return 0
    from dandd_b200 import hllfile
    from oracle import pyoracle as orc
    from tests.util import adversarial_fasta, kseq_fasta, random_bases, to_fasta
    dash = args.dashing
    rng = np.random.default_rng(7)
    tmp = tempfile.mkdtemp(prefix="dd_crosscheck_")
    inputs = {
        "adversarial": adversarial_fasta(rng, n=60000),
        "crlf": adversarial_fasta(rng, n=20000).replace(b"\n", b"\r\n"),
        "kseq_markers": kseq_fasta(rng, n=30000),
        "fastq": kseq_fasta(rng, n=30000, fastq=True),
        "polyT": to_fasta([(b"t", np.concatenate([random_bases(rng, 500), np.full(100, ord("T"), dtype=np.uint8), random_bases(rng, 500)]))]),
        "plain": to_fasta([(b"p", random_bases(rng, 300000))], width=80),
    }
    paths = {}
    for name, text in inputs.items():
        paths[name] = os.path.join(tmp, name + ".fa")
        with open(paths[name], "wb") as fh:
            fh.write(text)
    rep = {"status": "ran", "dashing": dash, "version": run(f"{dash} --version").stdout.strip()[:200], "cells": [], "verdicts": {}}

    def sketch_file(name, k, p, canon=True):
        d = os.path.join(tmp, f"{name}_k{k}_p{p}_{int(canon)}")
        os.makedirs(d, exist_ok=True)
        flag = "" if canon else "--no-canon"
        r = run(f"{dash} sketch {flag} -k{k} -S {p} --prefix {d} {paths[name]}")
        out = os.path.join(d, f"{name}.fa.w.{k}.spacing.{p}.hll")
        return out if os.path.exists(out) else None, r

    # registers
    ok_all = True
    for name in inputs:
        sym = orc.fasta_symbols(inputs[name])
        sym_t = orc.polyt_sentinel(sym)
        for k, p in [(1, 10), (4, 10), (15, 14), (16, 14), (17, 14), (21, 20), (31, 14), (32, 14)]:
            for canon in ((True, False) if name == "plain" else (True,)):
                out, r = sketch_file(name, k, p, canon)
                cell = {"input": name, "k": k, "p": p, "canon": canon}
                if out is None:
                    cell["error"] = (r.stderr or r.stdout)[-300:]
                    ok_all = False
                else:
                    raw = open(out, "rb").read()
                    cell["file_bytes"], cell["gzip"] = len(raw), raw[:2] == b"\x1f\x8b"
                    try:
                        regs, gp, cached = hllfile.read_hll(out)
                        cell["header_bytes"] = (len(raw) if not cell["gzip"] else len(__import__("zlib").decompress(raw, 31))) - (1 << gp)
                        cell["equal_plain"] = bool(np.array_equal(regs, orc.hll_sketch(sym, k, p, canon)))
                        cell["equal_polyt_sentinel"] = bool(np.array_equal(regs, orc.hll_sketch(sym_t, k, p, canon)))
                        ok_all &= cell["equal_plain"] or cell["equal_polyt_sentinel"]
                    except Exception as e:  # noqa: BLE001
                        cell["error"] = f"unreadable sketch file: {e}"
                        ok_all = False
                rep["cells"].append(cell)
    rep["verdicts"]["registers_bit_identical"] = ok_all
    pt = [c for c in rep["cells"] if c["input"] == "polyT" and c["k"] >= 31 and "equal_plain" in c]
    if pt:
        rep["verdicts"]["A.6_polyT"] = ("plain (poly-T is valid sequence)" if all(c["equal_plain"] for c in pt) else
                                        "all-ones sentinel (set DANDD_B200_POLYT_SENTINEL=1)" if all(c["equal_polyt_sentinel"] for c in pt)
                                        else "NEITHER restatement matches")
    hb = {c.get("header_bytes") for c in rep["cells"] if "header_bytes" in c}
    rep["verdicts"]["A.7_header_bytes"] = sorted(hb)
    # card / union / hll on the plain input
    a, _ = sketch_file("plain", 21, 14)
    b, _ = sketch_file("adversarial", 21, 14)
    if a and b:
        r = run(f"{dash} card --presketched {a} {b}")
        lines = [ln.split("\t") for ln in r.stdout.splitlines()[1:] if "\t" in ln]
        got = {ln[0]: float(ln[1]) for ln in lines}
        want = {a: orc.card(hllfile.read_hll(a)[0], 14), b: orc.card(hllfile.read_hll(b)[0], 14)}
        rep["card"] = {"header_line": r.stdout.splitlines()[0] if r.stdout else "", "got": got, "oracle": want}
        rep["verdicts"]["A.9_card_within_1e-6"] = all(abs(got.get(pth, 0) - want[pth]) <= 1e-6 * want[pth] + 1e-6 for pth in want)
        u = os.path.join(tmp, "u.hll")
        run(f"{dash} union -z -o {u} {a} {b}")
        if os.path.exists(u):
            rep["verdicts"]["A.8_union_is_max"] = bool(np.array_equal(hllfile.read_hll(u)[0],
                                                                      np.maximum(hllfile.read_hll(a)[0], hllfile.read_hll(b)[0])))
        r = run(f"{dash} hll -k 21 -S 14 {paths['plain']} {paths['adversarial']}")
        try:
            est = float(r.stdout.split()[-1])
            both = np.maximum(orc.hll_sketch(orc.fasta_symbols(inputs["plain"]), 21, 14), orc.hll_sketch(orc.fasta_symbols(inputs["adversarial"]), 21, 14))
            rep["verdicts"]["A.10_hll_last_token"] = abs(est - orc.card(both, 14)) <= 1e-6 * est + 1e-6
        except Exception as e:  # noqa: BLE001
            rep["verdicts"]["A.10_hll_last_token"] = f"could not parse: {e}"
    print(json.dumps(rep["verdicts"], indent=1))
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(rep, fh, indent=1)
    shutil.rmtree(tmp, ignore_errors=True)
    return 0 if all(v is True or not isinstance(v, bool) for v in rep["verdicts"].values()) else 1


if __name__ == "__main__":
    sys.exit(main())
