"""Runs this repository's drop-in `dandd` commands in-process and normalises their outputs the same
way tests/golden/make_reference_golden.py normalised the reference's, so the two can be compared."""
import csv
import json
import os
import pickle

import pytest

import dandd_b200

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs.json")
CARD_ABS = 2e-6    # the reference parses `dashing card` text printed with 6 decimals
CARD_REL = 1e-9


def gold_runs():
    with open(GOLD) as fh:
        return json.load(fh)["runs"]


def run_dandd(argv):
    dandd_b200.enable_compat()
    import dandd_cmd
    parser, _ = dandd_cmd.parse_arguments()
    args = parser.parse_args(argv)
    args.func(args)


def read_csv(path):
    with open(path, newline="") as fh:
        return list(csv.DictReader(fh))


def rel(path, base):
    return os.path.relpath(path, base) if path and os.path.isabs(path) else path


def norm_fastas(text, sep):
    return sep.join(os.path.basename(f.strip(" '[]")) for f in text.split(sep))


def collect_tree(outdir, prefix, sketchdir, tool):
    rows = [{"title": r["title"], "ngen": int(r["ngen"]), "k": int(r["k"]), "delta": float(r["delta"]),
             "card": float(r["card"]), "sketchloc": rel(r["sketchloc"], sketchdir), "fastas": norm_fastas(r["fastas"], "|")}
            for r in read_csv(os.path.join(outdir, prefix + "_deltas.csv"))]
    files = sorted(os.path.relpath(os.path.join(d, f), sketchdir) for d, _, fs in os.walk(sketchdir) for f in fs
                   if not f.endswith((".pickle", ".bkp")))
    cardkey = {}
    for name in os.listdir(sketchdir):
        if name.endswith(f"_{tool}_cardinalities.pickle"):
            with open(os.path.join(sketchdir, name), "rb") as fh:
                cardkey.update({rel(k, sketchdir): float(v) for k, v in pickle.load(fh).items()})
    with open(os.path.join(sketchdir, "dandd_fastahex.pickle"), "rb") as fh:
        fastahex = pickle.load(fh)
    with open(os.path.join(sketchdir, "dandd_sketchinfo.pickle"), "rb") as fh:
        sketchinfo = sorted(pickle.load(fh).keys())
    return {"deltas": rows, "files": files, "cardkey": cardkey, "fastahex": fastahex, "sketchinfo": sketchinfo}


def close(a, b):
    return a == pytest.approx(b, rel=CARD_REL, abs=CARD_ABS)


def assert_tree_matches(ours, gold, exact=False):
    assert [(r["title"], r["ngen"], r["k"], r["sketchloc"], r["fastas"]) for r in ours["deltas"]] == \
           [(r["title"], r["ngen"], r["k"], r["sketchloc"], r["fastas"]) for r in gold["deltas"]]
    for a, b in zip(ours["deltas"], gold["deltas"]):
        if exact:
            assert a["card"] == b["card"] and a["delta"] == pytest.approx(b["delta"], rel=1e-12)
        else:
            assert close(a["card"], b["card"]) and close(a["delta"], b["delta"]), (a, b)
    assert ours["files"] == gold["files"]
    assert sorted(ours["cardkey"]) == sorted(gold["cardkey"])
    for key, value in gold["cardkey"].items():
        assert (ours["cardkey"][key] == value) if exact else close(ours["cardkey"][key], value), key
    assert ours["fastahex"] == gold["fastahex"]
    assert ours["sketchinfo"] == gold["sketchinfo"]
