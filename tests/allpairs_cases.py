"""The all-pairs scenarios shared by the CPU run (oracle-backed store double, tests/test_allpairs.py) and
the GPU run (real store, tests/test_allpairs_gpu.py): dandd_b200/helpers/allpairs.py on the inputs of
tests/golden/make_allpairs_golden.py must write what the reference's own functions produced."""
import json
import os

import numpy as np
import pytest

from dandd_b200.helpers import allpairs
from tests.util import make_dataset

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "allpairs_golden.json")
CARD_ABS = 1e-6     # the reference parses `dashing hll` text printed with 6 decimals
REL = 1e-9


def gold_cases():
    with open(GOLD) as fh:
        return json.load(fh)["cases"]


def near(a, b, scale=1.0):
    return a == pytest.approx(b, rel=REL, abs=CARD_ABS * scale)


def write_dataset(tmp, case):
    """The case's FASTAs + an AFproject dataset file naming them (seqid = path without '.fasta')."""
    inputs = make_dataset(os.path.join(tmp, "data"), case["genomes"], case["length"], seed=case["seed"], sub=case["sub"])
    dataset = os.path.join(tmp, "dataset.json")
    with open(dataset, "w") as fh:
        json.dump({"seqids": [f[:-len(".fasta")] for f in inputs], "treids": []}, fh)
    return inputs, dataset


def argv_for(tmp, case, dataset, tool="dashing", **more):
    argv = ["--tool", tool, "--name", os.path.join(tmp, "run"), "--dataset", dataset,
            "--card-results", os.path.join(tmp, "card.tsv"), "--delta-results", os.path.join(tmp, "delta.tsv"),
            "--j-results-phylip", os.path.join(tmp, "sim.phylip"), "--ani-results-phylip", os.path.join(tmp, "ani.phylip"),
            "--nest", str(case["nest"]), "--klist", ",".join(map(str, case["klist"]))]
    if case["extra"]:
        argv += ["--extra=" + case["extra"]]          # (argparse takes a value that begins with "--" only in this form)
    for key, value in more.items():
        argv += ["--" + key.replace("_", "-"), str(value)]
    return argv


def read_tsv(path):
    with open(path) as fh:
        lines = fh.read().splitlines()
    return lines[0].split("\t"), [ln.split("\t") for ln in lines[1:]]


def parse_phylip(text):
    lines = text.strip().split("\n")
    rows = [ln.split(" ") for ln in lines[1:]]
    return int(lines[0]), [r[0] for r in rows], [[float(v) for v in r[1:]] for r in rows]


def check_against_gold(tmp, case, table):
    """card.tsv / delta.tsv / every PHYLIP matrix against what the reference's functions produced."""
    # cardinalities, in the reference's command order
    head, rows = read_tsv(os.path.join(tmp, "card.tsv"))
    assert head == ["tool", "name1", "name2", "k", "card"]
    want = case["results"]
    assert [(r[0], r[1], r[2], int(r[3])) for r in rows] == [tuple(w[:4]) for w in want]
    assert all(near(float(r[4]), w[4]) for r, w in zip(rows, want))
    assert [tuple(r[:4]) for r in table.results()] == [tuple(w[:4]) for w in want]
    # deltas: same argmax k per marginal and pair, rows sorted by name; the reference's column quirk kept
    head, rows = read_tsv(os.path.join(tmp, "delta.tsv"))
    assert head == ["tool", "name1", "name2", "delta", "card", "k"]
    want = case["delta_summary"]
    assert [(r[0], r[1], r[2], int(r[3])) for r in rows] == [(w[0], w[1], w[2], w[5]) for w in want]
    assert all(near(float(r[4]), w[4]) and near(float(r[5]), w[3]) for r, w in zip(rows, want))
    # KIJ / J tuples
    summ = allpairs.kij_summarize(table.delta_summary())
    for k in case["klist"]:
        summ += table.j_summary(k)
    want = case["summary"]
    assert [tuple(s[:4]) + tuple(s[5:]) for s in summ] == [tuple(w[:4]) + tuple(w[5:]) for w in want]
    assert all(near(s[4], w[4], scale=1e-2) for s, w in zip(summ, want))
    # PHYLIP matrices
    for k in [0] + case["klist"]:
        tag = "kij" if k == 0 else "k%d" % k
        for kind in ("sim", "ani"):
            with open(os.path.join(tmp, f"{kind}.{tag}.phylip")) as fh:
                ours = parse_phylip(fh.read())
            gold = parse_phylip(case["phylip"][kind + tag])
            assert ours[:2] == gold[:2]
            for a, b in zip(ours[2], gold[2]):
                assert len(a) == len(b) and all(near(x, y, scale=1e-2) for x, y in zip(a, b)), (kind, tag)


def scenario_gold(tmp, name):
    case = gold_cases()[name]
    inputs, dataset = write_dataset(tmp, case)
    table = allpairs.go(argv_for(tmp, case, dataset))
    assert table.names == case["names"] and os.path.isdir(os.path.join(tmp, "run"))
    check_against_gold(tmp, case, table)
    # the vectorised summaries are the tuple-list functions applied to the full result list
    results = table.results()
    assert table.delta_summary() == allpairs.delta_summarize(results)
    for k in case["klist"]:
        assert table.j_summary(k) == allpairs.j_summarize(results, k)
    assert_files_equal_tuple_path(tmp, table)
    with pytest.raises(RuntimeError, match="already exists"):
        allpairs.go(argv_for(tmp, case, dataset))
    return table


def assert_files_equal_tuple_path(tmp, table, summaries=allpairs):
    """The column-wise writers of CardTable produce, byte for byte, the files that the tuple-list functions
    (`summaries`: this package's, or the reference's module in the live test) write for the same table."""
    results = table.results()
    dsumm = summaries.delta_summarize(results)
    ids = {name: name for name in table.names}
    other = os.path.join(tmp, "tuple_path.phylip")
    for k in [0] + table.ks:
        summ = summaries.kij_summarize(dsumm) if k == 0 else summaries.j_summarize(results, k)
        for kind, ani in (("sim", False), ("ani", True)):
            summaries.summ_to_phylip(summ, ids, other, convert_to_ani=ani)
            with open(other) as a, open(os.path.join(tmp, "%s.%s.phylip" % (kind, "kij" if k == 0 else "k%d" % k))) as b:
                assert a.read() == b.read(), (kind, k)
    head, rows = read_tsv(os.path.join(tmp, "card.tsv"))
    assert rows == [[r[0], r[1], r[2], str(r[3]), str(r[4])] for r in results]
    head, rows = read_tsv(os.path.join(tmp, "delta.tsv"))
    assert rows == [[d[0], d[1], d[2], str(d[5]), str(d[4]), str(d[3])] for d in dsumm]


def scenario_exact(tmp, oracle_counts):
    """--tool kmc: exact distinct canonical k-mer counts of every input and of every pair's union."""
    case = {"genomes": 4, "length": 3000, "seed": 73, "sub": 0.08, "klist": [7, 11, 33], "nest": 4096, "extra": ""}
    inputs, dataset = write_dataset(tmp, case)
    table = allpairs.go(argv_for(tmp, case, dataset, tool="kmc"))
    for c, k in enumerate(case["klist"]):
        for i, f in enumerate(inputs):
            assert table.single[i, c] == oracle_counts([f], k)
        for r, (a, b) in enumerate(table.pairs):
            assert table.pair[r, c] == oracle_counts([inputs[a], inputs[b]], k)
    head, rows = read_tsv(os.path.join(tmp, "card.tsv"))
    assert len(rows) == 3 * (4 + 6) and all(float(r[4]).is_integer() for r in rows)
    assert np.all(table.pair >= np.maximum(table.single[table.pairs[:, 0]], table.single[table.pairs[:, 1]]))
