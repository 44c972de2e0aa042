"""The end-to-end drop-in scenarios, shared by the CPU run (oracle-backed store double,
tests/test_host_logic.py) and the GPU run (real store, tests/test_host_gpu.py).  Each scenario
replays on the same deterministic inputs what tests/golden/make_reference_golden.py ran through the
unmodified reference, and requires the same outputs."""
import os
import pickle

import pytest

from tests.host_harness import (assert_tree_matches, close, collect_tree, gold_runs, norm_fastas, read_csv, run_dandd)
from tests.util import make_dataset


def scenario_tree_hillclimb(tmp):
    data = make_dataset(os.path.join(tmp, "data5"), 5, 20000, seed=21)
    out = os.path.join(tmp, "outA")
    run_dandd(["tree", "-d", os.path.dirname(data[0]), "-s", "runA", "-k", "14", "-o", out])
    ours = collect_tree(out, "runA_5_dashing", os.path.join(out, "sketchdb"), "dashing")
    assert_tree_matches(ours, gold_runs()["A_tree_hillclimb"])
    return out


def scenario_rerun_is_fully_cached(tmp, store):
    """A second identical run must not sketch, union or estimate anything."""
    out = scenario_tree_hillclimb(tmp)
    before = dict(store.stats)
    run_dandd(["tree", "-d", os.path.join(tmp, "data5"), "-s", "runA", "-k", "14", "-o", out])
    after = store.stats
    assert after["leaf_passes"] == before["leaf_passes"] and after["union_launches"] == before["union_launches"]
    assert after["files_written"] == before["files_written"]


def scenario_ksweep_and_progressive(tmp):
    data = make_dataset(os.path.join(tmp, "data5"), 5, 20000, seed=21)
    out = os.path.join(tmp, "outB")
    sweep = ["--ksweep", "--mink", "10", "--maxk", "16"]
    run_dandd(["tree", "-d", os.path.dirname(data[0]), "-s", "runB", "-k", "14", "-o", out] + sweep)
    gold = gold_runs()
    assert_tree_matches(collect_tree(out, "runB_5_dashing", os.path.join(out, "sketchdb"), "dashing"), gold["B_tree_ksweep"])
    dtree = os.path.join(out, "runB_5_dashing_dtree.pickle")

    run_dandd(["progressive", "-d", dtree, "-n", "1", "-o", out] + sweep)
    rows = read_csv(os.path.join(out, "runB_progu1_5_dashing.csv"))
    want = gold["B_progressive_identity"]
    assert [(int(r["ngen"]), int(r["kval"]), r["delta"], int(r["ordering"]), norm_fastas(r["fastas"], ",")) for r in rows] == \
           [(r["ngen"], r["kval"], r["delta"], r["ordering"], r["fastas"]) for r in want["rows"]]
    summ = read_csv(os.path.join(out, "runB_progu1_5_dashingsummary.csv"))
    assert len(summ) == len(want["summary"])
    for a, b in zip(summ, want["summary"]):
        assert (int(a["ngen"]), int(a["kval"]), a["title"], int(a["ordering"])) == (b["ngen"], b["kval"], b["title"], b["ordering"])
        assert close(float(a["card"]), b["card"]) and close(float(a["delta_pos"]), b["delta_pos"])

    want = gold["B_progressive_orderings"]
    ofile = os.path.join(tmp, "orderings.pickle")
    with open(ofile, "wb") as fh:
        pickle.dump({tuple(o) for o in want["orderings"]}, fh)
    run_dandd(["progressive", "-d", dtree, "-r", ofile, "-s", "runBo", "-o", out] + sweep)
    summ = read_csv(os.path.join(out, "runBo_progu0_5_dashingsummary.csv"))
    cells = sorted({(r["title"], int(r["kval"]), float(r["card"])) for r in summ})
    assert [(t, k) for t, k, _ in cells] == [(t, k) for t, k, _ in want["cells"]]
    for (_, _, a), (_, _, b) in zip(cells, want["cells"]):
        assert close(a, b)


def scenario_progressive_hillclimb_and_kij(tmp):
    out = scenario_tree_hillclimb(tmp)
    gold = gold_runs()
    dtree = os.path.join(out, "runA_5_dashing_dtree.pickle")
    run_dandd(["progressive", "-d", dtree, "-n", "1", "-o", out])
    rows = read_csv(os.path.join(out, "runA_progu1_5_dashing.csv"))
    want = gold["F_progressive_hillclimb"]["rows"]
    assert [(int(r["ngen"]), int(r["kval"]), norm_fastas(r["fastas"], ",")) for r in rows] == \
           [(r["ngen"], r["kval"], r["fastas"]) for r in want]
    for a, b in zip(rows, want):
        assert close(float(a["delta"]), b["delta"])

    run_dandd(["kij", "-d", dtree, "-o", out, "--jaccard", "--mink", "12", "--maxk", "14"])
    kij = read_csv(os.path.join(out, "runA_5_dashing.kij.csv"))
    want = gold["C_kij"]
    assert len(kij) == len(want["kij"]) == 10
    for a, b in zip(kij, want["kij"]):
        assert (a["Atitle"], a["Btitle"], int(a["Ak"]), int(a["Bk"]), int(a["ABk"])) == \
               (b["Atitle"], b["Btitle"], b["Ak"], b["Bk"], b["ABk"])
        for f in ("Adelta", "Bdelta", "ABdelta"):
            assert close(float(a[f]), b[f])
        assert float(a["KIJ"]) == pytest.approx(b["KIJ"], rel=1e-7)
    jac = read_csv(os.path.join(out, "runA_5_dashing.j.csv"))
    assert len(jac) == len(want["jaccard"])
    for a, b in zip(jac, want["jaccard"]):
        assert (a["Atitle"], a["Btitle"], int(a["kval"])) == (b["Atitle"], b["Btitle"], b["kval"])
        for f in ("Acard", "Bcard", "ABcard"):
            assert close(float(a[f]), b[f])
        assert float(a["jaccard"]) == pytest.approx(b["jaccard"], rel=1e-6, abs=1e-9)


def scenario_tree_nchildren(tmp):
    data = make_dataset(os.path.join(tmp, "data7"), 7, 12000, seed=22, prefix="h")
    out = os.path.join(tmp, "outD")
    run_dandd(["tree", "-d", os.path.dirname(data[0]), "-s", "runD", "-k", "13", "-o", out, "-n", "3"])
    assert_tree_matches(collect_tree(out, "runD_7_dashing", os.path.join(out, "sketchdb"), "dashing"),
                        gold_runs()["D_tree_nchildren3"])


def scenario_config1_tree(tmp):
    """BASELINE.json config 1: `dandd tree -k 14` on 25 mito-like genomes (the example/ tutorial shape)."""
    data = make_dataset(os.path.join(tmp, "mito25"), 25, 16500, seed=1, sub=0.10, indel=0.003, prefix="mito")
    out = os.path.join(tmp, "outG")
    run_dandd(["tree", "-d", os.path.dirname(data[0]), "-s", "fish-mito", "-k", "14", "-o", out])
    assert_tree_matches(collect_tree(out, "fish-mito_25_dashing", os.path.join(out, "sketchdb"), "dashing"),
                        gold_runs()["G_config1_tree"])


def scenario_tree_exact(tmp):
    data = make_dataset(os.path.join(tmp, "data5"), 5, 20000, seed=21)
    out = os.path.join(tmp, "outE")
    run_dandd(["tree", "-d", os.path.dirname(data[0]), "-s", "runE", "-k", "14", "-o", out, "--exact"])
    assert_tree_matches(collect_tree(out, "runE_5_kmc", os.path.join(out, "sketchdb"), "kmc"),
                        gold_runs()["E_tree_exact"], exact=True)


def scenario_exact_sweep_above_k32(tmp):
    """`--exact --ksweep` across k = 32 (README.md:82: higher ks need --exact): every node's count at
    every k equals the oracle's distinct canonical k-mer count of its FASTAs."""
    from oracle import pyoracle as orc
    data = make_dataset(os.path.join(tmp, "data3"), 3, 12000, seed=33)
    out = os.path.join(tmp, "outH")
    run_dandd(["tree", "-d", os.path.dirname(data[0]), "-s", "runH", "-k", "31", "-o", out, "--exact", "--ksweep",
               "--mink", "30", "--maxk", "36"])
    with open(os.path.join(out, "runH_3_kmc_dtree.pickle"), "rb") as fh:
        tree = pickle.load(fh)
    syms = {f: orc.fasta_symbols(open(f, "rb").read()) for f in tree.fastas}
    todo, seen = [tree.root], 0
    while todo:
        node = todo.pop()
        todo.extend(node.children)
        for k in range(30, 37):
            assert node.ksketches[k] is not None, (node.node_title, k)
            assert int(node.ksketches[k].card) == orc.exact_count([syms[f] for f in node.fastas], k, True), (node.node_title, k)
            seen += 1
    assert seen >= 7 * 4


def scenario_option_coverage(tmp):
    """--no-canon / --registers, -f file lists with --safe, --fast, progressive -f and --step, kij --afproject:
    the same commands the golden generator ran through the unmodified reference."""
    gold = gold_runs()
    data = make_dataset(os.path.join(tmp, "data5"), 5, 20000, seed=21)
    ddir = os.path.dirname(data[0])
    outH = os.path.join(tmp, "outH")
    run_dandd(["tree", "-d", ddir, "-s", "runH", "-k", "12", "-o", outH, "-C", "-r", "12"])
    assert_tree_matches(collect_tree(outH, "runH_5_dashing", os.path.join(outH, "sketchdb"), "dashing"), gold["H_tree_noncanon_p12"])

    flist = os.path.join(tmp, "flist.txt")
    with open(flist, "w") as fh:
        fh.write("\n".join(os.path.join(ddir, f) for f in sorted(os.listdir(ddir))[1:4]) + "\n")
    outI = os.path.join(tmp, "outI")
    run_dandd(["tree", "-f", flist, "-s", "runI", "-k", "14", "-o", outI, "--safe"])
    assert_tree_matches(collect_tree(outI, "runI_3_dashing", os.path.join(outI, "sketchdb"), "dashing"), gold["I_tree_flist_safe"])

    outK = os.path.join(tmp, "outK")
    run_dandd(["tree", "-d", ddir, "-s", "runK", "-k", "14", "-o", outK, "--fast"])
    want = gold["K_tree_fast"]
    assert sorted(f for f in os.listdir(outK) if f != "sketchdb") == want["outputs"]
    rows = read_csv(os.path.join(outK, "runK_5_dashing_deltas.csv"))
    assert [(r["title"], int(r["ngen"]), int(r["k"])) for r in rows] == [(r["title"], r["ngen"], r["k"]) for r in want["deltas"]]
    for a, b in zip(rows, want["deltas"]):
        assert close(float(a["delta"]), b["delta"]) and close(float(a["card"]), b["card"])

    outA = os.path.join(tmp, "outA")
    run_dandd(["tree", "-d", ddir, "-s", "runA", "-k", "14", "-o", outA])
    dtree = os.path.join(outA, "runA_5_dashing_dtree.pickle")

    def prog_matches(csv_name, want_rows):
        got = read_csv(os.path.join(outA, csv_name))
        assert [(int(r["ngen"]), int(r["kval"]), norm_fastas(r["fastas"], ",")) for r in got] == \
               [(r["ngen"], r["kval"], r["fastas"]) for r in want_rows]
        for a, b in zip(got, want_rows):
            assert close(float(a["delta"]), b["delta"])

    run_dandd(["progressive", "-d", dtree, "-f", flist, "-n", "1", "-o", outA, "-s", "sub"])
    prog_matches("sub_progu1_5_dashing.csv", gold["L_progressive_subset"]["rows"])
    run_dandd(["progressive", "-d", dtree, "-n", "1", "--step", "2", "-o", outA, "-s", "st2"])
    prog_matches("st2_progu1_5_dashing.csv", gold["N_progressive_step2"]["rows"])

    run_dandd(["kij", "-d", dtree, "-o", outA, "--afproject", "-s", "af"])
    with open(os.path.join(outA, "af_5_dashing_AFtuples.pickle"), "rb") as fh:
        tuples = pickle.load(fh)
    want = gold["O_kij_afproject"]
    assert type(tuples).__name__ == want["type"] and len(tuples) == len(want["tuples"])
    key = lambda t: (t[1], t[2])
    for a, b in zip(sorted((list(t) for t in tuples), key=key), sorted(want["tuples"], key=key)):
        assert a[:4] == b[:4] and a[5:] == b[5:] and close(float(a[4]), float(b[4]))


def scenario_exact_sweep_progressive_and_binary_tree(tmp):
    """--exact with a k sweep, progressive exact unions (with the sweep and with the hill-climb) and a
    2-children tree, against the goldens of the unmodified reference."""
    gold = gold_runs()
    data = make_dataset(os.path.join(tmp, "data5"), 5, 20000, seed=21)
    ddir = os.path.dirname(data[0])
    outP = os.path.join(tmp, "outP")
    sweep = ["--ksweep", "--mink", "12", "--maxk", "15"]
    run_dandd(["tree", "-d", ddir, "-s", "runP", "-k", "14", "-o", outP, "--exact"] + sweep)
    assert_tree_matches(collect_tree(outP, "runP_5_kmc", os.path.join(outP, "sketchdb"), "kmc"), gold["P_tree_exact_ksweep"],
                        exact=True)
    dtree = os.path.join(outP, "runP_5_kmc_dtree.pickle")
    run_dandd(["progressive", "-d", dtree, "-n", "1", "-o", outP] + sweep)
    summ = read_csv(os.path.join(outP, "runP_progu1_5_kmcsummary.csv"))
    want = gold["Q_progressive_exact_ksweep"]["summary"]
    assert [(int(r["ngen"]), int(r["kval"]), r["title"], float(r["card"])) for r in summ] == \
           [(r["ngen"], r["kval"], r["title"], r["card"]) for r in want]
    for a, b in zip(summ, want):
        assert close(float(a["delta_pos"]), b["delta_pos"])
    run_dandd(["progressive", "-d", dtree, "-n", "1", "-o", outP, "-s", "hc"])
    got = read_csv(os.path.join(outP, "hc_progu1_5_kmc.csv"))
    want = gold["Q_progressive_exact_hillclimb"]["rows"]
    assert [(int(r["ngen"]), int(r["kval"]), norm_fastas(r["fastas"], ",")) for r in got] == \
           [(r["ngen"], r["kval"], r["fastas"]) for r in want]
    for a, b in zip(got, want):
        assert close(float(a["delta"]), b["delta"])

    outS = os.path.join(tmp, "outS")
    run_dandd(["tree", "-d", ddir, "-s", "runS", "-k", "13", "-o", outS, "-n", "2"])
    assert_tree_matches(collect_tree(outS, "runS_5_dashing", os.path.join(outS, "sketchdb"), "dashing"), gold["S_tree_nchildren2"])


def scenario_default_sweep(tmp):
    """`tree --ksweep` with the default range k = 2..32 (small k included) on three short genomes."""
    data = make_dataset(os.path.join(tmp, "data3"), 3, 6000, seed=23, prefix="s")
    out = os.path.join(tmp, "outT")
    run_dandd(["tree", "-d", os.path.dirname(data[0]), "-s", "runT", "-k", "10", "-o", out, "--ksweep"])
    assert_tree_matches(collect_tree(out, "runT_3_dashing", os.path.join(out, "sketchdb"), "dashing"),
                        gold_runs()["T_tree_default_sweep"])


def scenario_pickle_roundtrip(tmp):
    """The dtree pickle names classes by the reference's top-level module names."""
    out = scenario_tree_hillclimb(tmp)
    raw = open(os.path.join(out, "runA_5_dashing_dtree.pickle"), "rb").read()
    for name in (b"huffman_dandd", b"DeltaSpider", b"sketch_classes", b"DashSketchObj", b"SketchFilePath",
                 b"species_specifics", b"SpeciesSpecifics"):
        assert name in raw
    assert b"dandd_b200" not in raw
    tree = pickle.loads(raw)
    assert sorted(vars(tree)) == sorted(["_dt", "delta", "experiment", "fastas", "kstart", "maxk", "mink", "ngen", "root",
                                         "speciesinfo"])
    node = tree.root
    assert sorted(vars(node)) == sorted(["bestk", "card", "children", "delta", "experiment", "fastas", "ksketches", "maxk",
                                         "mink", "ngen", "node_title", "progeny", "speciesinfo"])
    assert len(node.ksketches) >= 100 and node.ksketches[0].kval == 0
    obj = node.ksketches[node.bestk]
    assert sorted(vars(obj)) == sorted(["kval", "sketch", "cmd", "sfp", "delta_pos", "card", "speciesinfo", "experiment",
                                        "_presketches"])
