"""Live differential runs against the UNMODIFIED reference Python, on randomly drawn scenarios.

tests/golden/reference_runs.json pins twenty fixed scenarios; here the same comparison is made on
inputs and option combinations drawn from a seeded generator, so that host-logic parity (naming,
caching, the k hill-climb, tree shapes, progressive and KIJ tables) does not rest on the fixtures
alone.  The reference runs with the oracle-backed dashing/kmc/parallel stand-ins first on PATH
(oracle/shims), the drop-in layer with the oracle-backed store double -- so both sides see the same
arithmetic and any difference is host logic.

Only where /root/reference exists (the build container); skipped elsewhere.  CPU only."""
import importlib.util
import os
import random

import pytest

from dandd_b200 import store as ddstore
from tests.host_harness import assert_tree_matches, close, collect_tree, read_csv, run_dandd
from tests.oracle_store import OracleStore
from tests.util import make_dataset

REF = "/root/reference/lib"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is not on this machine")


@pytest.fixture(scope="module")
def golden():
    """tests/golden/make_reference_golden.py as a module (run_ref, collect_tree) + the shims on a PATH dir."""
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_reference_golden", os.path.join(here, "golden", "make_reference_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture()
def oracle_store():
    st = OracleStore()
    ddstore.set_store(st)
    yield st
    ddstore.set_store(None)


def _draw(seed):
    rng = random.Random(seed)
    n = rng.randint(3, 8)
    case = {"n": n, "length": rng.choice([3000, 8000, 15000]), "seed": 100 + seed, "kstart": rng.randint(9, 15),
            "sub": rng.choice([0.01, 0.05, 0.15]), "registers": rng.choice([10, 12, 14]), "canon": rng.random() < 0.75,
            "nchildren": rng.choice([None, None, 2, 3, 4]), "sweep": None}
    if rng.random() < 0.5:
        lo = rng.randint(6, 12)
        case["sweep"] = (lo, lo + rng.randint(2, 5))
    return case


def _tree_argv(case, data, out, tag):
    argv = ["tree", "-d", data, "-s", tag, "-k", str(case["kstart"]), "-o", out, "-r", str(case["registers"])]
    if not case["canon"]:
        argv.append("-C")
    if case["nchildren"]:
        argv += ["-n", str(case["nchildren"])]
    if case["sweep"]:
        argv += ["--ksweep", "--mink", str(case["sweep"][0]), "--maxk", str(case["sweep"][1])]
    return argv


def _both(golden, bindir, ref_argv, our_argv):
    """Run one command on both sides.  True if both succeeded; False if the reference failed and the
    drop-in failed with the same exception (error behaviour is part of the interface: several option
    combinations crash the reference as shipped, e.g. n-ary shapes whose cursor runs off the node
    list, or `kij` hill-climbing out of the k range a sweep-built tree holds)."""
    import subprocess
    try:
        golden.run_ref(bindir, ref_argv)
    except subprocess.CalledProcessError as failed:
        last = failed.stderr.decode(errors="replace").strip().split("\n")[-1]
        with pytest.raises(Exception) as ours_err:
            run_dandd(our_argv)
        assert type(ours_err.value).__name__ in last, (last, repr(ours_err.value))
        return False
    run_dandd(our_argv)
    return True


@pytest.mark.parametrize("seed", range(int(os.environ.get("DANDD_LIVE_SEEDS", "8"))))
def test_random_tree_progressive_kij_match_the_reference(tmp_path, golden, oracle_store, seed):
    from oracle import pyoracle
    case = _draw(seed)
    bindir = pyoracle.install_shims(str(tmp_path / "bin"))
    data = str(tmp_path / "data")
    make_dataset(data, case["n"], case["length"], seed=case["seed"], sub=case["sub"])
    tag = f"r{seed}"
    prefix = f"{tag}_{case['n']}_dashing"
    ref_out, our_out = str(tmp_path / "ref"), str(tmp_path / "ours")
    # ---- tree
    if not _both(golden, bindir, _tree_argv(case, data, ref_out, tag), _tree_argv(case, data, our_out, tag)):
        print(f"LIVE seed {seed}: both fail at tree {case}")
        return
    want = golden.collect_tree(ref_out, prefix, os.path.join(ref_out, "sketchdb"), "dashing")
    ours = collect_tree(our_out, prefix, os.path.join(our_out, "sketchdb"), "dashing")
    assert_tree_matches(ours, want)
    # ---- progressive, identity ordering (the reference draws random orderings for -n > 1)
    sweep = ["--ksweep", "--mink", str(case["sweep"][0]), "--maxk", str(case["sweep"][1])] if case["sweep"] else []
    prog = lambda out: ["progressive", "-d", os.path.join(out, prefix + "_dtree.pickle"), "-n", "1", "-o", out] + sweep   # noqa: E731
    if not _both(golden, bindir, prog(ref_out), prog(our_out)):
        print(f"LIVE seed {seed}: both fail at progressive {case}")
        return
    name = f"{tag}_progu1_{case['n']}_dashing"
    def prog_rows(path):      # with --ksweep the reference leaves `delta` empty (SubSpider.delta stays None)
        from tests.host_harness import norm_fastas
        return [{"ngen": int(r["ngen"]), "kval": int(r["kval"]), "delta": float(r["delta"]) if r["delta"] else None,
                 "fastas": norm_fastas(r["fastas"], ",")} for r in read_csv(path)]
    rows_ref, rows_our = prog_rows(os.path.join(ref_out, name + ".csv")), prog_rows(os.path.join(our_out, name + ".csv"))
    assert [(r["ngen"], r["kval"], r["fastas"]) for r in rows_our] == [(r["ngen"], r["kval"], r["fastas"]) for r in rows_ref]
    assert all((a["delta"] is None and b["delta"] is None) or close(a["delta"], b["delta"]) for a, b in zip(rows_our, rows_ref))
    summ_ref = read_csv(os.path.join(ref_out, name + "summary.csv"))
    summ_our = read_csv(os.path.join(our_out, name + "summary.csv"))
    key = lambda r: (int(r["ngen"]), int(r["kval"]), r["title"])     # noqa: E731
    assert sorted(map(key, summ_our)) == sorted(map(key, summ_ref))
    cards_ref = {key(r): float(r["card"]) for r in summ_ref}
    assert all(close(float(r["card"]), cards_ref[key(r)]) for r in summ_our)
    # ---- kij (+ per-k Jaccard when a k range is known), batched pair table on our side
    jac = ["--jaccard", "--mink", str(case["sweep"][0]), "--maxk", str(case["sweep"][1])] if case["sweep"] else []
    kij = lambda out: ["kij", "-d", os.path.join(out, prefix + "_dtree.pickle"), "-o", out] + jac   # noqa: E731
    if not _both(golden, bindir, kij(ref_out), kij(our_out)):
        print(f"LIVE seed {seed}: both fail at kij {case}")
        return
    kij_ref = read_csv(os.path.join(ref_out, prefix + ".kij.csv"))
    kij_our = read_csv(os.path.join(our_out, prefix + ".kij.csv"))
    pair = lambda r: (r["Atitle"], r["Btitle"])     # noqa: E731
    assert sorted(map(pair, kij_our)) == sorted(map(pair, kij_ref)) and len(kij_ref) == case["n"] * (case["n"] - 1) // 2
    by_pair = {pair(r): r for r in kij_ref}
    for r in kij_our:
        g = by_pair[pair(r)]
        assert (int(r["Ak"]), int(r["Bk"]), int(r["ABk"])) == (int(g["Ak"]), int(g["Bk"]), int(g["ABk"])), pair(r)
        for col in ("Adelta", "Bdelta", "ABdelta"):
            assert close(float(r[col]), float(g[col])), (pair(r), col)
        assert float(r["KIJ"]) == pytest.approx(float(g["KIJ"]), rel=1e-6, abs=1e-6)
    if jac:
        j_ref = {(r["Atitle"], r["Btitle"], int(r["kval"])): r for r in read_csv(os.path.join(ref_out, prefix + ".j.csv"))}
        j_our = read_csv(os.path.join(our_out, prefix + ".j.csv"))
        assert sorted((r["Atitle"], r["Btitle"], int(r["kval"])) for r in j_our) == sorted(j_ref)
        for r in j_our:
            g = j_ref[(r["Atitle"], r["Btitle"], int(r["kval"]))]
            assert close(float(r["ABcard"]), float(g["ABcard"]))
            assert float(r["jaccard"]) == pytest.approx(float(g["jaccard"]), rel=1e-6, abs=1e-6)
    # ---- what the whole session left in the sketch database
    want_db = golden.collect_tree(ref_out, prefix, os.path.join(ref_out, "sketchdb"), "dashing")
    ours_db = collect_tree(our_out, prefix, os.path.join(our_out, "sketchdb"), "dashing")
    assert ours_db["files"] == want_db["files"]
    assert sorted(ours_db["cardkey"]) == sorted(want_db["cardkey"])
    assert all(close(ours_db["cardkey"][k], v) for k, v in want_db["cardkey"].items())
    assert ours_db["fastahex"] == want_db["fastahex"] and ours_db["sketchinfo"] == want_db["sketchinfo"]
    print(f"LIVE seed {seed}: tree + progressive + kij compared {case}")
