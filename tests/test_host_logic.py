"""Host logic of the drop-in layer (dandd_b200/lib) against the reference's own outputs
(tests/golden/reference_runs.json), on the CPU: the store is replaced by the oracle-backed double
so that naming, caching, the k hill-climb, tree shapes and every output table are exercised in a
container without a GPU.  tests/test_host_gpu.py runs the same scenarios on the real store."""
import os

import pytest

from dandd_b200 import store as ddstore
from tests import host_cases
from tests.oracle_store import OracleStore


@pytest.fixture()
def oracle_store():
    st = OracleStore()
    ddstore.set_store(st)
    yield st
    ddstore.set_store(None)


def test_tree_hillclimb(tmp_path, oracle_store):
    host_cases.scenario_tree_hillclimb(str(tmp_path))


def test_rerun_is_fully_cached(tmp_path, oracle_store):
    host_cases.scenario_rerun_is_fully_cached(str(tmp_path), oracle_store)


def test_ksweep_and_progressive(tmp_path, oracle_store):
    host_cases.scenario_ksweep_and_progressive(str(tmp_path))


def test_progressive_hillclimb_and_kij(tmp_path, oracle_store):
    host_cases.scenario_progressive_hillclimb_and_kij(str(tmp_path))


def test_tree_nchildren(tmp_path, oracle_store):
    host_cases.scenario_tree_nchildren(str(tmp_path))


def test_config1_tree(tmp_path, oracle_store):
    host_cases.scenario_config1_tree(str(tmp_path))


def test_tree_exact(tmp_path, oracle_store):
    host_cases.scenario_tree_exact(str(tmp_path))


def test_exact_sweep_above_k32(tmp_path, oracle_store):
    host_cases.scenario_exact_sweep_above_k32(str(tmp_path))


def test_option_coverage(tmp_path, oracle_store):
    host_cases.scenario_option_coverage(str(tmp_path))


def test_exact_sweep_progressive_and_binary_tree(tmp_path, oracle_store):
    host_cases.scenario_exact_sweep_progressive_and_binary_tree(str(tmp_path))


def test_default_sweep(tmp_path, oracle_store):
    host_cases.scenario_default_sweep(str(tmp_path))


def test_pickle_roundtrip(tmp_path, oracle_store):
    host_cases.scenario_pickle_roundtrip(str(tmp_path))


def test_cached_rerun_never_touches_the_store(tmp_path, oracle_store):
    """A second identical `tree` run is answered from the sketch database alone: with no store
    installed (creating the real one would need CUDA and fail loudly here) it must still succeed --
    which is what lets such runs skip importing torch and starting CUDA altogether."""
    import sys
    from dandd_b200 import store as ddstore
    from tests.host_harness import run_dandd
    out = host_cases.scenario_tree_hillclimb(str(tmp_path))
    first = open(os.path.join(out, "runA_5_dashing_deltas.csv")).read()
    ddstore.set_store(None)
    try:
        run_dandd(["tree", "-d", os.path.join(str(tmp_path), "data5"), "-s", "runA", "-k", "14", "-o", out])
        assert ddstore._store is None
    finally:
        ddstore.set_store(oracle_store)
    assert open(os.path.join(out, "runA_5_dashing_deltas.csv")).read() == first


def test_cached_rerun_as_a_process_imports_no_torch(tmp_path, oracle_store):
    """The real launcher, as a subprocess, on a fully cached sketch database: exits 0 in a container
    without a GPU and never imports torch (that is the whole start-up cost of such a run)."""
    import subprocess
    import sys
    out = host_cases.scenario_tree_hillclimb(str(tmp_path))
    launcher = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dandd_b200", "lib", "dandd")
    argv = [launcher, "tree", "-d", os.path.join(str(tmp_path), "data5"), "-s", "runA", "-k", "14", "-o", out]
    code = ("import sys, runpy\nsys.argv = %r\ntry:\n    runpy.run_path(sys.argv[0], run_name='__main__')\n"
            "except SystemExit as e:\n    assert not e.code, e.code\nprint('TORCH_IMPORTED', 'torch' in sys.modules)\n" % (argv,))
    env = dict(os.environ, WORLD_SIZE="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "TORCH_IMPORTED False" in r.stdout


def _kij_run(tmp, table, jaccard):
    """tree + kij on 7 genomes; returns everything the run leaves behind, paths made relative."""
    import csv
    import pickle
    from tests.host_harness import run_dandd
    from tests.util import make_dataset
    data = make_dataset(os.path.join(tmp, "data"), 7, 6000, seed=31)
    out = os.path.join(tmp, "out")
    os.environ["DANDD_B200_PAIR_TABLE"] = "1" if table else "0"
    try:
        run_dandd(["tree", "-d", os.path.dirname(data[0]), "-s", "kj", "-k", "11", "-o", out, "-r", "10"])
        argv = ["kij", "-d", os.path.join(out, "kj_7_dashing_dtree.pickle"), "-o", out, "--mink", "8", "--maxk", "13", "--afproject"]
        run_dandd(argv + (["--jaccard"] if jaccard else []))
    finally:
        os.environ.pop("DANDD_B200_PAIR_TABLE", None)
    db = os.path.join(out, "sketchdb")
    rel = lambda p: os.path.relpath(p, tmp)     # noqa: E731
    res = {"kij": [{k: (rel(v) if k in ("A", "B") else v) for k, v in r.items()} for r in csv.DictReader(open(os.path.join(out, "kj_7_dashing.kij.csv")))],
           "files": sorted(rel(os.path.join(d, f)) for d, _, fs in os.walk(db) for f in fs if not f.endswith(".bkp"))}
    if jaccard:
        res["j"] = [{k: (rel(v) if k in ("A", "B") else v) for k, v in r.items()} for r in csv.DictReader(open(os.path.join(out, "kj_7_dashing.j.csv")))]
    for name in ("kj_dashing_cardinalities", "dandd_fastahex", "dandd_sketchinfo"):
        with open(os.path.join(db, name + ".pickle"), "rb") as fh:
            obj = pickle.load(fh)
        res[name] = {(rel(k) if os.path.isabs(k) else k): v for k, v in obj.items()}
    with open(os.path.join(out, "kj_7_dashing_AFtuples.pickle"), "rb") as fh:
        res["af"] = sorted((t[0], os.path.basename(str(t[1])), os.path.basename(str(t[2])), t[3], round(t[4], 9), t[5], t[6], t[7])
                           for t in pickle.load(fh))
    return res


@pytest.mark.parametrize("jaccard", [False, True])
def test_kij_from_the_batched_pair_table_equals_one_spider_per_pair(tmp_path, oracle_store, jaccard):
    """DeltaTree._pair_from_table replays SubSpider + find_delta + kij_summarize on the K6 table; rows,
    sketch files, cardinality cache, name registrations and AFproject tuples must equal what the
    per-pair object path (the reference's shape, :666-695) leaves behind."""
    a = _kij_run(str(tmp_path / "table"), True, jaccard)
    ddstore.set_store(OracleStore())
    b = _kij_run(str(tmp_path / "spiders"), False, jaccard)
    ra = {k: {kk: (vv.replace("table/", "") if isinstance(vv, str) else vv) for kk, vv in r.items()} if isinstance(r, dict) else r
          for k, r in enumerate(a["kij"])}
    rb = {k: {kk: (vv.replace("spiders/", "") if isinstance(vv, str) else vv) for kk, vv in r.items()} if isinstance(r, dict) else r
          for k, r in enumerate(b["kij"])}
    assert ra == rb
    strip = lambda x, tag: x.replace(tag + "/", "")     # noqa: E731
    assert [strip(f, "table") for f in a["files"]] == [strip(f, "spiders") for f in b["files"]]
    for name in ("kj_dashing_cardinalities", "dandd_fastahex", "dandd_sketchinfo"):
        assert {strip(k, "table"): v for k, v in a[name].items()} == {strip(k, "spiders"): v for k, v in b[name].items()}, name
    assert a["af"] == b["af"]
    if jaccard:
        assert [{k: (strip(v, "table") if isinstance(v, str) else v) for k, v in r.items()} for r in a["j"]] == \
               [{k: (strip(v, "spiders") if isinstance(v, str) else v) for k, v in r.items()} for r in b["j"]]


def test_plan_tree_matches_the_built_shape(tmp_path, oracle_store):
    """DeltaTree.plan_tree (sizes only) must predict exactly the parents _build_tree creates, for the
    odd n-ary shapes of the reference."""
    import dandd_b200
    dandd_b200.enable_compat()
    import huffman_dandd
    for n in (3, 4, 7, 10, 13):
        for c in (2, 3, 4):
            plan = huffman_dandd.DeltaTree.plan_tree([1] * n, c)
            # replay with the real loop on stand-in nodes
            class N:      # noqa: N801
                def __init__(self, leaves):
                    self.leaves, self.ngen = leaves, len(leaves)
            dt, made, cursor, insert_at, nch = [N([i]) for i in range(n)], [], 0, 0, c
            while cursor != len(dt) - 1:
                stride = nch - 1
                group = dt[cursor:cursor + nch]
                parent = N([x for g in group for x in g.leaves])
                made.append(parent.leaves)
                while insert_at < len(dt) - stride and dt[insert_at + stride].ngen <= parent.ngen:
                    insert_at += stride
                dt.insert(insert_at + stride, parent)
                cursor += nch
                if insert_at + stride > len(dt) - 1:
                    nch = len(dt) - cursor
            assert plan == made, (n, c)


def _tree_run(tmp, batch):
    import csv
    import pickle
    from tests.host_harness import run_dandd
    from tests.util import make_dataset
    data = make_dataset(os.path.join(tmp, "data"), 7, 5000, seed=41)
    out = os.path.join(tmp, "out")
    os.environ["DANDD_B200_TREE_BATCH"] = "1" if batch else "0"
    try:
        run_dandd(["tree", "-d", os.path.dirname(data[0]), "-s", "tb", "-o", out, "-r", "10", "--nchildren", "2", "--ksweep",
                   "--mink", "9", "--maxk", "13"])
    finally:
        os.environ.pop("DANDD_B200_TREE_BATCH", None)
    db = os.path.join(out, "sketchdb")
    rel = lambda p: os.path.relpath(p, tmp)     # noqa: E731
    res = {"files": sorted(rel(os.path.join(d, f)) for d, _, fs in os.walk(db) for f in fs if not f.endswith(".bkp")),
           "deltas": [{k: (v if k not in ("sketchloc", "fastas") else os.path.basename(v)) for k, v in r.items()}
                      for r in csv.DictReader(open(os.path.join(out, "tb_7_dashing_deltas.csv")))]}
    for name in ("tb_dashing_cardinalities", "dandd_fastahex", "dandd_sketchinfo"):
        with open(os.path.join(db, name + ".pickle"), "rb") as fh:
            res[name] = {(rel(k) if os.path.isabs(k) else k): v for k, v in pickle.load(fh).items()}
    return res


def test_tree_sweep_batched_unions_equal_per_node_unions(tmp_path, oracle_store):
    """`tree --ksweep --nchildren 2`: all inner nodes x all k through ONE store.union_many call must leave
    the same database, cardinalities and deltas as one union_sketches call per node."""
    a = _tree_run(str(tmp_path / "batch"), True)
    launches_batched = oracle_store.stats["union_launches"]
    st2 = OracleStore()
    ddstore.set_store(st2)
    b = _tree_run(str(tmp_path / "nodes"), False)
    strip = lambda x, tag: x.replace(tag + "/", "")     # noqa: E731
    assert [strip(f, "batch") for f in a["files"]] == [strip(f, "nodes") for f in b["files"]]
    assert a["deltas"] == b["deltas"]
    for name in ("tb_dashing_cardinalities", "dandd_fastahex", "dandd_sketchinfo"):
        assert {strip(k, "batch"): v for k, v in a[name].items()} == {strip(k, "nodes"): v for k, v in b[name].items()}, name
    assert launches_batched == 1 and st2.stats["union_launches"] >= 3


def test_large_fresh_leaves_are_sketched_before_they_are_named(tmp_path, oracle_store, monkeypatch):
    """Single-process `tree`: FASTAs above the streaming threshold that the database has never seen get
    their all-k device pass BEFORE any blake2b name is asked for (the hashing runs in background
    threads meanwhile); a cached re-run warms nothing and never creates the store; the outputs are
    those of the plain path."""
    from dandd_b200 import ingest, store as ddstore
    from tests.host_harness import run_dandd
    from tests.util import make_dataset
    data = str(tmp_path / "data5")
    make_dataset(data, 5, 20000, seed=21)
    events = []
    real_digest = ingest.digest

    def spy_digest(path):
        events.append(("digest", os.path.basename(path)))
        return real_digest(path)

    def warm_leaf(path, p, canon):
        events.append(("warm", os.path.basename(path)))

    monkeypatch.setattr(ingest, "digest", spy_digest)
    oracle_store.warm_leaf = warm_leaf
    oracle_store.cache_bytes = 64 << 30
    argv = ["tree", "-d", data, "-s", "runW", "-k", "14", "-o"]
    run_dandd(argv + [str(tmp_path / "out_plain")])                  # 20 kbp files: below the threshold
    assert not [e for e in events if e[0] == "warm"]
    plain = open(tmp_path / "out_plain" / "runW_5_dashing_deltas.csv").read()

    events.clear()
    monkeypatch.setattr(ingest, "HASH_ONLY_MIN_BYTES", 1000)
    run_dandd(argv + [str(tmp_path / "out_warm")])
    warms = [i for i, e in enumerate(events) if e[0] == "warm"]
    digests = [i for i, e in enumerate(events) if e[0] == "digest"]
    assert len(warms) == 5 and digests and max(warms) < min(digests)
    assert sorted(e[1] for e in events if e[0] == "warm") == sorted(os.listdir(data))
    assert open(tmp_path / "out_warm" / "runW_5_dashing_deltas.csv").read().replace("out_warm", "out_plain") == plain

    events.clear()
    ddstore.set_store(None)
    try:
        run_dandd(argv + [str(tmp_path / "out_warm")])               # cached: nothing to warm, no store
        assert ddstore._store is None and not [e for e in events if e[0] == "warm"]
    finally:
        ddstore.set_store(oracle_store)
    oracle_store.cache_bytes = 0                                     # no room to keep blocks: nothing is warmed
    run_dandd(argv + [str(tmp_path / "out_nofit")])
    assert not [e for e in events if e[0] == "warm"]


def test_gpus_flag_starts_workers_only_for_fresh_fastas(tmp_path, oracle_store, monkeypatch):
    """`tree --gpus N`: ranks 1..N-1 are started (as copies of the command line, with the rendezvous
    environment) when some FASTA has never been named by the sketch database -- and not at all on a
    re-run over the same files, which then stays a single process that never imports torch."""
    import argparse
    import subprocess
    import dandd_b200
    from tests.host_harness import run_dandd
    from tests.util import make_dataset
    dandd_b200.enable_compat()
    import dandd_cmd
    data = str(tmp_path / "data5")
    make_dataset(data, 5, 20000, seed=21)
    out = str(tmp_path / "out")
    started = []

    class FakePopen:
        def __init__(self, argv, env=None):
            started.append((argv, env))
            self.args = argv

        def wait(self):
            return 0

    monkeypatch.setattr(subprocess, "Popen", FakePopen)
    saved = dict(os.environ)
    fresh_bytes = sum(os.path.getsize(os.path.join(data, f)) for f in os.listdir(data))
    try:
        for key in ("WORLD_SIZE", "RANK", "LOCAL_RANK", "CUDA_VISIBLE_DEVICES"):
            os.environ.pop(key, None)
        args = argparse.Namespace(gpus=3, exact=False, genomedir=data, flist_loc=None, sketchdir=None, outdir=out)
        # N is an upper bound: one process per DANDD_B200_BYTES_PER_GPU of fresh FASTA
        os.environ.pop("DANDD_B200_BYTES_PER_GPU", None)
        assert dandd_cmd._self_launch(args) == [] and not started          # 100 kB is far below 12 GiB
        os.environ["DANDD_B200_BYTES_PER_GPU"] = str(fresh_bytes // 2)
        assert len(dandd_cmd._self_launch(args)) == 1                      # two processes' worth
        assert started[0][1]["WORLD_SIZE"] == "2"
    finally:
        os.environ.clear()
        os.environ.update(saved)
    started.clear()
    saved["DANDD_B200_BYTES_PER_GPU"] = "0"                                # from here on: all N whenever anything is fresh
    os.environ["DANDD_B200_BYTES_PER_GPU"] = "0"
    try:
        for key in ("WORLD_SIZE", "RANK", "LOCAL_RANK", "CUDA_VISIBLE_DEVICES"):
            os.environ.pop(key, None)
        children = dandd_cmd._self_launch(args)
        assert len(children) == 2 and len(started) == 2
        envs = [env for _, env in started]
        assert sorted(e["RANK"] for e in envs) == ["1", "2"] and all(e["WORLD_SIZE"] == "3" for e in envs)
        assert all(e["MASTER_ADDR"] == "127.0.0.1" and e["MASTER_PORT"] == os.environ["MASTER_PORT"] for e in envs)
        assert os.environ["RANK"] == "0" and os.environ["WORLD_SIZE"] == "3"
    finally:
        os.environ.clear()
        os.environ.update(saved)
    run_dandd(["tree", "-d", data, "-s", "runG", "-k", "14", "-o", out])          # names every FASTA
    started.clear()
    try:
        for key in ("WORLD_SIZE", "RANK", "LOCAL_RANK", "CUDA_VISIBLE_DEVICES"):
            os.environ.pop(key, None)
        assert dandd_cmd._self_launch(args) == [] and not started and "WORLD_SIZE" not in os.environ
        args.exact = True                      # exact counting shards over ranks whatever the names say
        assert len(dandd_cmd._self_launch(args)) == 2
    finally:
        os.environ.clear()
        os.environ.update(saved)
    # one more FASTA: fresh again
    with open(os.path.join(data, "zz_new.fa"), "wb") as fh:
        fh.write(b">n\\nACGTACGTAC\\n")
    started.clear()
    try:
        for key in ("WORLD_SIZE", "RANK", "LOCAL_RANK", "CUDA_VISIBLE_DEVICES"):
            os.environ.pop(key, None)
        args.exact = False
        assert len(dandd_cmd._self_launch(args)) == 2
    finally:
        os.environ.clear()
        os.environ.update(saved)
        os.environ.pop("DANDD_B200_BYTES_PER_GPU", None)


@pytest.mark.parametrize("jaccard", [False, True])
def test_kij_rerun_is_answered_by_the_database(tmp_path, oracle_store, jaccard):
    """A second `kij` over the same tree finds every pair union on record (file + cardinality) and must
    not create the store -- no batched device job, no torch import -- while writing the same tables.
    Deleting ONE union file brings the batched job back (for that cell the database has no answer)."""
    import csv
    from tests.host_harness import run_dandd
    first = _kij_run(str(tmp_path), True, jaccard)
    out = os.path.join(str(tmp_path), "out")
    argv = ["kij", "-d", os.path.join(out, "kj_7_dashing_dtree.pickle"), "-o", out, "--mink", "8", "--maxk", "13", "--afproject"]
    argv += ["--jaccard"] if jaccard else []

    def rows():
        return [{k: v for k, v in r.items() if k not in ("A", "B")} for r in csv.DictReader(open(os.path.join(out, "kj_7_dashing.kij.csv")))]
    want = rows()

    def trap(*a, **k):
        raise AssertionError("a fully cached kij re-run created the sketch store")
    real_cls = ddstore.GpuSketchStore
    ddstore.set_store(None)
    ddstore.GpuSketchStore = trap
    try:
        run_dandd(argv)
        assert ddstore._store is None
    finally:
        ddstore.GpuSketchStore = real_cls
        ddstore.set_store(oracle_store)
    assert rows() == want and len(want) == 21
    # one union sketch goes missing: the batched job runs again and restores it
    db = os.path.join(out, "sketchdb", "ngen2")
    victim = sorted(os.path.join(d, f) for d, _, fs in os.walk(db) for f in fs if f.endswith(".hll"))[3]
    os.remove(victim)
    calls = {"n": 0}
    real_pairs = oracle_store.pair_unions

    def counting(*a, **k):
        calls["n"] += 1
        return real_pairs(*a, **k)
    oracle_store.pair_unions = counting
    run_dandd(argv)
    assert calls["n"] == 1 and os.path.getsize(victim) > 0 and rows() == want
    assert len(first["kij"]) == 21


@pytest.mark.parametrize("sweep", [False, True])
@pytest.mark.parametrize("n,nchildren", [(3, 4), (4, 3), (6, 4)])
def test_nary_shapes_the_reference_cannot_build_fail_the_same_way(tmp_path, oracle_store, n, nchildren, sweep):
    """For some (genomes, --nchildren) the reference's parent-insertion cursor runs off the node list and it
    dies naming a parent without members: FileNotFoundError on '' (lib/huffman_dandd.py:412-438,
    lib/sketch_classes.py:12-18; found by tests/test_reference_live.py).  The drop-in must fail the
    same way -- also with --ksweep, where the batched tree-union job would otherwise be handed an empty set."""
    from tests.host_harness import run_dandd
    from tests.util import make_dataset
    data = str(tmp_path / "d")
    make_dataset(data, n, 1500, seed=7 + n, sub=0.05)
    argv = ["tree", "-d", data, "-s", "x", "-k", "10", "-r", "10", "-n", str(nchildren), "-o", str(tmp_path / "out")]
    with pytest.raises(FileNotFoundError) as err:
        run_dandd(argv + (["--ksweep", "--mink", "9", "--maxk", "11"] if sweep else []))
    assert "''" in str(err.value)


def test_store_cache_evicts_a_leaf_block_together_with_its_views():
    """GpuSketchStore's byte budget: per-k sketches that are views of an all-k leaf block cost nothing
    by themselves, so when the block is evicted they must go too (they would keep its memory alive,
    uncounted); tensors a caller still holds stay valid.  Cache plumbing only: no device is touched."""
    import torch
    st = ddstore.GpuSketchStore(engine=object(), cache_bytes=100)
    blocks = {}
    for name in ("a", "b", "c"):
        regs = torch.full((4, 10), ord(name), dtype=torch.uint8)          # 40 "bytes" per block
        ent = {"regs": regs, "cards": [0.0] * 4, "ks": {k: k for k in range(4)}, "views": set()}
        blocks[name] = ent
        st._put(("leaf", name, 10, True), ent, regs.numel())
        for k in range(4):
            ent["views"].add(f"/db/{name}.{k}")
            st._remember(f"/db/{name}.{k}", regs[k], view_of_block=True)
    # 3 x 40 > 100: block a went, and with it its four views; b and c are whole
    assert st._get(("leaf", "a", 10, True)) is None and all(st._get(("sketch", f"/db/a.{k}")) is None for k in range(4))
    assert all(st._get(("sketch", f"/db/{n}.{k}")) is not None for n in "bc" for k in range(4))
    assert st._bytes == 80 and sum(st._cost.values()) == 80 and set(st._cost) == set(st._lru)
    assert int(blocks["a"]["regs"][2, 0]) == ord("a")                      # a held tensor is untouched
    # a standalone sketch (read from a file, or a union) is charged its own size and evicts the oldest block
    st._remember("/db/u", torch.zeros(30, dtype=torch.uint8))
    assert st._get(("leaf", "b", 10, True)) is None and st._get(("sketch", "/db/b.0")) is None
    assert st._get(("sketch", "/db/c.3")) is not None and st._bytes == 70
    st.forget("/db/u")
    assert st._bytes == 40


def test_info_command(tmp_path, oracle_store, capsys):
    """`dandd info`: same flags as the reference's handler-less sub-parser; prints the delta table of a saved
    tree and, with --ksweep, writes one row per (node, k) after sketching whatever the range still lacks."""
    import csv
    import dandd_b200
    from oracle import pyoracle as orc
    from tests.host_harness import run_dandd
    from tests.util import make_dataset
    files = make_dataset(str(tmp_path / "data"), 3, 4000, seed=31)
    out = str(tmp_path / "out")
    run_dandd(["tree", "-d", str(tmp_path / "data"), "-s", "inf", "-k", "11", "-o", out, "-r", "10"])
    pickle_path = os.path.join(out, "inf_3_dashing_dtree.pickle")
    capsys.readouterr()
    run_dandd(["info", "-d", pickle_path, "-o", out])
    shown = capsys.readouterr().out.splitlines()
    with open(os.path.join(out, "inf_3_dashing_deltas.csv")) as fh:
        assert shown == fh.read().splitlines()                      # the table `tree` saved, nothing else
    assert not os.path.exists(os.path.join(out, "inf_3_dashing_info.csv"))
    passes = oracle_store.stats["leaf_passes"]
    run_dandd(["info", "-d", pickle_path, "-o", out, "-l", "wide", "--ksweep", "--mink", "8", "--maxk", "40"])
    with open(os.path.join(out, "inf_wide_3_dashing_info.csv"), newline="") as fh:
        rows = list(csv.DictReader(fh))
    assert len(rows) == 4 * (32 - 8 + 1)                            # 3 leaves + the root, k = 8..32 (Dashing's limit)
    assert oracle_store.stats["leaf_passes"] > passes               # the range was wider than what the hill-climb had visited
    sym = orc.fasta_symbols(open(files[1], "rb").read())
    for r in rows:
        assert float(r["delta_pos"]) == pytest.approx(float(r["card"]) / int(r["kval"]), rel=1e-12)
        if r["title"] == "g1" and r["kval"] in ("8", "21", "32"):
            assert float(r["card"]) == pytest.approx(orc.card(orc.hll_sketch(sym, int(r["kval"]), 10), 10), rel=1e-9)
    parser, commands = __import__("dandd_cmd").parse_arguments()
    assert commands == ["tree", "progressive", "kij"]
    with pytest.raises(SystemExit):
        parser.parse_args(["info"])                                  # -d is required, as in the reference
    assert dandd_b200.LIB_DIR


def test_lowmem_union_over_a_trusted_child_whose_file_is_gone(tmp_path, oracle_store):
    """--lowmem trusts the cardinality on record of a multi-FASTA sketch whose file was deleted and does not
    rebuild it; a parent union at a k the parent has never seen must then be built from that child's own
    members (the leaves), not from the missing file.  (Found by tests/test_reference_live.py: the store
    raised FileNotFoundError.)  A hill-climbed binary tree gives such cells: every node visits its own ks."""
    import pickle
    import numpy as np
    from dandd_b200 import hllfile
    from oracle import pyoracle as orc
    from tests.host_harness import run_dandd
    from tests.util import make_dataset
    files = make_dataset(str(tmp_path / "data"), 5, 4000, seed=77, sub=0.2)
    out = str(tmp_path / "out")
    base = ["tree", "-d", str(tmp_path / "data"), "-s", "lm", "-k", "9", "-o", out, "-r", "10", "-n", "2"]
    run_dandd(base)
    db = os.path.join(out, "sketchdb")
    with open(os.path.join(db, "lm_dashing_cardinalities.pickle"), "rb") as fh:
        before = pickle.load(fh)
    # one inner node of that tree, swept on its own over ks the whole tree never visited: its unions are now on
    # record (and on disk) at k = 4..6, the root's are not
    from tests.host_harness import read_csv
    inner = [r for r in read_csv(os.path.join(out, "lm_5_dashing_deltas.csv")) if int(r["ngen"]) == 2][0]
    flist = str(tmp_path / "inner.txt")
    with open(flist, "w") as fh:
        fh.write("\n".join(inner["fastas"].split("|")) + "\n")
    run_dandd(["tree", "-f", flist, "-s", "lm", "-k", "5", "-o", str(tmp_path / "out2"), "-r", "10", "-c", os.path.join(out, "sketchdb"),
               "--ksweep", "--mink", "4", "--maxk", "6"])
    with open(os.path.join(db, "lm_dashing_cardinalities.pickle"), "rb") as fh:
        before = pickle.load(fh)
    unions = [p for p in before if os.sep + "ngen1" + os.sep not in p]
    assert any(os.sep + "ngen2" + os.sep + "k4" + os.sep in p for p in unions) and not any(os.sep + "ngen5" + os.sep + "k4" + os.sep in p for p in unions)
    for p in unions:
        os.remove(p)
    oracle_store._regs.clear()
    run_dandd(base + ["--ksweep", "--mink", "4", "--maxk", "14", "--lowmem"])
    syms = [orc.fasta_symbols(open(f, "rb").read()) for f in files]
    with open(os.path.join(db, "lm_dashing_cardinalities.pickle"), "rb") as fh:
        cardkey = pickle.load(fh)
    checked = 0
    for k in range(4, 15):                                      # the root at every k: rebuilt from whatever was at hand, or trusted
        directory = os.path.join(db, "ngen5", f"k{k}")
        paths = [os.path.join(directory, f) for f in (os.listdir(directory) if os.path.isdir(directory) else [])]
        want = orc.union_max([orc.hll_sketch(s, k, 10) for s in syms])
        if paths:
            assert np.array_equal(hllfile.read_hll(paths[0])[0], want), k
            assert cardkey[paths[0]] == pytest.approx(orc.card(want, 10), rel=1e-9)
            checked += 1
        else:
            trusted = [p for p in before if os.sep + "ngen5" + os.sep + f"k{k}" + os.sep in p]
            assert trusted and cardkey[trusted[0]] == before[trusted[0]] > 0
    assert checked >= 5


def test_nchildren_one_is_refused(tmp_path, oracle_store):
    """`--nchildren 1` over several FASTAs: the reference loops forever (each new node has one child, the node
    list never shrinks; verified with a 25 s timeout in the build container); the drop-in refuses."""
    from tests.host_harness import run_dandd
    from tests.util import make_dataset
    make_dataset(str(tmp_path / "data"), 3, 2000, seed=5)
    with pytest.raises(ValueError, match="--nchildren 1"):
        run_dandd(["tree", "-d", str(tmp_path / "data"), "-s", "t", "-k", "11", "-o", str(tmp_path / "out"), "-r", "10", "-n", "1"])
