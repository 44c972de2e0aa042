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
