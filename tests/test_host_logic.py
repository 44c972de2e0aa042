"""Host logic of the drop-in layer (dandd_b200/lib) against the reference's own outputs
(tests/golden/reference_runs.json), on the CPU: the store is replaced by the oracle-backed double
so that naming, caching, the k hill-climb, tree shapes and every output table are exercised in a
container without a GPU.  tests/test_host_gpu.py runs the same scenarios on the real store."""
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


def test_pickle_roundtrip(tmp_path, oracle_store):
    host_cases.scenario_pickle_roundtrip(str(tmp_path))
