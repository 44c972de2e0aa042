"""The drop-in scenarios of tests/host_cases.py on the real GPU store: `dandd tree / progressive /
kij / --exact` end to end through the C ABI, compared with the reference's own outputs."""
import pytest

from tests import host_cases

pytestmark = pytest.mark.gpu


@pytest.fixture()
def gpu_store():
    from dandd_b200 import build
    build.build()
    from dandd_b200 import store as ddstore
    st = ddstore.GpuSketchStore()
    ddstore.set_store(st)
    yield st
    ddstore.set_store(None)


def test_tree_hillclimb(tmp_path, gpu_store):
    host_cases.scenario_tree_hillclimb(str(tmp_path))
    assert gpu_store.stats["leaf_passes"] == 5      # one fused all-k pass per FASTA, whatever the hill-climb visits


def test_rerun_is_fully_cached(tmp_path, gpu_store):
    host_cases.scenario_rerun_is_fully_cached(str(tmp_path), gpu_store)


def test_ksweep_and_progressive(tmp_path, gpu_store):
    host_cases.scenario_ksweep_and_progressive(str(tmp_path))


def test_progressive_hillclimb_and_kij(tmp_path, gpu_store):
    host_cases.scenario_progressive_hillclimb_and_kij(str(tmp_path))


def test_tree_nchildren(tmp_path, gpu_store):
    host_cases.scenario_tree_nchildren(str(tmp_path))


def test_config1_tree(tmp_path, gpu_store):
    host_cases.scenario_config1_tree(str(tmp_path))


def test_tree_exact(tmp_path, gpu_store):
    host_cases.scenario_tree_exact(str(tmp_path))


def test_exact_sweep_above_k32(tmp_path, gpu_store):
    host_cases.scenario_exact_sweep_above_k32(str(tmp_path))


def test_option_coverage(tmp_path, gpu_store):
    host_cases.scenario_option_coverage(str(tmp_path))


def test_exact_sweep_progressive_and_binary_tree(tmp_path, gpu_store):
    host_cases.scenario_exact_sweep_progressive_and_binary_tree(str(tmp_path))


def test_default_sweep(tmp_path, gpu_store):
    host_cases.scenario_default_sweep(str(tmp_path))


def test_stub_union_files(tmp_path, gpu_store):
    """DANDD_B200_UNION_FILES=stub keeps union registers in HBM and writes marker files only; the
    reported numbers do not change."""
    gpu_store.union_files = "stub"
    host_cases.scenario_tree_hillclimb(str(tmp_path))


def test_tree_sweep_batched_unions_on_the_gpu_store(tmp_path, gpu_store):
    """`tree --ksweep --nchildren 2` through GpuSketchStore.union_many (one batched job for every inner
    node x k) == one union per node, full and stub union files alike."""
    from tests.test_host_logic import _tree_run
    from dandd_b200 import store as ddstore
    from dandd_b200.store import GpuSketchStore
    a = _tree_run(str(tmp_path / "batch"), True)
    batched = gpu_store.stats["union_launches"]
    st2 = GpuSketchStore(engine=gpu_store.engine, union_files="stub")
    ddstore.set_store(st2)
    try:
        b = _tree_run(str(tmp_path / "nodes"), False)
    finally:
        ddstore.set_store(gpu_store)
    strip = lambda x, tag: x.replace(tag + "/", "")     # noqa: E731
    assert [strip(f, "batch") for f in a["files"]] == [strip(f, "nodes") for f in b["files"]]
    assert a["deltas"] == b["deltas"]
    ca = {strip(k, "batch"): v for k, v in a["tb_dashing_cardinalities"].items()}
    cb = {strip(k, "nodes"): v for k, v in b["tb_dashing_cardinalities"].items()}
    assert ca == cb
    assert batched <= 3 and st2.stats["union_launches"] >= 6      # (batched: one launch per member-count group)
