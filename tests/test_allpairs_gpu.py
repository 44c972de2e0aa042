"""dandd_b200/helpers/allpairs.py on the real GPU store: one fused all-k pass per FASTA (K1 + K2), one
batched pair job (K6; bit planes at p = 12, bytes at p = 10), exact pair unions (K5) -- compared with
what the reference's helpers/allpairs.py functions produced (tests/golden/allpairs_golden.json) and
with the oracle's exact counts."""
import pytest

from tests import allpairs_cases as cases

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


def _oracle_counts(fastas, k):
    from oracle import pyoracle as orc
    return orc.exact_count([orc.fasta_symbols(open(f, "rb").read()) for f in fastas], k, True)


@pytest.mark.parametrize("name", sorted(cases.gold_cases()))
def test_tables_match_the_reference_functions(tmp_path, gpu_store, name):
    table = cases.scenario_gold(str(tmp_path), name)
    assert gpu_store.stats["leaf_passes"] == len(table.names)      # every FASTA packed and sketched once for all k
    assert gpu_store.stats["files_written"] == 0                   # `dashing hll` leaves no sketch behind


def test_exact_tool(tmp_path, gpu_store):
    cases.scenario_exact(str(tmp_path), _oracle_counts)


def test_registers_of_a_block_equal_the_oracle(tmp_path, gpu_store):
    """leaf_block rows follow the requested k order and are bit-identical to the oracle's registers."""
    import numpy as np
    from oracle import pyoracle as orc
    case = cases.gold_cases()["five_nocanon_unsorted_klist"]
    inputs, _ = cases.write_dataset(str(tmp_path), case)
    ks = [21, 9, 15]
    regs, cards = gpu_store.leaf_block(inputs[2], ks, 10, False)
    sym = orc.fasta_symbols(open(inputs[2], "rb").read())
    for row, k in enumerate(ks):
        want = orc.hll_sketch(sym, k, 10, False)
        assert np.array_equal(regs[row].cpu().numpy(), want)
        assert cards[row] == pytest.approx(orc.card(want, 10), rel=1e-9)
