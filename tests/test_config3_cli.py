"""Config 3 through the drop-in command line, scaled down (8 x 50 Mbp instead of 8 x 3.1 Gbp):
`dandd tree --ksweep` (k = 2..32) + `dandd progressive -n 1 --ksweep` on FASTA files on disk, the
files large enough (> 32 MiB) to take the streaming ingest path; the sketch files the run leaves
behind are compared with the oracle (registers bit-exact, cardinalities 1e-9) and the argmax k of a
leaf and of a prefix union with the oracle's.  tools/config3_cli.py is the same driver that
produced profiles/r02_config3_cli*.json at full size."""
import argparse
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _args(tmp_path, gpus, **kw):
    base = dict(gpus=gpus, genomes=8, bases=50e6, kmin=2, kmax=32, workdir=str(tmp_path / "cfg3"), union_files="full",
                cpu_sample_bytes=0, oracle_ks="2,18,32", keep=True, out=None, _generate=False, launcher="self",
                also_torchrun=False)
    base.update(kw)
    return argparse.Namespace(**base)


def _check(rep, args):
    from dandd_b200 import hllfile
    from oracle import pyoracle as orc
    assert rep["oracle_check"]["all_registers_equal"]
    assert rep["oracle_check"]["max_card_rel_err"] <= 1e-9
    assert len(rep["oracle_check"]["cells"]) == 8 * 3
    # the streaming path was taken (files are 50 MB) and its stage times were recorded
    stages = rep["tree_stages"]
    assert stages["stream_wall_s"]["max"] > 0 and stages["stream_gpu_span_s"]["max"] > 0
    assert "prefetch_blake2b" in stages or "stream_blake2b_s" in stages      # the name hash ran beside the sketching
    # argmax k / delta of leaf 0 and of the prefix {0, 1} against the oracle over the whole sweep
    ks = list(range(2, 33))
    data = os.path.join(args.workdir, "fasta")
    syms = [orc.fasta_symbols(open(os.path.join(data, f"genome{g}.fa"), "rb").read()) for g in (0, 1)]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max(1, min(31, (os.cpu_count() or 2) - 1))) as ex:
        regs0 = list(ex.map(lambda k: orc.hll_sketch(syms[0], k, 20), ks))
        regs1 = list(ex.map(lambda k: orc.hll_sketch(syms[1], k, 20), ks))
    d0 = [orc.card(r, 20) / k for r, k in zip(regs0, ks)]
    d01 = [orc.card(np.maximum(a, b), 20) / k for a, b, k in zip(regs0, regs1, ks)]
    assert rep["leaf_argmax_k"][0] == ks[int(np.argmax(d0))]
    assert rep["leaf_delta"][0] == pytest.approx(max(d0), rel=1e-6)        # north-star: delta within 1e-6
    assert rep["prefix_argmax_k"][1] == ks[int(np.argmax(d01))]
    assert rep["prefix_delta"][1] == pytest.approx(max(d01), rel=1e-6)
    assert rep["prefix_delta"][0] == pytest.approx(rep["leaf_delta"][0], rel=1e-12)
    # every prefix union exists as a file DandD can find again
    assert rep["sketch_files"] >= 8 * 31 + 7 * 31
    assert rep["tree_cached_rerun_wall_s"] < rep["tree_wall_s"]
    del hllfile


def test_config3_cli_single_gpu(tmp_path):
    from tools import config3_cli
    args = _args(tmp_path, gpus=1)
    rep = config3_cli.run(args)
    _check(rep, args)


def test_config3_cli_two_ranks_same_results(tmp_path):
    """Two ranks (one GPU each) must leave the same cardinalities behind as one."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from tools import config3_cli
    one = config3_cli.run(_args(tmp_path / "a", gpus=1, oracle_ks=""))
    two = config3_cli.run(_args(tmp_path / "b", gpus=2, oracle_ks="", also_torchrun=True))   # --gpus 2 and torchrun
    assert one["leaf_cards"] == two["leaf_cards"] and two["torchrun_cards_identical"]
    assert one["prefix_delta"] == two["prefix_delta"] and one["prefix_argmax_k"] == two["prefix_argmax_k"]
