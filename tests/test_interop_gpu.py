"""the `dashing`-argv front end on the GPU library interoperates file-for-file
with the oracle-backed `dashing` stand-in (same names, same layout, same numbers), .gz FASTA input
works, and the background ingest pool hands back the right bytes and digests."""
import gzip
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from dandd_b200 import hllfile
from oracle import pyoracle as orc
from tests.util import make_dataset

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_CLI = os.path.join(ROOT, "dandd_b200", "lib", "dashing_b200")


def test_ingest_pool_bytes_and_digests(tmp_path):
    from dandd_b200 import ingest
    paths = make_dataset(str(tmp_path / "d"), 3, 5000, seed=5)
    gz = str(tmp_path / "d" / "z.fasta.gz")
    with gzip.open(gz, "wb") as fh:
        fh.write(open(paths[0], "rb").read())
    ingest.prefetch(paths + [gz])
    for p in paths + [gz]:
        assert ingest.digest(p) == hashlib.blake2b(open(p, "rb").read()).hexdigest()
    assert ingest.fasta_bytes(gz) == open(paths[0], "rb").read()
    assert ingest.fasta_bytes(paths[1]) == open(paths[1], "rb").read()
    assert ingest.fasta_bytes(paths[1]) == open(paths[1], "rb").read()   # second time: loaded on demand


@pytest.mark.gpu
def test_dashing_cli_interop(tmp_path):
    from dandd_b200 import build
    build.build()
    paths = make_dataset(str(tmp_path / "d"), 2, 30000, seed=6)
    gz = str(tmp_path / "d" / "g0.fasta.gz")
    with gzip.open(gz, "wb") as fh:
        fh.write(open(paths[0], "rb").read())
    cpu_bin = orc.install_shims(str(tmp_path / "bin"))
    a, b = tmp_path / "gpu", tmp_path / "cpu"
    a.mkdir(), b.mkdir()
    run_gpu = lambda *argv: subprocess.run([sys.executable, GPU_CLI, *argv], check=True, capture_output=True, text=True).stdout
    run_cpu = lambda *argv: subprocess.run([os.path.join(cpu_bin, "dashing"), *argv], check=True, capture_output=True, text=True).stdout
    for run, d in ((run_gpu, a), (run_cpu, b)):
        run("sketch", "-k21", "-S", "14", "--prefix", str(d), paths[0], paths[1], gz)
        run("union", "-z", "-o", str(d / "u.hll"), str(d / "g0.fasta.w.21.spacing.14.hll"), str(d / "g1.fasta.w.21.spacing.14.hll"))
    names = sorted(os.listdir(a))
    assert names == sorted(os.listdir(b)) and "g0.fasta.gz.w.21.spacing.14.hll" in names
    for n in names:
        ra, pa, _ = hllfile.read_hll(str(a / n))
        rb, pb, _ = hllfile.read_hll(str(b / n))
        assert pa == pb == 14 and np.array_equal(ra, rb), n
    assert np.array_equal(hllfile.read_hll(str(a / "g0.fasta.gz.w.21.spacing.14.hll"))[0],
                          hllfile.read_hll(str(a / "g0.fasta.w.21.spacing.14.hll"))[0])
    # each tool estimates the OTHER tool's files and they agree with themselves
    ca = run_gpu("card", "--presketched", str(b / "u.hll"), str(b / "g1.fasta.w.21.spacing.14.hll")).splitlines()
    cb = run_cpu("card", "--presketched", str(a / "u.hll"), str(a / "g1.fasta.w.21.spacing.14.hll")).splitlines()
    assert ca[0] == cb[0] == "#Path\tSize (est.)"
    for la, lb in zip(ca[1:], cb[1:]):
        assert float(la.split("\t")[1]) == pytest.approx(float(lb.split("\t")[1]), rel=1e-9, abs=2e-6)
    ha = float(run_gpu("hll", "-k", "17", "-S", "12", paths[0], paths[1]).split()[-1])
    hb = float(run_cpu("hll", "-k", "17", "-S", "12", paths[0], paths[1]).split()[-1])
    assert ha == pytest.approx(hb, rel=1e-9, abs=2e-6)
