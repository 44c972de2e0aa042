#!/usr/bin/env python
"""`dandd tree` / `progressive` wall-time at config-2 scale through the drop-in command line
(files, pickles and CSVs included): 12 x 5 Mbp FASTAs on local disk, k = 10..32, 30 orderings.
Run once with full union files (what `dashing union -o` leaves behind) and once with
DANDD_B200_UNION_FILES=stub."""
import json
import os
import pickle
import random
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

CLI = os.path.join(ROOT, "dandd_b200", "lib", "dandd")


def run(argv, env=None):
    t0 = time.perf_counter()
    subprocess.run([sys.executable, CLI] + argv, check=True, stdout=subprocess.DEVNULL, env=dict(os.environ, **(env or {})))
    return time.perf_counter() - t0


def main():
    work = tempfile.mkdtemp(prefix="dd_cfg2_")
    rep = {}
    try:
        data = os.path.join(work, "fastas")
        os.makedirs(data)
        for i, (text, _) in enumerate(bench.make_genomes(seed=2)):
            with open(os.path.join(data, f"genome{i:02d}.fasta"), "wb") as fh:
                fh.write(text)
        random.seed(2)
        orderings = set()
        while len(orderings) < 30:
            orderings.add(tuple(random.sample(range(12), 12)))
        ofile = os.path.join(work, "orderings.pickle")
        with open(ofile, "wb") as fh:
            pickle.dump(orderings, fh)
        sweep = ["--ksweep", "--mink", "10", "--maxk", "32"]
        for mode in ("stub", "full"):
            out = os.path.join(work, "out_" + mode)
            env = {"DANDD_B200_UNION_FILES": mode}
            rep[mode] = {"tree_s": run(["tree", "-d", data, "-s", "cfg2", "-k", "14", "-o", out] + sweep, env)}
            dtree = os.path.join(out, "cfg2_12_dashing_dtree.pickle")
            rep[mode]["progressive_s"] = run(["progressive", "-d", dtree, "-r", ofile, "-o", out] + sweep, env)
            rep[mode]["tree_again_s"] = run(["tree", "-d", data, "-s", "cfg2", "-k", "14", "-o", out] + sweep, env)
            files = sum(len(fs) for _, _, fs in os.walk(os.path.join(out, "sketchdb")))
            size = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(os.path.join(out, "sketchdb")) for f in fs)
            rep[mode].update(sketchdb_files=files, sketchdb_GB=size / 1e9)
            shutil.rmtree(out)
        rep["tree_hillclimb_s"] = run(["tree", "-d", data, "-s", "cfg2h", "-k", "14", "-o", os.path.join(work, "out_h")])
    finally:
        shutil.rmtree(work, ignore_errors=True)
    print(json.dumps(rep))


if __name__ == "__main__":
    main()
