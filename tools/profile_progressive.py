#!/usr/bin/env python
"""cProfile of `dandd progressive` (config-2 scale, stub union files) on the GPU box."""
import cProfile, io, os, pickle, pstats, random, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tests.host_harness import run_dandd
os.environ["DANDD_B200_UNION_FILES"] = os.environ.get("DANDD_B200_UNION_FILES", "stub")
work = tempfile.mkdtemp(prefix="dd_prof_")
data = os.path.join(work, "fastas"); os.makedirs(data)
for i, (text, _) in enumerate(bench.make_genomes(seed=2)):
    open(os.path.join(data, f"genome{i:02d}.fasta"), "wb").write(text)
random.seed(2); ords = set()
while len(ords) < 30: ords.add(tuple(random.sample(range(12), 12)))
of = os.path.join(work, "o.pickle"); pickle.dump(ords, open(of, "wb"))
sweep = ["--ksweep", "--mink", "10", "--maxk", "32"]
out = os.path.join(work, "out")
t0 = time.perf_counter(); run_dandd(["tree", "-d", data, "-s", "c2", "-k", "14", "-o", out] + sweep); t_tree = time.perf_counter() - t0
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter()
run_dandd(["progressive", "-d", os.path.join(out, "c2_12_dashing_dtree.pickle"), "-r", of, "-o", out] + sweep)
t_prog = time.perf_counter() - t0; pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30)
print("tree_s", round(t_tree, 2), "progressive_s", round(t_prog, 2)); print(s.getvalue()[:5500])
