#!/usr/bin/env python
"""cProfile of one `dandd tree --ksweep` at config-2 scale (12 x 5 Mbp, k = 10..32) in-process, to see
where the command line's wall time goes once the kernels take milliseconds."""
import cProfile
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dandd_b200", "lib"))
t_start = time.perf_counter()
import bench  # noqa: E402


def main():
    work = tempfile.mkdtemp(prefix="dd_cliprof_")
    data = os.path.join(work, "fastas")
    os.makedirs(data)
    for i, (text, _) in enumerate(bench.make_genomes(seed=2)):
        with open(os.path.join(data, f"genome{i:02d}.fasta"), "wb") as fh:
            fh.write(text)
    t0 = time.perf_counter()
    import dandd_cmd
    t1 = time.perf_counter()
    print(f"import dandd_cmd (torch, library): {t1 - t0:.2f} s", flush=True)
    parser, _ = dandd_cmd.parse_arguments()
    argv = ["tree", "-d", data, "-s", "cfg2", "-k", "14", "-o", os.path.join(work, "out"), "--ksweep", "--mink", "10", "--maxk", "32"]
    args = parser.parse_args(argv)
    prof = cProfile.Profile()
    t2 = time.perf_counter()
    prof.enable()
    args.func(args)
    prof.disable()
    print(f"tree --ksweep body: {time.perf_counter() - t2:.2f} s", flush=True)
    pstats.Stats(prof).sort_stats("cumulative").print_stats(45)


if __name__ == "__main__":
    main()
