#!/usr/bin/env python
"""Pack one synthetic FASTA text a few times (K1 only) -- a target for ncu launch lists / captures.
usage: pack_probe.py [bases] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.scale_check import synth_fasta  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    from dandd_b200.engine import Engine
    eng = Engine(0)
    text = synth_fasta(n, 24, seed=3, device=eng.device)
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        seq = eng.pack(text, start=0)
        e1.record()
        torch.cuda.synchronize()
        print("pack %d bytes -> %d symbols: %.3f ms" % (text.numel(), seq.nsym, e0.elapsed_time(e1)), flush=True)


if __name__ == "__main__":
    main()
