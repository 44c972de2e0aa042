#!/usr/bin/env python
"""Human-scale check of the sketch path (config 3 shape): chromosome-like records,
1 % N runs, 50 % soft-masked (lower-case) bases, k = 2..32 (31 k), p = 20, generated on the GPU.

For each size: pack time, all-k sketch time with and without the min-register floor filter, Gbp/s,
and the invariants that need no oracle (floor on/off give identical registers; sketching the
concatenation equals the max of the parts).  At <= 200 Mbp one k is also checked bit-exact against
the CPU oracle.  Usage: python tools/scale_check.py [--sizes 100e6,1e9,3.1e9]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.synth import synth_fasta  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="100e6,1e9")
    ap.add_argument("--chunk", type=float, default=64e6)
    ap.add_argument("--out", default=None)
    ap.add_argument("--per-k", action="store_true", help="also time every k on its own (floor filter on)")
    args = ap.parse_args()
    from dandd_b200 import build
    build.build()
    from dandd_b200.engine import Engine
    eng = Engine(0)
    ks, p = list(range(2, 33)), 20
    report = []
    for size in [int(float(s)) for s in args.sizes.split(",")]:
        text = synth_fasta(size, 24, seed=3, device=eng.device)
        seq = eng.pack(text, start=0)                                   # first call pays the allocations
        del seq
        seq, t_pack = timed(lambda: eng.pack(text, start=0))
        nsym = seq.nsym
        eng.sketch(seq, [21], p=p)                                       # warm-up
        # default path: dd_sketch_update_sched (floor refreshed at 16 x 2^p x 2^i symbols)
        (r_floor, c_floor), t_floor = timed(lambda: eng.sketch(seq, ks, p=p))
        row = {"bases": size, "text_bytes": int(text.numel()), "symbols": nsym, "pack_ms": t_pack,
               "sketch_floor_ms": t_floor, "gbp_s_floor": size / t_floor / 1e6, "chunk": int(args.chunk),
               "pack_gb_s": text.numel() / t_pack / 1e6}
        if size <= 1.2e9:
            (r_plain, _), t_plain = timed(lambda: eng.sketch(seq, ks, p=p, ranges=[(0, nsym)]))   # no floor filter at all
            row.update(sketch_plain_ms=t_plain, gbp_s_plain=size / t_plain / 1e6,
                       floor_equals_plain=bool((r_plain == r_floor).all()))
            del r_plain
        if size <= 2e8:
            from oracle import pyoracle as orc
            sym = orc.fasta_symbols(text.cpu().numpy().tobytes())
            t0 = time.perf_counter()
            want = orc.hll_sketch(sym, 21, p)
            row.update(oracle_k21_s=time.perf_counter() - t0, oracle_k21_equal=bool(np.array_equal(r_floor[19].cpu().numpy(), want)),
                       nsym_equal=bool(sym.size == nsym))
        if args.per_k:
            row["per_k_ms"] = {}
            for k in ks:
                _, t = timed(lambda: eng.sketch(seq, [k], p=p))
                row["per_k_ms"][k] = round(t, 2)
        cards = c_floor.cpu().numpy()
        deltas = cards / np.array(ks)
        row.update(argmax_k=int(ks[int(np.argmax(deltas))]), delta=float(deltas.max()))
        report.append(row)
        print(json.dumps(row), flush=True)
        del text, seq, r_floor
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(report, fh, indent=1)


if __name__ == "__main__":
    main()
