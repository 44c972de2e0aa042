#!/usr/bin/env python
"""Config 5 shape (helpers/allpairs.py): N x N union cardinalities / KIJ / per-k Jaccard from
HBM-resident sketches.  Synthetic sketches (random registers with a realistic rank distribution) are
enough to time K6; a small subset is checked against the oracle.
    python tools/allpairs_check.py --n 200 --p 18 --kmin 10 --kmax 32"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--p", type=int, default=18)
    ap.add_argument("--kmin", type=int, default=10)
    ap.add_argument("--kmax", type=int, default=32)
    ap.add_argument("--tile", type=int, default=20000, help="pairs per launch")
    args = ap.parse_args()
    from dandd_b200 import build
    build.build()
    from dandd_b200.engine import Engine
    from oracle import pyoracle as orc
    eng = Engine(0)
    n, p, nk = args.n, args.p, args.kmax - args.kmin + 1
    m = 1 << p
    g = torch.Generator(device=eng.device)
    g.manual_seed(5)
    # geometric ranks: floor(-log2(u)) + 1, shifted so that values look like a filled sketch
    u = torch.rand((n, nk, m), device=eng.device, generator=g)
    regs = (torch.clamp((-torch.log2(u)).floor() + 3, max=64 - p + 1)).to(torch.uint8)
    del u
    pairs = np.array([(a, b) for a in range(n) for b in range(a + 1, n)], dtype=np.int32)
    single = eng.cards(regs, p)                                        # [n, nk]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = torch.empty((len(pairs), nk), dtype=torch.float64, device=eng.device)
    e0.record()
    for t0 in range(0, len(pairs), args.tile):
        out[t0:t0 + args.tile] = eng.pairwise_cards(regs, pairs[t0:t0 + args.tile], p)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    karr = torch.arange(args.kmin, args.kmax + 1, device=eng.device, dtype=torch.float64)
    d_single = (single / karr).max(dim=1).values
    d_pair = (out / karr).max(dim=1).values
    pa, pb = torch.as_tensor(pairs[:, 0], device=eng.device).long(), torch.as_tensor(pairs[:, 1], device=eng.device).long()
    kij = (d_single[pa] + d_single[pb] - d_pair) / d_pair
    jac = (single[pa] + single[pb] - out) / out
    h = regs[:3].cpu().numpy()
    ok = True
    for j, (a, b) in enumerate(pairs[:2].tolist()):
        for i in (0, nk - 1):
            want = orc.card(np.maximum(h[a, i], h[b, i]), p)
            ok &= abs(float(out[j, i]) - want) <= 1e-9 * want
    cells = len(pairs) * nk
    print(json.dumps({"n": n, "p": p, "nk": nk, "pairs": len(pairs), "ms": ms, "pair_k_per_s": cells / ms * 1e3,
                      "register_bytes_per_s": cells * 2 * m / ms * 1e3, "oracle_ok": bool(ok),
                      "kij_mean": float(kij.mean()), "jaccard_mean": float(jac.mean())}))


if __name__ == "__main__":
    main()
