"""Generates tests/golden/oracle_vectors.json from the numpy/pure-Python restatement
(oracle/ref_numpy.py), NOT from the C oracle that the fixture is used to check.

Run from the repository root:  python tests/golden/make_oracle_vectors.py
No reference goldens exist for this path (parity unpinned); these vectors pin the
oracle's behaviour across rounds."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_numpy as ref  # noqa: E402
from tests.util import adversarial_fasta  # noqa: E402


def main():
    cases = []
    for seed, n, p in [(11, 4000, 10), (12, 20000, 12), (13, 60000, 14)]:
        rng = np.random.default_rng(seed)
        txt = adversarial_fasta(rng, n=n)
        sym = ref.fasta_symbols_py(txt)
        rows = []
        for k, canon in [(4, True), (12, True), (16, True), (17, True), (21, False), (31, True), (32, True), (32, False)]:
            regs = ref.hll_sketch_np(sym, k, p, canon)
            hist = np.bincount(regs, minlength=66)
            last = int(np.flatnonzero(hist)[-1]) + 1
            rows.append({
                "k": k, "canon": canon, "hist": hist[:last].tolist(),
                "checksum": int(np.dot(regs.astype(np.int64), np.arange(regs.size) % 251)),
                "card": ref.ertl_mle_py(hist, p),
                "exact": ref.exact_count_np([sym], k, canon),
            })
        cases.append({"seed": seed, "n": n, "p": p, "nsym": int(sym.size), "nbreak": int((sym == 4).sum()), "rows": rows})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.json")
    with open(out, "w") as f:
        json.dump({"generator": "tests/golden/make_oracle_vectors.py (oracle/ref_numpy.py)", "cases": cases}, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
