"""OracleStore: a CPU double of dandd_b200.store.GpuSketchStore for the `-m "not gpu"` tests.

TEST INFRASTRUCTURE.  It lets the host logic (naming, caching, hill-climb, tree shapes, CSV and
pickle outputs) be exercised in a container without a GPU by answering the store's questions from
the CPU oracle.  The product never imports this module; on the GPU box the same tests run against
the real store (tests/test_host_gpu.py)."""
import gzip
import os

import numpy as np

from dandd_b200 import hllfile
from oracle import pyoracle as orc


def _read_fasta(path):
    raw = open(path, "rb").read()
    return gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw


class OracleStore:
    def __init__(self, union_files="full"):
        self.union_files = union_files
        self._regs = {}
        self._syms = {}
        self.stats = {"leaf_passes": 0, "union_launches": 0, "files_written": 0, "files_read": 0, "exact_calls": 0}
        self.exact_workers = None

    def symbols(self, fasta):
        if fasta not in self._syms:
            self._syms[fasta] = orc.fasta_symbols(_read_fasta(fasta))
        return self._syms[fasta]

    def registers(self, path):
        if path not in self._regs:
            self._regs[path] = hllfile.read_hll(path)[0]
            self.stats["files_read"] += 1
        return self._regs[path]

    def forget(self, path):
        self._regs.pop(path, None)

    def _write(self, path, regs, p, card):
        os.makedirs(os.path.dirname(path), exist_ok=True)
        hllfile.write_hll(path, regs, p, card)
        self._regs[path] = regs
        self.stats["files_written"] += 1

    def leaf_sketches(self, fasta, ks, p, canon, out_paths, split=None):
        self.stats["leaf_passes"] += 1
        out = {}
        sym = self.symbols(fasta)
        if split is not None:   # collective: my part of the file, registers max-reduced over the ranks (gloo)
            import torch
            from dandd_b200 import dist as dd_dist
            sym = orc.fasta_symbols(dd_dist.split_fasta(_read_fasta(fasta), split[1], only=split[0])[split[0]])
        for k in sorted(ks):
            regs = orc.hll_sketch(sym, int(k), p, canon)
            if split is not None:
                t = torch.from_numpy(regs.copy())
                dd_dist.union_over_ranks(t)
                regs = t.numpy()
            out[k] = orc.card(regs, p)
            if split is None or split[0] == 0:
                self._write(out_paths[k], regs, p, out[k])
        return out

    def union_sketches(self, members_by_k, p, out_paths):
        self.stats["union_launches"] += 1
        out = {}
        for k, members in members_by_k.items():
            regs = orc.union_max([self.registers(m) for m in members])
            out[k] = orc.card(regs, p)
            self._write(out_paths[k], regs, p, out[k])
        return out

    def union_many(self, jobs, p, chunk_bytes=0):
        self.stats["union_launches"] += 1
        out = []
        for members_by_k, out_paths in jobs:
            row = {}
            for k, members in members_by_k.items():
                regs = orc.union_max([self.registers(m) for m in members])
                row[k] = orc.card(regs, p)
                self._write(out_paths[k], regs, p, row[k])
            out.append(row)
        return out

    def prefix_unions(self, leaf_paths_by_k, orderings, p, out_paths=None, chunk_bytes=0):
        self.stats["union_launches"] += 1
        ks = sorted(leaf_paths_by_k)
        orderings = np.asarray(orderings)
        out = np.zeros((orderings.shape[0], orderings.shape[1], len(ks)))
        for o, order in enumerate(orderings):
            for i, k in enumerate(ks):
                run = np.zeros(1 << p, dtype=np.uint8)
                for st, g in enumerate(order):
                    if g >= 0:
                        run = np.maximum(run, self.registers(leaf_paths_by_k[k][g]))
                    out[o, st, i] = orc.card(run, p)
                    path = (out_paths or {}).get((o, st, k))
                    if path and not os.path.exists(path):
                        self._write(path, run.copy(), p, out[o, st, i])
        return out

    def pair_unions(self, leaf_paths_by_k, p, tile_pairs=0, remember=True):
        """[n(n-1)/2, nk] union cardinalities of every pair of leaves (the K6 job of the real store)."""
        self.stats["union_launches"] += 1
        ks = sorted(leaf_paths_by_k)
        n = len(leaf_paths_by_k[ks[0]])
        pairs = [(a, b) for a in range(n) for b in range(a + 1, n)]
        out = np.zeros((len(pairs), len(ks)))
        for j, (a, b) in enumerate(pairs):
            for i, k in enumerate(ks):
                out[j, i] = orc.card(np.maximum(self.registers(leaf_paths_by_k[k][a]), self.registers(leaf_paths_by_k[k][b])), p)
        return out

    def leaf_block(self, fasta, ks, p, canon, out=None):
        """(registers [len(ks), 2^p] torch uint8, cardinalities [len(ks)]): GpuSketchStore.leaf_block on the CPU."""
        import torch
        self.stats["leaf_passes"] += 1
        sym = self.symbols(fasta)
        regs = np.stack([orc.hll_sketch(sym, int(k), p, canon) for k in ks])
        t = torch.from_numpy(regs)
        return (t if out is None else out.copy_(t)), np.asarray([orc.card(r, p) for r in regs], dtype=np.float64)

    def pair_cards(self, regs, pairs, p, tile_pairs=0):
        """[t, nk] union cardinalities of the listed pairs of regs [n, nk, 2^p] (GpuSketchStore.pair_cards)."""
        self.stats["union_launches"] += 1
        regs = regs.numpy()
        pairs = np.asarray(pairs).reshape(-1, 2)
        out = np.zeros((len(pairs), regs.shape[1]))
        for j, (a, b) in enumerate(pairs):
            for i in range(regs.shape[1]):
                out[j, i] = orc.card(np.maximum(regs[a, i], regs[b, i]), p)
        return out

    def materialize_union(self, path, p, card, members):
        self._write(path, orc.union_max([self.registers(m) for m in members]), p, card)
        return float(card)

    def card_of_file(self, path, p=None):
        regs = self.registers(path)
        return orc.card(regs, int(regs.size).bit_length() - 1)

    def exact_count(self, fastas, k, canon):
        return self.exact_prefix_counts(fastas, k, canon)[-1]

    def exact_prefix_counts(self, fastas, k, canon):
        self.stats["exact_calls"] += 1
        if self.exact_workers is not None:
            return self.exact_workers.counts(list(fastas), int(k), bool(canon))
        return self._exact_shard_counts(fastas, k, canon, None)

    def _exact_shard_counts(self, fastas, k, canon, shard):
        """The CPU oracle has no key-range filter: rank 0's "shard" is the whole set, the other ranks
        contribute zeros -- enough to exercise the request / serve / sum protocol."""
        self.stats["shard_calls"] = self.stats.get("shard_calls", 0) + 1
        if shard is not None and shard[0] != 0:
            return [0] * len(fastas)
        return [orc.exact_count([self.symbols(f) for f in fastas[:i + 1]], int(k), canon) for i in range(len(fastas))]

    def start_exact_workers(self):
        from dandd_b200 import dist as dd_dist
        self.exact_workers = dd_dist.ExactWorkers(self._exact_shard_counts)
        return self.exact_workers
