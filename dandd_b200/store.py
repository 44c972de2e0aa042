"""SketchStore: what the drop-in sketch objects (dandd_b200/lib/sketch_classes.py) call instead of
shelling out to dashing / kmc / kmc_tools / GNU parallel.

The reference's only state between commands is the sketchdb directory (files + cardinality
pickle).  The store keeps that contract -- every sketch it produces is written to the path DandD
expects -- and adds an HBM-resident cache of registers and packed sequences so that unions,
progressive prefix unions and pairwise unions never re-read a file or re-sketch a FASTA:

    dashing sketch (x nk via parallel)  -> leaf_sketches()      one fused all-k pass on the GPU
    dashing union  (x nk via parallel)  -> union_sketches()     one batched max+histogram+MLE launch
    dashing card                        -> card_of_file()
    kmc + kmc_tools info / complex      -> exact_count()        GPU k-mer set (bitmap / hash set)

One store per process; it talks to one Engine (one GPU).  There is no CPU path here.
"""
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import hllfile, ingest, timing
from .engine import Engine, get_engine

ALL_HLL_KS = tuple(range(1, 33))
STREAM_MIN_BYTES = int(os.environ.get("DANDD_B200_STREAM_MIN", str(32 << 20)))   # larger FASTAs take the chunked pipeline


def read_fasta_bytes(path: str) -> bytes:
    """Whole FASTA as bytes; .gz is inflated on the host (dashing/kmc read .gz transparently).
    Served from the background ingest pool when the file was prefetched."""
    return ingest.fasta_bytes(path)


class GpuSketchStore:
    def __init__(self, engine: Engine = None, prefetch_all_k: bool = True, cache_bytes: int = 64 << 30,
                 hll_compresslevel: int = None, union_files: str = None):
        self.engine = engine or get_engine()
        self.prefetch_all_k = prefetch_all_k
        self.cache_bytes = cache_bytes
        self.hll_compresslevel = int(os.environ.get("DANDD_B200_HLL_GZIP", "0")) if hll_compresslevel is None else hll_compresslevel
        # 'full': union sketches are written out like `dashing union -o` does; 'stub': a header-only
        # marker file (registers stay in HBM / are recomputed from the leaves on demand)
        self.union_files = union_files or os.environ.get("DANDD_B200_UNION_FILES", "full")
        # all-pairs jobs leave N(N-1)/2 x (a few k) union sketches behind: markers unless full files were asked for
        self.pair_union_files = union_files or os.environ.get("DANDD_B200_UNION_FILES", "stub")
        # One LRU over everything resident in HBM, keyed by kind: ("sketch", path) -> [2^p] u8,
        # ("leaf", fasta, p, canon) -> {"regs","cards","ks"}, ("packed", fasta) -> PackedSeq.  A sketch that
        # is a VIEW of a leaf block costs nothing by itself (the block is what occupies memory).
        self._lru: "OrderedDict[tuple, object]" = OrderedDict()
        self._cost: Dict[tuple, int] = {}
        self._bytes = 0
        self.pair_table = None   # cardinalities of every two-leaf union, filled by one batched K6 job (pair_unions)
        self.stats = {"leaf_passes": 0, "union_launches": 0, "files_written": 0, "files_read": 0, "exact_calls": 0}
        self.exact_workers = None   # set by start_exact_workers() for `--exact` under torchrun
        self.stream_stats: List[dict] = []   # per streamed FASTA: stage times (read, blake2b, GPU span, wall)

    # ------------------------------------------------------------------ cache plumbing
    def _put(self, key: tuple, value, cost: int) -> None:
        if key in self._lru:
            self._bytes -= self._cost.pop(key)
            del self._lru[key]
        self._lru[key] = value
        self._cost[key] = int(cost)
        self._bytes += int(cost)
        while self._bytes > self.cache_bytes and len(self._lru) > 1:
            old, gone = self._lru.popitem(last=False)
            self._bytes -= self._cost.pop(old)
            if old[0] == "leaf":       # the per-k sketches that are views of this block would keep it in HBM, uncounted
                for path in gone.get("views", ()):
                    self.forget(path)

    def _get(self, key: tuple):
        value = self._lru.get(key)
        if value is not None:
            self._lru.move_to_end(key)
        return value

    def _remember(self, path: str, regs: torch.Tensor, view_of_block: bool = False) -> None:
        self._put(("sketch", path), regs, 0 if view_of_block else regs.numel())

    def registers(self, path: str) -> torch.Tensor:
        """Device registers of the sketch stored at `path` (HBM cache, else read the file).  The
        returned tensor stays valid for as long as the caller holds it, whatever the cache evicts."""
        t = self._get(("sketch", path))
        if t is None:
            with timing.span("read_sketch_files"):
                try:
                    regs, _p, _ = hllfile.read_hll(path)
                    t = torch.from_numpy(regs).to(self.engine.device)
                except hllfile.StubSketch as stub:       # union marker: rebuild from its members
                    t = self.engine.union([self.registers(m) for m in stub.members])
            self.stats["files_read"] += 1
            self._remember(path, t)
        return t

    def forget(self, path: str) -> None:
        key = ("sketch", path)
        if key in self._lru:
            del self._lru[key]
            self._bytes -= self._cost.pop(key)

    def _write(self, path: str, regs: torch.Tensor, p: int, card: float, leaf: bool, members=None) -> None:
        with timing.span("write_sketch_files"):
            os.makedirs(os.path.dirname(path), exist_ok=True)
            if leaf or self.union_files == "full" or not members:
                hllfile.write_hll(path, regs.cpu().numpy(), p, card, self.hll_compresslevel)
            else:
                hllfile.write_stub(path, p, card, members)
        self.stats["files_written"] += 1

    # ------------------------------------------------------------------ dashing sketch
    def _pack_text(self, text):
        """K1, with the FASTQ detour: the packer only reports a line beginning with '+'; the text is
        then rewritten record by record the way kseq reads it (host side) and packed again."""
        from .engine import FastqInput
        try:
            return self.engine.pack(text).check()
        except FastqInput:
            return self.engine.pack(self.engine.fastq_to_fasta(text)).check()

    def packed(self, fasta: str):
        seq = self._get(("packed", fasta))
        if seq is None:
            seq = self._pack_text(read_fasta_bytes(fasta))
            self._put(("packed", fasta), seq, seq.codes.numel() + seq.invalid.numel())
        return seq

    def leaf_sketches(self, fasta: str, ks: Sequence[int], p: int, canon: bool, out_paths: Dict[int, str],
                      split: Optional[Tuple[int, int]] = None) -> Dict[int, float]:
        """Sketch `fasta` for every k in `ks` (one fused pass), write each sketch to out_paths[k]
        and return {k: cardinality}.  With prefetch_all_k the pass covers k = 1..32 once and later
        requests for other k of the same FASTA are served from HBM.

        split=(rank, world): a COLLECTIVE call -- every rank sketches its part of the file
        (dist.split_fasta), the registers are max-reduced over the ranks, every rank gets the
        cardinalities of the whole file and rank 0 writes the sketch files (fewer
        genomes than GPUs)."""
        need = [int(k) for k in ks]
        ent = self._leaf_entry(fasta, need, p, canon, split)
        out = {}
        for k in need:
            i = ent["ks"][k]
            card = float(ent["cards"][i])
            ent.setdefault("views", set()).add(out_paths[k])     # (first: the next line may evict the block itself)
            self._remember(out_paths[k], ent["regs"][i], view_of_block=True)
            if split is None or split[0] == 0:
                self._write(out_paths[k], ent["regs"][i], p, card, leaf=True)
            out[k] = card
        return out

    def leaf_block(self, fasta: str, ks: Sequence[int], p: int, canon: bool,
                   out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, np.ndarray]:
        """(registers [len(ks), 2^p] uint8 on the device, cardinalities [len(ks)]) of one FASTA, rows in the
        order of `ks`; nothing is named or written (`dashing hll` leaves no sketch behind:
        reference helpers/allpairs.py:32-35).  Same fused pass as leaf_sketches, for exactly these k; a
        block already in the HBM cache is used, a new one is not added to it (the caller keeps the
        registers: with `out` they are written straight into its [len(ks), 2^p] slice)."""
        need = [int(k) for k in ks]
        ent = self._leaf_entry(fasta, need, p, canon, None, all_k=False, keep=False)
        rows = [ent["ks"][k] for k in need]
        cards = np.asarray([float(ent["cards"][i]) for i in rows], dtype=np.float64)
        if out is not None and (tuple(out.shape) != (len(rows), 1 << p) or out.dtype != torch.uint8 or not out.is_contiguous()):
            raise ValueError(f"leaf_block: out must be a contiguous uint8 [{len(rows)}, {1 << p}] tensor, got {tuple(out.shape)}")
        if rows == list(range(ent["regs"].shape[0])):
            regs = ent["regs"] if out is None else out.copy_(ent["regs"])
        else:
            index = torch.as_tensor(rows, dtype=torch.int64, device=ent["regs"].device)
            regs = torch.index_select(ent["regs"], 0, index) if out is None else torch.index_select(ent["regs"], 0, index, out=out)
        return regs, cards

    def _leaf_entry(self, fasta: str, need: List[int], p: int, canon: bool, split: Optional[Tuple[int, int]],
                    all_k: Optional[bool] = None, keep: bool = True) -> dict:
        """The all-k block of one FASTA: {"regs": [nk, 2^p] device tensor, "cards": [nk], "ks": {k: row}},
        from the HBM cache or one fused pass (streamed above STREAM_MIN_BYTES).  all_k: sketch k = 1..32
        whatever is asked for (default: the store's prefetch_all_k); keep: remember a new block in the cache."""
        key = ("leaf", fasta, int(p), bool(canon))
        ent = self._get(key)
        if split is not None or ent is None or any(k not in ent["ks"] for k in need):
            all_k = self.prefetch_all_k if all_k is None else all_k
            run_ks = list(ALL_HLL_KS) if all_k else sorted(set(need) | set(ent["ks"] if ent else ()))
            streamed = None
            if split is None and os.path.getsize(fasta) >= STREAM_MIN_BYTES:
                streamed = self._stream_leaf(fasta, run_ks, p, canon)
            text = read_fasta_bytes(fasta) if streamed is None else None
            if split is not None:
                from dandd_b200 import dist as dd_dist
                run_ks = sorted(need)                      # identical on every rank, whatever each one has cached
                if b"\n+" in text or text[:1] == b"+":     # FASTQ: records as kseq reads them, then split
                    text = self.engine.fastq_to_fasta(text)
                text = dd_dist.split_fasta(text, split[1], only=split[0])[split[0]]
            if streamed is not None:
                regs, cards = streamed
            else:
                seq = self._pack_text(text)
                regs, cards = self.engine.sketch(seq, run_ks, p=p, canon=canon)
                if split is not None:
                    dd_dist.union_over_ranks(regs)
                    cards = self.engine.cards(regs, p)
                cards = cards.cpu().numpy()
            ent = {"regs": regs, "cards": cards, "ks": {k: i for i, k in enumerate(run_ks)}}
            if keep:
                self._put(key, ent, regs.numel())
            self.stats["leaf_passes"] += 1
        return ent

    def warm_leaf(self, fasta: str, p: int, canon: bool) -> None:
        """Sketch a large FASTA for all k NOW and keep the block in HBM, without naming or writing
        anything: lets the caller start the GPU work before the file's blake2b name (which the sketch
        database paths need, and which takes seconds for a multi-GB file) has been computed."""
        key = ("leaf", fasta, int(p), bool(canon))
        if not self.prefetch_all_k or self._get(key) is not None or not os.path.isfile(fasta) \
                or os.path.getsize(fasta) < STREAM_MIN_BYTES:
            return
        run_ks = list(ALL_HLL_KS)
        streamed = self._stream_leaf(fasta, run_ks, p, canon)
        if streamed is not None:
            regs, cards = streamed
            self._put(key, {"regs": regs, "cards": cards, "ks": {k: i for i, k in enumerate(run_ks)}}, regs.numel())
            self.stats["leaf_passes"] += 1

    def _stream_leaf(self, fasta, run_ks, p, canon):
        """Large FASTA: disk (or the prefetched bytes) -> pinned ring -> H2D -> K1 -> K2 in 64 MiB chunks,
        the blake2b name computed from the same chunks (dandd_b200/streaming.py).  None if the text is
        FASTQ (the whole-file path then rewrites it first)."""
        from . import streaming
        from .engine import FastqInput
        text = ingest.fasta_bytes(fasta) if ingest.is_prefetched(fasta) else None
        try:
            regs, cards, digest, stats = streaming.sketch_file(self.engine, fasta, run_ks, p, canon, text=text,
                                                               want_digest=not ingest.has_digest(fasta))
        except FastqInput:
            return None
        if digest:
            ingest.set_digest(fasta, digest)
        self.stream_stats.append(stats)
        for key in ("read_s", "blake2b_s", "inflate_s", "gpu_span_s", "wall_s"):
            if key in stats:
                timing.add("stream_" + key, stats[key])
        return regs, cards

    # ------------------------------------------------------------------ dashing union (+ card)
    def union_sketches(self, members_by_k: Dict[int, List[str]], p: int, out_paths: Dict[int, str]) -> Dict[int, float]:
        """For every k: union of the sketches stored at members_by_k[k]; written to out_paths[k];
        returns {k: cardinality}.  One launch for all k."""
        out = {}
        for k in sorted(members_by_k):   # cells a batched pair job (pair_unions) has already evaluated
            card = self._pair_lookup(members_by_k[k], p)
            if card is not None:
                out[k] = self.materialize_union(out_paths[k], p, card, members_by_k[k])
        ks = sorted(k for k in members_by_k if k not in out)
        if not ks:
            return out
        width = max(len(members_by_k[k]) for k in ks)
        tensors = [[self.registers(pth) for pth in members_by_k[k]] for k in ks]
        ptrs = np.zeros((len(ks), width), dtype=np.int64)
        for i, row in enumerate(tensors):
            for j, t in enumerate(row):
                ptrs[i, j] = t.data_ptr()
        with timing.span("union_gpu"):
            cards, unions = self.engine.union_sets(ptrs, p, final_only=True, materialize=True)
            self.stats["union_launches"] += 1
            cards = cards.cpu().numpy().reshape(len(ks))
        for i, k in enumerate(ks):
            regs = unions[i, 0]
            self._remember(out_paths[k], regs)
            self._write(out_paths[k], regs, p, float(cards[i]), leaf=False, members=list(members_by_k[k]))
            out[k] = float(cards[i])
        return out

    def union_many(self, jobs: Sequence[Tuple[Dict[int, List[str]], Dict[int, str]]], p: int,
                   chunk_bytes: int = 4 << 30) -> List[Dict[int, float]]:
        """union_sketches for MANY nodes at once (every inner node of a tree x every k): jobs[i] =
        (members_by_k, out_paths).  The (node, k) cells are grouped by member count (powers of two, so the
        pointer tables stay dense) and each group goes to the device as one launch -- or a few, when full
        union files are wanted and the materialised unions of a group exceed chunk_bytes.  Returns one
        {k: cardinality} per job; every out_path exists afterwards."""
        cells = [(i, k, list(members[k])) for i, (members, _) in enumerate(jobs) for k in sorted(members)]
        out: List[Dict[int, float]] = [dict() for _ in jobs]
        groups: Dict[int, list] = {}
        for cell in cells:
            groups.setdefault(max(1, len(cell[2]) - 1).bit_length(), []).append(cell)
        want_regs = self.union_files == "full"
        m = 1 << p
        for _, group in sorted(groups.items()):
            step = max(1, chunk_bytes // m) if want_regs else len(group)
            for g0 in range(0, len(group), step):
                part = group[g0:g0 + step]
                width = max(len(c[2]) for c in part)
                held = [[self.registers(pth) for pth in c[2]] for c in part]
                ptrs = np.zeros((len(part), width), dtype=np.int64)
                for r, row in enumerate(held):
                    ptrs[r, :len(row)] = [t.data_ptr() for t in row]
                with timing.span("union_gpu"):
                    res = self.engine.union_sets(ptrs, p, final_only=True, materialize=want_regs)
                    cards, unions = (res if want_regs else (res, None))
                    cards = cards.cpu().numpy().reshape(len(part))
                self.stats["union_launches"] += 1
                for r, (i, k, members) in enumerate(part):
                    path = jobs[i][1][k]
                    card = float(cards[r])
                    if want_regs:
                        self._remember(path, unions[r, 0])
                        self._write(path, unions[r, 0], p, card, leaf=False, members=members)
                    else:
                        self._write(path, None, p, card, leaf=False, members=members)
                    out[i][k] = card
                del held
        return out

    def prefix_unions(self, leaf_paths_by_k: Dict[int, List[str]], orderings: Sequence[Sequence[int]], p: int,
                      out_paths: Dict[tuple, str] = None, chunk_bytes: int = 4 << 30) -> np.ndarray:
        """Cardinalities of every prefix union: [n_orderings, n_steps, nk] for the k values (sorted)
        of leaf_paths_by_k, each listing the leaf sketches in genome-index order.  A running max per
        (ordering, k): n sketch reads instead of the n(n+1)/2 of re-unioning every prefix.
        out_paths maps (ordering index, step index, k) -> file path for the prefix unions that must
        also exist as sketch files; those are materialised (chunk_bytes of HBM at a time) and
        written once per distinct path."""
        ks = sorted(leaf_paths_by_k)
        orderings = np.asarray(orderings, dtype=np.int64)
        n_ord, n_steps = orderings.shape
        m = 1 << p
        # hold every leaf tensor for the duration of the call: a cache miss further down this list may
        # evict (and free) sketches whose addresses are already in the table
        held = [[self.registers(pth) for pth in leaf_paths_by_k[k]] for k in ks]
        base = np.array([[t.data_ptr() for t in row] for row in held], dtype=np.int64)  # [nk, n]
        # rows are laid out [ordering-chunk][k][ordering] at launch time: sets that read the same
        # sketches (same k, different ordering) are adjacent, which is what keeps them in L2
        ptrs = np.zeros((n_ord, len(ks), n_steps), dtype=np.int64)
        for i in range(len(ks)):
            ptrs[:, i, :] = np.where(orderings >= 0, base[i][np.clip(orderings, 0, None)], 0)
        out = np.empty((n_ord, n_steps, len(ks)), dtype=np.float64)
        want_files = bool(out_paths)
        per_ord = len(ks) * n_steps * m
        step_ords = max(1, chunk_bytes // per_ord) if want_files else n_ord
        written = set()
        for o0 in range(0, n_ord, step_ords):
            o1 = min(n_ord, o0 + step_ords)
            block = np.ascontiguousarray(ptrs[o0:o1].transpose(1, 0, 2)).reshape(len(ks) * (o1 - o0), n_steps)  # [k][o]
            if want_files:
                cards, unions = self.engine.union_sets(block, p, final_only=False, materialize=True)
                unions = unions.view(len(ks), o1 - o0, n_steps, m).transpose(0, 1)   # -> [o][k][step][m]
            else:
                cards = self.engine.union_sets(block, p, final_only=False)
            self.stats["union_launches"] += 1
            cards = cards.cpu().numpy().reshape(len(ks), o1 - o0, n_steps).transpose(1, 0, 2)   # [o][k][step]
            out[o0:o1] = cards.transpose(0, 2, 1)
            if want_files:
                for o in range(o0, o1):
                    for st in range(n_steps):
                        for i, k in enumerate(ks):
                            path = out_paths.get((o, st, k))
                            if path and path not in written:
                                written.add(path)
                                members = [leaf_paths_by_k[k][g] for g in orderings[o][:st + 1] if g >= 0]
                                self._write(path, unions[o - o0, i, st], p, float(cards[o - o0, i, st]), leaf=False,
                                            members=members)
        del held   # every launch above was followed by a device->host read of its results, so nothing still uses them
        return out

    # ------------------------------------------------------------------ all pairs (K6)
    def pair_unions(self, leaf_paths_by_k: Dict[int, List[str]], p: int, tile_pairs: int = 1 << 16,
                    remember: bool = True) -> np.ndarray:
        """Cardinality of the union of EVERY unordered pair of leaves, for every k: [n(n-1)/2, nk], pairs
        in the order (0,1), (0,2) .. (n-2,n-1), k sorted.  leaf_paths_by_k[k] lists the leaf sketches in
        leaf-index order.  The sketches are transposed into bit planes ONCE; the pair list then goes
        through K6 in tiles.  With `remember` the table also answers later union_sketches() calls for
        two-leaf unions (`dandd kij` builds one SubSpider per pair: reference lib/huffman_dandd.py:666-695)."""
        ks = sorted(leaf_paths_by_k)
        n = len(leaf_paths_by_k[ks[0]])
        m = 1 << p
        eng = self.engine
        regs = torch.empty((n, len(ks), m), dtype=torch.uint8, device=eng.device)
        for i, k in enumerate(ks):
            for g, pth in enumerate(leaf_paths_by_k[k]):
                regs[g, i].copy_(self.registers(pth))
        iu = np.triu_indices(n, 1)
        pairs = np.stack(iu, axis=1).astype(np.int32)
        out = self.pair_cards(regs, pairs, p, tile_pairs)
        if remember:
            self.pair_table = {"p": int(p), "n": n, "col": {k: i for i, k in enumerate(ks)}, "cards": out,
                               "leaf": {pth: (g, k) for k in ks for g, pth in enumerate(leaf_paths_by_k[k])}}
        return out

    def pair_cards(self, regs: torch.Tensor, pairs, p: int, tile_pairs: int = 1 << 16) -> np.ndarray:
        """card(A u B) for the listed pairs (int [t, 2] genome indices into regs [n, nk, 2^p], on the device)
        at every k: float64 [t, nk] on the host.  The sketches are transposed into bit planes once, the
        pair list goes through K6 in tiles (reference helpers/allpairs.py:360-370 runs one
        `dashing hll A B` -- two FASTA passes -- per pair and k)."""
        eng = self.engine
        n, nk = int(regs.shape[0]), int(regs.shape[1])
        pairs = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(-1, 2))
        out = np.empty((pairs.shape[0], nk), dtype=np.float64)
        if not pairs.shape[0]:
            return out
        planes = eng.to_planes(regs, p) if p >= 12 else None
        for t0 in range(0, pairs.shape[0], tile_pairs):
            tile = pairs[t0:t0 + tile_pairs]
            if planes is not None:
                cards = eng.pairwise_cards(None, tile, p, planes=planes, n_genomes=n, nk=nk)
            else:
                cards = eng.pairwise_cards(regs, tile, p)
            out[t0:t0 + tile.shape[0]] = cards.cpu().numpy()
            self.stats["union_launches"] += 1
        return out

    def _pair_lookup(self, members: Sequence[str], p: int):
        """Cardinality of the union of two leaf sketches if the last pair_unions() job covered it."""
        tab = self.pair_table
        if tab is None or len(members) != 2 or tab["p"] != int(p):
            return None
        a, b = tab["leaf"].get(members[0]), tab["leaf"].get(members[1])
        if a is None or b is None or a[1] != b[1] or a[0] == b[0]:
            return None
        i, j = (a[0], b[0]) if a[0] < b[0] else (b[0], a[0])
        n = tab["n"]
        return float(tab["cards"][i * n - i * (i + 1) // 2 + (j - i - 1), tab["col"][a[1]]])

    def materialize_union(self, path: str, p: int, card: float, members) -> float:
        """A union whose cardinality a batched job already produced is asked for as a FILE: write the
        marker (or, with DANDD_B200_UNION_FILES=full, build the registers from the members) -- no
        estimator run.  The marker holds p, the cardinality and the member paths; registers() rebuilds
        the union from them on demand."""
        if self.pair_union_files == "full":
            regs = self.engine.union([self.registers(mb) for mb in members])
            self._remember(path, regs)
            self._write(path, regs, p, card, leaf=False, members=list(members))
        else:
            with timing.span("write_union_markers"):
                try:
                    hllfile.write_stub(path, p, card, list(members))
                except FileNotFoundError:
                    os.makedirs(os.path.dirname(path), exist_ok=True)
                    hllfile.write_stub(path, p, card, list(members))
            self.stats["files_written"] += 1
        return float(card)

    # ------------------------------------------------------------------ dashing card
    def card_of_file(self, path: str, p: int = None) -> float:
        regs = self.registers(path)
        pp = int(regs.numel()).bit_length() - 1
        return float(self.engine.cards(regs.view(1, -1), pp)[0])

    # ------------------------------------------------------------------ kmc / kmc_tools
    def exact_count(self, fastas: Sequence[str], k: int, canon: bool) -> int:
        """Number of distinct (canonical) k-mers in the union of the FASTAs (KMC semantics)."""
        return self.exact_prefix_counts(fastas, k, canon)[-1]

    def exact_prefix_counts(self, fastas: Sequence[str], k: int, canon: bool) -> List[int]:
        self.stats["exact_calls"] += 1
        if self.exact_workers is not None:
            return self.exact_workers.counts(list(fastas), int(k), bool(canon))
        return self._exact_shard_counts(fastas, k, canon, None)

    def _exact_shard_counts(self, fastas, k, canon, shard):
        return self.engine.exact_counts([self.packed(f) for f in fastas], int(k), canon, shard=shard)

    def start_exact_workers(self):
        """Multi-rank exact mode (dandd_b200.dist.ExactWorkers): rank 0 keeps going, the others serve."""
        from dandd_b200 import dist as dd_dist
        self.exact_workers = dd_dist.ExactWorkers(self._exact_shard_counts)
        return self.exact_workers


import threading as _threading  # noqa: E402

_store = None
_store_lock = _threading.Lock()
_bound_threads = set()


def get_store():
    """The process-wide store.  Thread-safe: the command line starts it on a background thread (CUDA
    start-up takes seconds) while the main thread joins the process group."""
    global _store
    if _store is None:
        with _store_lock:
            if _store is None:
                from dandd_b200 import _startup
                if _startup.TRIM_REQUESTED:          # command-line start-up trim, see _startup.py
                    _startup.trim_torch_cuda_init()
                with timing.span("engine_start"):    # CUDA context + library load (torch itself was imported with this module)
                    _store = GpuSketchStore()
                timing.mark("engine_ready")
    ident = _threading.get_ident()
    if ident not in _bound_threads and hasattr(getattr(_store, "engine", None), "bind_thread"):
        _store.engine.bind_thread()                  # the store may have been started on another thread
        _bound_threads.add(ident)
    return _store


def set_store(store) -> None:
    """Install another store (a different device; the test-suite installs an oracle-backed double
    here to exercise the host logic on machines without a GPU)."""
    global _store
    _store = store
