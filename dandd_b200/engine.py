"""Engine: the batched GPU operations behind the drop-in sketch objects.

PyTorch is used only to own device memory and streams; every operation is one or a few calls
into the C ABI (include/dandd_b200.h) on torch's current stream, so `torch.cuda.Event` timing
and stream semantics work as usual.  One Engine per process / per GPU."""
import ctypes as C
import os
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import DD_HIST_BINS, DD_PACK_FLAG_FASTQ, DD_PACK_FLAG_OVERFLOW, DandDError, check


class FastqInput(DandDError):
    """The packed text holds a line beginning with '+': FASTQ.  The caller rewrites the text with
    Engine.fastq_to_fasta (kseq's walk of the records, on the host) and packs that instead."""


def kmask_of(ks: Iterable[int]) -> int:
    m = 0
    for k in ks:
        k = int(k)
        if not 1 <= k <= 32:
            raise ValueError(f"k={k} outside 1..32 (HLL estimation supports k<=32; reference README.md:82)")
        m |= 1 << (k - 1)
    if m == 0:
        raise ValueError("no k values given")
    return m


def ks_of(kmask: int) -> List[int]:
    return [k for k in range(1, 33) if (kmask >> (k - 1)) & 1]


@dataclass
class PackedSeq:
    """A FASTA packed into the 2-bit symbol stream (layout: include/dandd_b200.h, K1)."""
    codes: torch.Tensor       # uint8 storage holding the u32 code words
    invalid: torch.Tensor     # uint8 storage holding the u32 break-bit words
    state: torch.Tensor       # 32 bytes: dd_pack_state
    cap_symbols: int
    text_bytes: int
    _nsym: Optional[int] = None

    @property
    def nsym(self) -> int:
        """Number of symbols (reads the device state: synchronises)."""
        if self._nsym is None:
            st = self.state.cpu().numpy().view(np.uint64)
            if int(st[3]) & DD_PACK_FLAG_OVERFLOW:
                raise DandDError("packed stream overflowed its capacity")
            if int(st[3]) & DD_PACK_FLAG_FASTQ:
                raise FastqInput("FASTQ text (a line begins with '+') must go through Engine.fastq_to_fasta first")
            self._nsym = int(st[0])
        return self._nsym

    def check(self) -> "PackedSeq":
        """Raise if the packer flagged the text (overflow, FASTQ).  Synchronises."""
        self.nsym  # noqa: B018
        return self


class Engine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise DandDError("no CUDA device visible to torch; dandd_b200 has no CPU fallback")
        check(self.lib.dd_init(int(device)), "dd_init")
        self.device = torch.device("cuda", int(device))
        torch.cuda.set_device(self.device)
        sm, maj, mnr, l2, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t(), C.c_size_t()
        check(self.lib.dd_device_info(int(device), C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(l2), C.byref(mem)))
        self.sm_count, self.cc, self.l2_bytes, self.total_mem = sm.value, (maj.value, mnr.value), l2.value, mem.value
        self._ws = {}
        # SURVEY.md A.6 (UNVERIFIED encoder quirk): opt-in emulation, for every entry point of this engine
        self.polyt_sentinel = os.environ.get("DANDD_B200_POLYT_SENTINEL", "0") not in ("", "0")
        check(self.lib.dd_set_option(b"polyt_sentinel", int(self.polyt_sentinel)), "dd_set_option")

    # ---- plumbing ---------------------------------------------------------------------------
    def bind_thread(self) -> None:
        """Make this engine's GPU the calling thread's current device.  The CUDA runtime's current
        device is per host thread: an engine created on one thread (the command line starts it in the
        background) must be bound again on every other thread that launches through it."""
        torch.cuda.set_device(self.device)

    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _buf(self, nbytes: int, tag: str) -> torch.Tensor:
        """Grow-only scratch tensors keyed by role (uint8)."""
        t = self._ws.get(tag)
        if t is None or t.numel() < nbytes:
            t = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws[tag] = t
        return t

    def _dev_u8(self, data) -> torch.Tensor:
        if isinstance(data, torch.Tensor):
            return data.to(self.device, dtype=torch.uint8).contiguous().view(-1)
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.uint8)
        arr = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        if arr.size == 0:
            return torch.empty(0, dtype=torch.uint8, device=self.device)
        return torch.from_numpy(arr.copy() if not arr.flags.writeable else arr).to(self.device, non_blocking=False)

    # ---- K1 ---------------------------------------------------------------------------------
    @staticmethod
    def skip_preamble(text) -> int:
        """Offset of the first '>' or '@' (kseq ignores everything before the first record marker,
        SURVEY.md A.1); len if there is none."""
        if isinstance(text, torch.Tensor) and text.is_cuda:
            n = int(text.numel())
            block = 1 << 26
            for off in range(0, n, block):            # the marker is almost always in the first block
                part = text[off:off + block]
                hit = ((part == 62) | (part == 64)).nonzero()
                if hit.numel():
                    return off + int(hit[0])
            return n
        if isinstance(text, torch.Tensor):
            ptr, n = text.data_ptr(), int(text.numel())
        else:
            arr = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else \
                np.ascontiguousarray(text, dtype=np.uint8).reshape(-1)
            ptr, n = arr.ctypes.data, int(arr.size)
        return int(_lib.load().dd_fasta_first_record_host(ptr, n)) if n else 0

    @staticmethod
    def fastq_to_fasta(text) -> bytes:
        """kseq_read()'s walk over FASTA/FASTQ text, written back as plain FASTA (host side, no GPU):
        what the packer must be given when it reports DD_PACK_FLAG_FASTQ."""
        arr = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else \
            np.ascontiguousarray(text, dtype=np.uint8).reshape(-1)
        out = np.empty(arr.size + 16, dtype=np.uint8)
        n = _lib.load().dd_fastq_to_fasta_host(arr.ctypes.data, arr.size, out.ctypes.data) if arr.size else 0
        return out[:n].tobytes()

    def pack(self, text, chunk_bytes: Optional[int] = None, start: Optional[int] = None, ws_tag: str = "pack") -> PackedSeq:
        """FASTA text (bytes / numpy / torch uint8) -> PackedSeq.  `chunk_bytes` packs in several
        chunks through the carried state (used by the tests to exercise streaming).  `start` = offset
        of the first '>' if the caller knows it (device-resident text is otherwise scanned for it)."""
        if start is None and not isinstance(text, torch.Tensor):
            start = self.skip_preamble(text)          # on the host, before the upload
        d_text = self._dev_u8(text)
        if start is None:
            start = self.skip_preamble(d_text) if d_text.numel() else 0
        off = min(int(start), int(d_text.numel()))
        n = int(d_text.numel()) - off
        if off % 16 and n > 0:
            d_text = d_text[off:].clone()   # keep the 128-bit loads aligned
        elif off:
            d_text = d_text[off:]
        cb, ib = self.lib.dd_pack_codes_bytes(max(n, 1)), self.lib.dd_pack_invalid_bytes(max(n, 1))
        codes = torch.empty(cb, dtype=torch.uint8, device=self.device)
        invalid = torch.empty(ib, dtype=torch.uint8, device=self.device)
        state = torch.empty(32, dtype=torch.uint8, device=self.device)
        st = self.stream
        check(self.lib.dd_pack_reset(codes.data_ptr(), cb, invalid.data_ptr(), ib, state.data_ptr(), st), "dd_pack_reset")
        chunk = int(chunk_bytes) if chunk_bytes else max(n, 1)
        wsb = self.lib.dd_pack_workspace_bytes(min(chunk, max(n, 1)))
        ws = self._buf(wsb, ws_tag)   # one scratch buffer per stream in flight
        pos = 0
        while pos < n:
            ln = min(chunk, n - pos)
            part = d_text[pos:pos + ln]
            if part.data_ptr() % 16:
                part = part.clone()
            check(self.lib.dd_pack_fasta(part.data_ptr(), ln, codes.data_ptr(), invalid.data_ptr(), max(n, 1),
                                         state.data_ptr(), ws.data_ptr(), ws.numel(), st), "dd_pack_fasta")
            if self.polyt_sentinel:
                check(self.lib.dd_pack_polyt_sentinel(codes.data_ptr(), invalid.data_ptr(), state.data_ptr(), 0, 0, ln, st),
                      "dd_pack_polyt_sentinel")
            pos += ln
        return PackedSeq(codes, invalid, state, max(n, 1), n)

    # ---- K2 (+K4 for the leaf cardinalities) ------------------------------------------------------
    def sketch(self, seq: PackedSeq, ks: Sequence[int], p: int = 20, canon: bool = True,
               out: Optional[torch.Tensor] = None, ranges=None, floor_every: Optional[int] = None,
               hist_out: Optional[torch.Tensor] = None, ws_tag: str = "sketch"):
        """All-k HLL sketch of one packed sequence.  Returns (regs [nk, 2^p] uint8, cards [nk] f64),
        both on the device.  `ranges` (list of (begin, end) symbol ranges) forces chunked updates.
        With `hist_out` ([nk, 64] int32) the register histograms go there and the estimator is left to
        a later batched mle() call (cards is then None): one MLE launch per batch, not per genome."""
        kmask = kmask_of(ks)
        nk = bin(kmask).count("1")
        m = 1 << p
        wsb = self.lib.dd_sketch_workspace_bytes(nk, p)
        ws = self._buf(wsb, ws_tag)
        st = self.stream
        regs = out if out is not None else torch.empty((nk, m), dtype=torch.uint8, device=self.device)
        assert regs.is_contiguous() and regs.numel() == nk * m
        defer = hist_out is not None
        hist = hist_out if defer else torch.empty((nk, DD_HIST_BINS), dtype=torch.int32, device=self.device)
        assert hist.is_contiguous() and hist.numel() == nk * DD_HIST_BINS
        cards = None if defer else torch.empty(nk, dtype=torch.float64, device=self.device)
        check(self.lib.dd_sketch_begin(ws.data_ptr(), ws.numel(), nk, p, st), "dd_sketch_begin")
        if ranges is None and floor_every:
            n = seq.nsym                      # needs the symbol count on the host: one sync
            ranges = [(b, min(n, b + floor_every)) for b in range(0, n, floor_every)]
        if ranges is None:
            # whole stream in one launch; the symbol count stays on the device (no host sync)
            self._update_from_state(seq, kmask, p, canon, ws, st)
        else:
            seen = 0
            for (b, e) in ranges:
                check(self.lib.dd_sketch_update_range(seq.codes.data_ptr(), seq.invalid.data_ptr(), int(b), int(e), kmask,
                                                      p, int(canon), ws.data_ptr(), ws.numel(), st), "dd_sketch_update_range")
                seen += e - b
                if floor_every and seen >= (16 << p):
                    check(self.lib.dd_sketch_refresh_floor(ws.data_ptr(), ws.numel(), kmask, p, st), "dd_sketch_refresh_floor")
        check(self.lib.dd_sketch_end(ws.data_ptr(), ws.numel(), nk, p, regs.data_ptr(), hist.data_ptr(),
                                     None if defer else cards.data_ptr(), st), "dd_sketch_end")
        return regs, cards

    def mle(self, hist: torch.Tensor, p: int) -> torch.Tensor:
        """Ertl-MLE cardinalities from register histograms [..., 64] (int32, device)."""
        flat = hist.contiguous().view(-1, DD_HIST_BINS)
        out = torch.empty(flat.shape[0], dtype=torch.float64, device=self.device)
        check(self.lib.dd_mle_from_hist(flat.data_ptr(), flat.shape[0], p, out.data_ptr(), self.stream), "dd_mle_from_hist")
        return out.view(hist.shape[:-1])

    def _update_from_state(self, seq, kmask, p, canon, ws, st):
        # the pack state says [prev_nsym, nsym); for a whole-stream sketch rewind prev_nsym to 0
        state = seq.state.clone()
        state.view(torch.int64)[1] = 0
        # (dd_sketch_update_sched == dd_sketch_update for streams shorter than 16 x 2^p symbols; longer ones
        # are cut at the floor schedule's boundaries with a refresh at each)
        check(self.lib.dd_sketch_update_sched(seq.codes.data_ptr(), seq.invalid.data_ptr(), state.data_ptr(), 0, 0,
                                              seq.text_bytes, 0, kmask, p, int(canon), ws.data_ptr(), ws.numel(), st),
              "dd_sketch_update_sched")
        self._keep = state  # keep alive until the stream has consumed it

    def sketch_fasta_host(self, text: bytes, ks: Sequence[int], p: int = 20, canon: bool = True,
                          want_regs: bool = True, pinned_regs: Optional[torch.Tensor] = None,
                          out_dev: Optional[torch.Tensor] = None, cards_out: Optional[torch.Tensor] = None,
                          ws_tag: str = "host", sync: bool = True):
        """The host-buffer C-ABI path: FASTA bytes in host memory -> (regs numpy [nk, 2^p] or None,
        cards numpy [nk]).  H2D, pack, sketch, estimate, D2H all inside the one call.  `out_dev`
        (uint8 [nk, 2^p] on the device) additionally keeps the registers resident in HBM.
        sync=False enqueues on the current stream and returns (dd_sketch_fasta_host_async): give every
        stream in flight its own `ws_tag` and a pinned `cards_out` (float64 [nk]), and synchronise
        the stream before reading the results."""
        kmask = kmask_of(ks)
        nk = bin(kmask).count("1")
        m = 1 << p
        if isinstance(text, torch.Tensor):
            h_ptr, n = text.data_ptr(), text.numel()
        else:
            arr = np.frombuffer(text, dtype=np.uint8)
            h_ptr, n = arr.ctypes.data, arr.size
        wsb = self.lib.dd_sketch_fasta_host_workspace_bytes(n, nk, p)
        ws = self._buf(wsb, ws_tag)
        if cards_out is not None:
            assert cards_out.dtype == torch.float64 and cards_out.numel() == nk and cards_out.is_contiguous()
            cards, c_ptr = cards_out, cards_out.data_ptr()
        else:
            cards = np.empty(nk, dtype=np.float64)
            c_ptr = cards.ctypes.data
        regs = None
        r_ptr = None
        if want_regs:
            if pinned_regs is not None:
                regs, r_ptr = pinned_regs, pinned_regs.data_ptr()
            else:
                regs = np.empty((nk, m), dtype=np.uint8)
                r_ptr = regs.ctypes.data
        d_ptr = None
        if out_dev is not None:
            assert out_dev.is_contiguous() and out_dev.numel() == nk * m and out_dev.dtype == torch.uint8
            d_ptr = out_dev.data_ptr()
        fn = self.lib.dd_sketch_fasta_host if sync else self.lib.dd_sketch_fasta_host_async
        check(fn(h_ptr, n, kmask, p, int(canon), r_ptr, c_ptr, d_ptr, ws.data_ptr(), ws.numel(), self.stream),
              "dd_sketch_fasta_host")
        return regs, cards

    # ---- K4 ---------------------------------------------------------------------------------
    def cards(self, regs: torch.Tensor, p: int) -> torch.Tensor:
        """Ertl-MLE cardinality of each sketch in regs [..., 2^p] (uint8, device)."""
        m = 1 << p
        flat = regs.contiguous().view(-1, m)
        nsk = flat.shape[0]
        hist = torch.empty((nsk, DD_HIST_BINS), dtype=torch.int32, device=self.device)
        out = torch.empty(nsk, dtype=torch.float64, device=self.device)
        check(self.lib.dd_card_ertl_mle(flat.data_ptr(), nsk, p, out.data_ptr(), hist.data_ptr(), self.stream), "dd_card_ertl_mle")
        return out.view(regs.shape[:-1])

    # ---- K3 ---------------------------------------------------------------------------------
    def union(self, sketches: Sequence[torch.Tensor]) -> torch.Tensor:
        """Register-wise max of equally sized uint8 device tensors (`dashing union`)."""
        n = sketches[0].numel()
        keep = [s.contiguous() for s in sketches]
        ptrs = torch.tensor([s.data_ptr() for s in keep], dtype=torch.int64).to(self.device)
        out = torch.empty_like(keep[0])
        check(self.lib.dd_union_max(ptrs.data_ptr(), len(keep), n, out.data_ptr(), self.stream), "dd_union_max")
        self._keep = (keep, ptrs)
        return out

    def prefix_union_cards(self, regs: torch.Tensor, orders, p: int, final_only: bool = False,
                           materialize: bool = False):
        """regs [n_genomes, nk, 2^p]; orders [n_ord, n_steps] genome indices (-1 = skip).
        Returns cards [n_ord, n_steps or 1, nk] f64 (and the unions if materialize)."""
        n_g, nk, m = regs.shape
        assert m == 1 << p and regs.is_contiguous()
        order = torch.as_tensor(np.asarray(orders, dtype=np.int32)).to(self.device).contiguous()
        n_ord, n_steps = order.shape
        osteps = 1 if final_only else n_steps
        hist = torch.empty((n_ord, osteps, nk, DD_HIST_BINS), dtype=torch.int32, device=self.device)
        cards = torch.empty((n_ord, osteps, nk), dtype=torch.float64, device=self.device)
        unions = torch.empty((n_ord, osteps, nk, m), dtype=torch.uint8, device=self.device) if materialize else None
        ws = None
        if not materialize and p >= 12:   # scratch for the bit-plane kernel (transposed sketches + identical-prefix table)
            ws = self._buf(self.lib.dd_prefix_union_workspace_bytes(n_ord, n_steps, n_g, nk, p), "prefix")
        check(self.lib.dd_prefix_union_card(regs.data_ptr(), order.data_ptr(), n_ord, n_steps, n_g, nk, p, int(final_only),
                                            cards.data_ptr(), hist.data_ptr(), unions.data_ptr() if materialize else None,
                                            ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                                            self.stream), "dd_prefix_union_card")
        self._keep = order
        return (cards, unions) if materialize else cards

    # ---- bit planes (K3/K6 on sketches that are unioned many times) -------------------------------
    def to_planes(self, regs: torch.Tensor, p: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """regs [..., 2^p] uint8 -> bit planes [n_sketches, 6, 2^p/32] int32 (dd_to_planes), once;
        prefix_union_cards_planes / pairwise_cards(planes=...) then never transpose again."""
        m = 1 << p
        flat = regs.contiguous().view(-1, m)
        nsk = flat.shape[0]
        if out is None:
            out = torch.empty((nsk, 6, m // 32), dtype=torch.int32, device=self.device)
        assert out.is_contiguous() and out.numel() * 4 == self.lib.dd_planes_bytes(nsk, p)
        check(self.lib.dd_to_planes(flat.data_ptr(), nsk, p, out.data_ptr(), self.stream), "dd_to_planes")
        return out

    def prefix_union_cards_planes(self, planes: torch.Tensor, n_genomes: int, nk: int, orders, p: int,
                                  final_only: bool = False) -> torch.Tensor:
        """prefix_union_cards on sketches already transposed by to_planes ([n_genomes * nk, 6, 2^p/32])."""
        order = torch.as_tensor(np.asarray(orders, dtype=np.int32)).to(self.device).contiguous()
        n_ord, n_steps = order.shape
        osteps = 1 if final_only else n_steps
        hist = torch.empty((n_ord, osteps, nk, DD_HIST_BINS), dtype=torch.int32, device=self.device)
        cards = torch.empty((n_ord, osteps, nk), dtype=torch.float64, device=self.device)
        ws = self._buf(self.lib.dd_prefix_union_workspace_bytes(n_ord, n_steps, 0, 0, p), "prefix_dedup")
        check(self.lib.dd_prefix_union_card_planes(planes.data_ptr(), order.data_ptr(), n_ord, n_steps, n_genomes, nk, p,
                                                   int(final_only), cards.data_ptr(), hist.data_ptr(), ws.data_ptr(),
                                                   ws.numel(), self.stream), "dd_prefix_union_card_planes")
        self._keep = order
        return cards

    def union_sets(self, member_ptrs, p: int, final_only: bool = True, materialize: bool = False):
        """member_ptrs: int64 [n_sets, n_steps] device addresses of 2^p-byte sketches (0 = skip).
        Returns cards [n_sets, n_steps or 1] f64 (and unions [n_sets, n_steps or 1, 2^p] if asked).
        The caller keeps the member tensors alive until the stream has run."""
        ptrs = torch.as_tensor(np.ascontiguousarray(member_ptrs, dtype=np.int64)).to(self.device)
        n_sets, n_steps = ptrs.shape
        osteps = 1 if final_only else n_steps
        m = 1 << p
        hist = torch.empty((n_sets, osteps, DD_HIST_BINS), dtype=torch.int32, device=self.device)
        cards = torch.empty((n_sets, osteps), dtype=torch.float64, device=self.device)
        unions = torch.empty((n_sets, osteps, m), dtype=torch.uint8, device=self.device) if materialize else None
        check(self.lib.dd_union_sets_card(ptrs.data_ptr(), n_sets, n_steps, p, int(final_only), cards.data_ptr(),
                                          hist.data_ptr(), unions.data_ptr() if materialize else None, self.stream),
              "dd_union_sets_card")
        self._keep = ptrs
        return (cards, unions) if materialize else cards

    def pairwise_cards(self, regs: Optional[torch.Tensor], pairs, p: int, planes: Optional[torch.Tensor] = None,
                       n_genomes: Optional[int] = None, nk: Optional[int] = None,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """card(A u B) for every listed pair and every k: [n_pairs, nk] f64.  Either `regs`
        [n_genomes, nk, 2^p] (transposed into scratch on every call) or `planes` from to_planes()
        with n_genomes / nk given (tiles of a big pair matrix share one transpose)."""
        if isinstance(pairs, torch.Tensor):
            pr = pairs.to(self.device, dtype=torch.int32).contiguous().view(-1, 2)
        else:
            pr = torch.as_tensor(np.asarray(pairs, dtype=np.int32).reshape(-1, 2)).to(self.device).contiguous()
        n_pairs = pr.shape[0]
        if planes is None:
            n_genomes, nk, m = regs.shape
        hist = self._buf(n_pairs * nk * DD_HIST_BINS * 4, "pair_hist")
        cards = out if out is not None else torch.empty((n_pairs, nk), dtype=torch.float64, device=self.device)
        assert cards.is_contiguous() and cards.numel() == n_pairs * nk and cards.dtype == torch.float64
        if planes is not None:
            check(self.lib.dd_pairwise_union_card_planes(planes.data_ptr(), n_genomes, nk, p, pr.data_ptr(), n_pairs,
                                                         cards.data_ptr(), hist.data_ptr(), self.stream),
                  "dd_pairwise_union_card_planes")
        else:
            ws = self._buf(self.lib.dd_prefix_union_workspace_bytes(0, 0, n_genomes, nk, p), "prefix") if p >= 12 else None
            check(self.lib.dd_pairwise_union_card(regs.data_ptr(), n_genomes, nk, p, pr.data_ptr(), n_pairs, cards.data_ptr(),
                                                  hist.data_ptr(), ws.data_ptr() if ws is not None else None,
                                                  ws.numel() if ws is not None else 0, self.stream), "dd_pairwise_union_card")
        self._keep = pr
        return cards

    # ---- K5 ---------------------------------------------------------------------------------
    def exact_counts(self, seqs: Sequence[PackedSeq], k: int, canon: bool = True, capacity: Optional[int] = None,
                     shard: Optional[tuple] = None) -> List[int]:
        """Insert the sequences one after another into one k-mer set and return the number of
        distinct (canonical) k-mers after each: [|S1|, |S1 u S2|, ...] (exact, KMC semantics).
        shard=(rank, world): only the k-mers of this rank's key range are stored and counted; the
        ranks' results add up to the unsharded counts (dandd_b200.dist.sum_counts)."""
        nsyms = [s.nsym for s in seqs]
        srank, sworld = (int(shard[0]), int(shard[1])) if shard is not None else (0, 1)
        if capacity is None:
            capacity = 1024
            while capacity < 2 * max(1, sum(nsyms)) // sworld + 1024:
                capacity *= 2
        wsb = self.lib.dd_exact_workspace_bytes(k, capacity)
        ws = self._buf(wsb, "exact")
        st = self.stream
        counts = torch.zeros(len(seqs), dtype=torch.int64, device=self.device)
        check(self.lib.dd_exact_begin(ws.data_ptr(), ws.numel(), k, capacity, st), "dd_exact_begin")
        for i, (s, n) in enumerate(zip(seqs, nsyms)):
            check(self.lib.dd_exact_insert_shard(s.codes.data_ptr(), s.invalid.data_ptr(), 0, n, k, int(canon), ws.data_ptr(),
                                                 ws.numel(), capacity, srank, sworld, st), "dd_exact_insert_shard")
            check(self.lib.dd_exact_count(ws.data_ptr(), ws.numel(), k, capacity, counts[i:].data_ptr(), st), "dd_exact_count")
        out = counts.cpu().numpy().astype(np.uint64)
        if (out == np.uint64(0xFFFFFFFFFFFFFFFF)).any():
            raise DandDError("exact k-mer table overflowed; pass a larger capacity")
        return [int(v) for v in out]


_default_engine = None


def get_engine() -> Engine:
    """Process-wide engine on LOCAL_RANK's GPU (torchrun) or device 0."""
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_engine


def set_engine(engine) -> None:
    """Install a different engine object (another device; tests install an oracle-backed double
    here to exercise the host logic on machines without a GPU)."""
    global _default_engine
    _default_engine = engine
