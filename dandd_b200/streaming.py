"""Streaming leaf sketches: disk -> pinned ring -> H2D -> K1 -> K2, with the FASTA's blake2b name
computed from the very same chunks ("host ingest").

For a 3.1 GB assembly the kernels need ~0.2 s; what `dandd tree` waits for is reading the file
(page cache: ~1 s), hashing it for the sketch-database name (hashlib.blake2b: ~0.6-1 GB/s, one
core, inherently sequential -- reference lib/sketch_classes.py:12-18) and copying it to the GPU.
Done one after another that is ~7 s per genome; here they run as one pipeline over 64 MiB chunks:

    reader thread   file.readinto(pinned slot)                       (releases the GIL)
    hasher thread   blake2b.update(slot)                             (releases the GIL)
    feeder (caller) cudaMemcpyAsync(slot -> device buffer) on a copy stream, then on the compute
                    stream dd_pack_fasta + dd_sketch_update_sched for that chunk; two device text
                    buffers, so chunk i+1 is copied while chunk i is packed and sketched

A slot goes back to the reader once both the hasher and the copy are done with it.  Wall time is
the slowest stage (the hash), not the sum.  Gzip input is inflated first (ingest.gunzip, BGZF
member-parallel) and then fed through the same loop from memory; the digest is always that of the
FILE bytes, compressed or not, as in the reference.
"""
import hashlib
import os
import queue
import threading
import time
from typing import Optional, Sequence

import numpy as np
import torch

from . import ingest
from ._lib import DD_HIST_BINS, DD_PACK_FLAG_FASTQ, DD_PACK_FLAG_OVERFLOW, check
from .engine import Engine, FastqInput, kmask_of

CHUNK_BYTES = int(os.environ.get("DANDD_B200_STREAM_CHUNK", str(64 << 20)))
RING_SLOTS = 6
# threads that copy a file from the page cache into ring slots at the same time: one kernel-to-user copy
# runs at 6-7 GB/s, a B200 packs and sketches 20 GB/s of text
READERS = max(1, int(os.environ.get("DANDD_B200_STREAM_READERS", "3")))


class _Ring:
    """RING_SLOTS pinned host buffers; a slot is handed reader -> (hasher, feeder) -> reader."""

    def __init__(self, chunk_bytes: int, pin: bool = True):
        self.chunk = chunk_bytes
        self.slots = [torch.empty(chunk_bytes, dtype=torch.uint8) for _ in range(RING_SLOTS)]
        if pin:
            self.slots = [s.pin_memory() for s in self.slots]
        self.views = [memoryview(s.numpy()) for s in self.slots]
        self.free: "queue.Queue[int]" = queue.Queue()
        for i in range(RING_SLOTS):
            self.free.put(i)
        self.pending = [0] * RING_SLOTS     # consumers still using the slot
        self.lock = threading.Lock()

    def release(self, slot: int) -> None:
        with self.lock:
            self.pending[slot] -= 1
            done = self.pending[slot] == 0
        if done:
            self.free.put(slot)


_rings = {}


def _ring_for(device, chunk_bytes) -> _Ring:
    key = (str(device), chunk_bytes, threading.get_ident())
    if key not in _rings:
        _rings[key] = _Ring(chunk_bytes)
    return _rings[key]


def _reader(path, ring: _Ring, out_q, hash_q, stats, raw_text: Optional[bytes], stop: threading.Event):
    """Fill slots with consecutive chunks of the file (or of already inflated text) until the end, or until
    the feeder sets `stop` (it failed: the rest of the file is of no use to anyone)."""
    t_busy = 0.0
    try:
        if raw_text is None:
            fh = open(path, "rb", buffering=0)
        else:
            raw_view = memoryview(raw_text)      # slices of a view do not copy
        pos = seq = 0
        while True:
            slot = ring.free.get()
            if stop.is_set():
                ring.free.put(slot)
                break
            t0 = time.perf_counter()
            if raw_text is None:
                n = fh.readinto(ring.views[slot])
                while 0 < n < ring.chunk:           # short reads: fill the slot
                    more = fh.readinto(ring.views[slot][n:])
                    if not more:
                        break
                    n += more
            else:
                n = min(ring.chunk, len(raw_text) - pos)
                ring.views[slot][:n] = raw_view[pos:pos + n]
                pos += n
            t_busy += time.perf_counter() - t0
            if not n:
                ring.free.put(slot)
                break
            with ring.lock:
                ring.pending[slot] = 2 if hash_q is not None else 1
            if hash_q is not None:
                hash_q.put((slot, n))
            out_q.put((seq, slot, n))
            seq += 1
        if raw_text is None:
            fh.close()
    except BaseException as e:  # noqa: BLE001 -- handed to the feeder, which re-raises
        out_q.put(e)
    finally:
        stats["read_s"] = t_busy
        if hash_q is not None:
            hash_q.put(None)
        out_q.put(None)


class _Cursor:
    """Which chunk of the file the next free slot gets (shared by the positional readers)."""

    def __init__(self, size: int, chunk: int):
        self.size, self.chunk = size, chunk
        self.off = self.seq = 0
        self.busy = 0.0
        self.lock = threading.Lock()

    def claim(self):
        with self.lock:
            if self.off >= self.size:
                return None
            got = (self.seq, self.off, min(self.chunk, self.size - self.off))
            self.off += got[2]
            self.seq += 1
            return got


def _reader_at(fd: int, ring: _Ring, out_q, cursor: _Cursor, stop: threading.Event):
    """One of READERS threads filling slots with pread(): a thread first owns a slot and only then claims
    the next chunk, so the chunk the feeder is waiting for always has a buffer to land in."""
    t_busy = 0.0
    try:
        while True:
            slot = ring.free.get()
            piece = None if stop.is_set() else cursor.claim()
            if piece is None:
                ring.free.put(slot)
                break
            seq, off, n = piece
            t0 = time.perf_counter()
            got = 0
            try:
                while got < n:
                    more = os.preadv(fd, [ring.views[slot][got:n]], off + got)
                    if not more:
                        raise OSError("file shrank while it was being read")
                    got += more
            except BaseException:
                ring.free.put(slot)
                raise
            t_busy += time.perf_counter() - t0
            with ring.lock:
                ring.pending[slot] = 1
            out_q.put((seq, slot, n))
    except BaseException as e:  # noqa: BLE001 -- handed to the feeder, which re-raises
        stop.set()
        out_q.put(e)
    finally:
        with cursor.lock:
            cursor.busy += t_busy
        out_q.put(None)


def _in_order(out_q, producers: int):
    """What the readers put on out_q, chunks in file order: ("chunk", slot, n) / ("error", exc, None), and once
    every producer has finished ("orphan", slot, n) for chunks behind a gap a failed reader left."""
    waiting, want = {}, 0
    while producers:
        item = out_q.get()
        if item is None:
            producers -= 1
        elif isinstance(item, BaseException):
            yield "error", item, None
        else:
            waiting[item[0]] = item[1:]
            while want in waiting:
                yield ("chunk",) + waiting.pop(want)
                want += 1
    for seq in sorted(waiting):
        yield ("orphan",) + waiting[seq]


def _hasher(ring: _Ring, hash_q, result, stats):
    h = hashlib.blake2b()
    t_busy = 0.0
    while True:
        item = hash_q.get()
        if item is None:
            break
        slot, n = item
        t0 = time.perf_counter()
        h.update(ring.views[slot][:n])
        t_busy += time.perf_counter() - t0
        ring.release(slot)
    result["digest"] = h.hexdigest()
    stats["blake2b_s"] = t_busy


def _retire(ring: _Ring, slot: int, copied, quiet: bool = False):
    """Give a pinned slot back once the copy engine has finished reading it.  With `quiet` (an error is
    already being propagated) a failing synchronize must not mask the first error nor strand the slot."""
    try:
        copied.synchronize()
    except Exception:  # noqa: BLE001
        if not quiet:
            ring.release(slot)
            raise
    ring.release(slot)


def sketch_file(eng: Engine, path: str, ks: Sequence[int], p: int = 20, canon: bool = True,
                chunk_bytes: int = CHUNK_BYTES, out: Optional[torch.Tensor] = None, text: Optional[bytes] = None,
                want_digest: bool = True):
    """All-k HLL sketch of the FASTA at `path`, streamed.  Returns (regs [nk, 2^p] u8 on the device,
    cards numpy [nk], blake2b hex digest of the file bytes or None, stats dict of stage times).
    `text`: the (decompressed) bytes of the file if the caller already holds them (ingest.prefetch);
    they are then fed from memory and no digest is computed here.
    `want_digest=False`: the caller gets the name elsewhere (ingest hashes the file on another thread);
    the chunks are then not hashed here (a gzip file still is: its raw bytes are in hand anyway).
    Raises FastqInput if the text turns out to be FASTQ (the caller takes the whole-file detour)."""
    lib = eng.lib
    kmask = kmask_of(ks)
    nk = bin(kmask).count("1")
    m = 1 << p
    dev = eng.device
    stats = {"path": path, "bytes": os.path.getsize(path) if text is None else len(text)}
    t_start = time.perf_counter()
    magic = b""
    if text is None:
        with open(path, "rb") as fh:
            magic = fh.read(2)
    raw_text = text
    digest = {"digest": None}
    if magic == b"\x1f\x8b":                       # gzip: inflate first (the hash is of the compressed file)
        with open(path, "rb") as fh:
            raw = fh.read()
        t0 = time.perf_counter()
        digest["digest"] = hashlib.blake2b(raw).hexdigest()
        stats["blake2b_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        raw_text = ingest.gunzip(raw)
        stats["inflate_s"] = time.perf_counter() - t0
        del raw
    total = len(raw_text) if raw_text is not None else stats["bytes"]
    size = 1 << 20                                  # ring slots come in powers of two: few distinct pinned rings
    while size < min(chunk_bytes, total):
        size <<= 1
    chunk_bytes = size
    ring = _ring_for(dev, chunk_bytes)
    out_q: "queue.Queue" = queue.Queue()
    stop = threading.Event()
    hash_q = queue.Queue() if (raw_text is None and want_digest) else None
    # device side
    cb, ib = lib.dd_pack_codes_bytes(max(total, 1)), lib.dd_pack_invalid_bytes(max(total, 1))
    codes = torch.empty(cb, dtype=torch.uint8, device=dev)
    invalid = torch.empty(ib, dtype=torch.uint8, device=dev)
    state = torch.empty(32, dtype=torch.uint8, device=dev)
    d_text = [eng._buf(chunk_bytes + 256, f"stream_text{i}") for i in range(2)]
    pack_ws = eng._buf(lib.dd_pack_workspace_bytes(chunk_bytes), "stream_pack")
    sk_ws = eng._buf(lib.dd_sketch_workspace_bytes(nk, p), "stream_sketch")
    regs = out if out is not None else torch.empty((nk, m), dtype=torch.uint8, device=dev)
    hist = torch.empty((nk, DD_HIST_BINS), dtype=torch.int32, device=dev)
    cards = torch.empty(nk, dtype=torch.float64, device=dev)
    compute = torch.cuda.current_stream(dev)
    copy = eng.__dict__.setdefault("_copy_stream", torch.cuda.Stream(device=dev))
    st, cs = compute.cuda_stream, copy.cuda_stream
    check(lib.dd_pack_reset(codes.data_ptr(), cb, invalid.data_ptr(), ib, state.data_ptr(), st), "dd_pack_reset")
    check(lib.dd_sketch_begin(sk_ws.data_ptr(), sk_ws.numel(), nk, p, st), "dd_sketch_begin")
    copy.wait_stream(compute)
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    freed = [torch.cuda.Event(), torch.cuda.Event()]
    ev_gpu = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    ev_gpu[0].record(compute)
    seen = 0
    i = 0
    started = False           # first record marker found
    in_flight = []            # (slot, copied event) whose host slot is still being read by the copy engine
    err = None

    def feed(slot, n):
        """Enqueue the copy, pack and sketch of one chunk; returns True if the slot is now in flight."""
        nonlocal seen, i, started
        off = 0
        if not started:       # kseq ignores everything before the first record marker
            off = int(lib.dd_fasta_first_record_host(ring.slots[slot].data_ptr(), n))
            started = off < n
        ln = n - off
        if ln <= 0:
            return False
        b = i & 1
        if i >= 2:
            copy.wait_event(freed[b])
        with torch.cuda.stream(copy):
            d_text[b][:ln].copy_(ring.slots[slot][off:off + ln], non_blocking=True)
            copied[b].record(copy)
        in_flight.append((slot, copied[b]))
        copied[b] = torch.cuda.Event()     # a fresh event per chunk: the old one is still referenced by in_flight
        compute.wait_event(in_flight[-1][1])
        check(lib.dd_pack_fasta(d_text[b].data_ptr(), ln, codes.data_ptr(), invalid.data_ptr(), max(total, 1),
                                state.data_ptr(), pack_ws.data_ptr(), pack_ws.numel(), st), "dd_pack_fasta")
        freed[b].record(compute)
        if eng.polyt_sentinel:
            check(lib.dd_pack_polyt_sentinel(codes.data_ptr(), invalid.data_ptr(), state.data_ptr(), 0, 0, ln, st),
                  "dd_pack_polyt_sentinel")
        check(lib.dd_sketch_update_sched(codes.data_ptr(), invalid.data_ptr(), state.data_ptr(), 0, 0, ln, seen, kmask, p,
                                         int(canon), sk_ws.data_ptr(), sk_ws.numel(), st), "dd_sketch_update_sched")
        seen += ln
        i += 1
        return True

    # host side: the threads that fill the ring (started last: nothing above can fail and leave them waiting)
    fd, cursor = None, None
    if raw_text is None and hash_q is None and READERS > 1 and total > chunk_bytes:
        # a plain file whose name comes from elsewhere: several threads pread() it into the ring
        fd = os.open(path, os.O_RDONLY)
        cursor = _Cursor(total, chunk_bytes)
        producers = min(READERS, -(-total // chunk_bytes))
        threads = [threading.Thread(target=_reader_at, args=(fd, ring, out_q, cursor, stop), daemon=True) for _ in range(producers)]
    else:
        producers = 1
        threads = [threading.Thread(target=_reader, args=(path, ring, out_q, hash_q, stats, raw_text, stop), daemon=True)]
        if hash_q is not None:
            threads.append(threading.Thread(target=_hasher, args=(ring, hash_q, digest, stats), daemon=True))
    for t in threads:
        t.start()

    for kind, slot, n in _in_order(out_q, producers):   # drains the readers' queue to the end whatever happens, so that
        if kind == "error":                             # every slot finds its way back into the ring and the threads end
            err = err or slot
            stop.set()
            continue
        queued = False
        if err is None and kind == "chunk":
            before = len(in_flight)
            try:
                queued = feed(slot, n)
            except BaseException as e:  # noqa: BLE001 -- re-raised below, after the pipeline has been drained
                err = e
                stop.set()
                queued = len(in_flight) > before     # the copy was enqueued before the failure
        if not queued:
            ring.release(slot)
        while len(in_flight) > (1 if (err is None and queued) else 0):   # keep one copy in flight, give the rest back to the reader
            _retire(ring, *in_flight.pop(0), quiet=err is not None)
    for s0, ev in in_flight:
        _retire(ring, s0, ev, quiet=err is not None)
    for t in threads:
        t.join()
    if fd is not None:
        os.close(fd)
        stats["read_s"] = cursor.busy / max(1, producers)      # per reader, comparable with the single-reader figure
        stats["readers"] = producers
    if err is not None:
        raise err
    check(lib.dd_sketch_end(sk_ws.data_ptr(), sk_ws.numel(), nk, p, regs.data_ptr(), hist.data_ptr(), cards.data_ptr(), st),
          "dd_sketch_end")
    ev_gpu[1].record(compute)
    flags = int(state.cpu().numpy().view(np.uint64)[3])     # synchronises the compute stream
    if flags & DD_PACK_FLAG_OVERFLOW:
        raise RuntimeError("packed stream overflowed its capacity")
    if flags & DD_PACK_FLAG_FASTQ:
        raise FastqInput(f"{path}: FASTQ text (a line begins with '+')")
    stats["gpu_span_s"] = ev_gpu[0].elapsed_time(ev_gpu[1]) / 1e3
    stats["wall_s"] = time.perf_counter() - t_start
    stats["chunks"] = i
    return regs, cards.cpu().numpy(), digest["digest"], stats
