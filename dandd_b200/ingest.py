"""Host ingest: once the kernels are fast, `dandd tree` wall-time is reading,
inflating and naming the FASTAs.  A small thread pool does all three per file in the background --
read, gunzip if needed (zlib releases the GIL), blake2b digest (hashlib releases the GIL) -- so that
by the time a leaf is sketched its bytes and its name are already there, and file i+1.. are being
prepared while the GPU works on file i.

    prefetch(paths)          start background work for these files (idempotent)
    fasta_bytes(path)        decompressed bytes (waits for the READ only; drops the cache entry)
    drop(path)               release a prefetched file whose bytes are not needed after all
    digest(path)             blake2b hex digest of the FILE bytes (what the reference hashes, compressed
                             or not: lib/sketch_classes.py:12-18)
"""
import gzip
import hashlib
import os
import threading
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Dict, Iterable

_MAX_CACHED_BYTES = int(os.environ.get("DANDD_B200_INGEST_BYTES", str(32 << 30)))
# uncompressed files from this size on are streamed from disk by the sketch (store.STREAM_MIN_BYTES, the
# same variable): prefetching prepares their NAME only
HASH_ONLY_MIN_BYTES = int(os.environ.get("DANDD_B200_STREAM_MIN", str(32 << 20)))
_pool = None
_hash_pool = None
_lock = threading.Lock()
_lock2 = threading.Lock()
_jobs: Dict[str, "_Job"] = {}
_digests: Dict[tuple, str] = {}
_cached = 0


def _pool_get():
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=max(2, min(16, (os.cpu_count() or 4) // 2)), thread_name_prefix="dd-ingest")
    return _pool


def _hash_pool_get():
    global _hash_pool
    with _lock2:
        if _hash_pool is None:      # its own pool: a digest must not queue behind the reads of a thousand later files
            # naming is the one stage of a fresh run that only host cores can do (one sequential blake2b per
            # file, ~0.5 GB/s per busy core): it gets all of them but a few
            _hash_pool = ThreadPoolExecutor(max_workers=max(2, min(32, (os.cpu_count() or 4) - 4)), thread_name_prefix="dd-blake2b")
    return _hash_pool


def _key(path):
    st = os.stat(path)
    return (os.path.abspath(path), st.st_size, st.st_mtime_ns)


# ---- BGZF (bgzip) ------------------------------------------------------------------------------------
# A plain .gz stream can only be inflated front to back.  bgzip output -- common for genome FASTAs --
# is a series of independent <= 64 KiB gzip members whose header carries the member's size in a
# "BC" extra field, so the members can be found without inflating and inflated in parallel
# (zlib releases the GIL).  Anything that is not BGZF goes through gzip.decompress as before.
_inflate_pool = None
_BGZF_BATCH = 256          # members per task (~16 MiB of text)


def _bgzf_members(raw: bytes):
    """[(payload_begin, payload_end, isize, crc32)] of every member if `raw` is BGZF from end to end, else None."""
    out, at, n = [], 0, len(raw)
    while at < n:
        if n - at < 18 or raw[at:at + 4] != b"\x1f\x8b\x08\x04":
            return None
        xlen = int.from_bytes(raw[at + 10:at + 12], "little")
        extra, bsize, q = raw[at + 12:at + 12 + xlen], None, 0
        while q + 4 <= len(extra):
            slen = int.from_bytes(extra[q + 2:q + 4], "little")
            if extra[q:q + 2] == b"BC" and slen == 2:
                bsize = int.from_bytes(extra[q + 4:q + 6], "little") + 1
            q += 4 + slen
        if bsize is None or at + bsize > n or bsize < 12 + xlen + 8:
            return None
        out.append((at + 12 + xlen, at + bsize - 8, int.from_bytes(raw[at + bsize - 4:at + bsize], "little"),
                    int.from_bytes(raw[at + bsize - 8:at + bsize - 4], "little")))
        at += bsize
    return out


def _inflate_batch(raw, members):
    import zlib
    parts = []
    for begin, end, isize, crc in members:
        data = zlib.decompress(raw[begin:end], wbits=-15, bufsize=max(isize, 1))
        if len(data) != isize or zlib.crc32(data) != crc:
            raise ValueError("corrupt BGZF member (size or CRC-32 does not match its trailer)")
        parts.append(data)
    return b"".join(parts)


def gunzip(raw: bytes) -> bytes:
    """gzip -> bytes; BGZF input is inflated by a pool of threads."""
    global _inflate_pool
    members = _bgzf_members(raw) if raw[:4] == b"\x1f\x8b\x08\x04" else None
    if not members or len(members) <= _BGZF_BATCH:
        return gzip.decompress(raw)
    if _inflate_pool is None:      # its own pool: _load itself runs on the ingest pool
        _inflate_pool = ThreadPoolExecutor(max_workers=max(2, min(32, os.cpu_count() or 4)), thread_name_prefix="dd-inflate")
    view = memoryview(raw)
    jobs = [_inflate_pool.submit(_inflate_batch, view, members[i:i + _BGZF_BATCH]) for i in range(0, len(members), _BGZF_BATCH)]
    return b"".join(j.result() for j in jobs)


def _read(path):
    from . import timing
    with timing.span("prefetch_read"):
        with open(path, "rb") as fh:
            return fh.read()


def _hash(raw):
    from . import timing
    with timing.span("prefetch_blake2b"):
        return hashlib.blake2b(raw).hexdigest()


def _hash_file(path, piece=128 << 20):
    """blake2b of a file without holding its bytes: the page cache is mapped and hashed in large
    pieces.  Large, because every hashlib call gives up the GIL and has to win it back afterwards --
    with 1 MiB reads and a main thread busy importing torch that wait, not the hashing, set the pace
    (measured: 4 files x 512 MiB, busy main thread: 10.6 s with 1 MiB reads, 1.6 s mapped)."""
    import mmap
    from . import timing
    with timing.span("prefetch_blake2b"):
        h = hashlib.blake2b()
        with open(path, "rb", buffering=0) as fh:
            try:
                mapped = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            except (ValueError, OSError):          # empty file, or something that cannot be mapped
                mapped = None
            if mapped is None:
                buf = bytearray(min(piece, 64 << 20))
                while True:
                    n = fh.readinto(buf)
                    if not n:
                        break
                    h.update(memoryview(buf)[:n])
                return h.hexdigest()
            try:
                if hasattr(mapped, "madvise") and hasattr(mmap, "MADV_SEQUENTIAL"):
                    mapped.madvise(mmap.MADV_SEQUENTIAL)
                view = memoryview(mapped)
                try:
                    for at in range(0, len(mapped), piece):
                        h.update(view[at:at + piece])
                finally:
                    view.release()
            finally:
                mapped.close()
        return h.hexdigest()


def _is_gzip(path) -> bool:
    try:
        with open(path, "rb") as fh:
            return fh.read(2) == b"\x1f\x8b"
    except OSError:
        return False


def _text(raw):
    return gunzip(raw) if raw[:2] == b"\x1f\x8b" else raw


def _load(path):
    """(decompressed text, blake2b hex digest of the file bytes), in the calling thread."""
    raw = _read(path)
    return (gunzip(raw) if raw[:2] == b"\x1f\x8b" else raw), hashlib.blake2b(raw).hexdigest()


class _Job:
    """Background work for one file: the text becomes available as soon as the file is read (and
    inflated), the digest -- the slow part, one core at ~0.6-1 GB/s -- on its own thread, so a sketch
    can start before the name is known.

    No pool task ever waits for another one: the read task hands the bytes to the hashing pool and
    goes on to inflate them.  (Tasks that block on a sibling's future starve a bounded pool: with 8
    workers and 8 files queued as read/text/hash triples only 3 files made progress at a time, which
    tripled the naming time of the 8 x 3.1 GB run.)"""

    def __init__(self, path, pool, hash_only=False, want_digest=True):
        # want_digest=False: bytes only (callers that never name the file: helpers/allpairs.py)
        self.digest: Future = Future() if want_digest else None
        if hash_only:
            # a large uncompressed file: nobody needs its bytes in one piece (streaming.py feeds the GPU from
            # the file itself, through the page cache), so only the name is prepared -- hashed in 1 MiB
            # reads, without materialising a multi-GB bytes object (whose page faults slowed everything
            # else in the process down, CUDA start-up included)
            self.text = None
            _hash_pool_get().submit(self._run_hash_file, path)
            return
        self.text: Future = Future()
        pool.submit(self._run, path)

    def _run(self, path):
        try:
            raw = _read(path)
        except BaseException as err:  # noqa: BLE001 -- delivered to whoever asks for the text or the digest
            self.text.set_exception(err)
            if self.digest is not None:
                self.digest.set_exception(err)
            return
        if self.digest is not None:
            _hash_pool_get().submit(self._run_hash, raw)
        try:
            self.text.set_result(_text(raw))
        except BaseException as err:  # noqa: BLE001
            self.text.set_exception(err)

    def _run_hash_file(self, path):
        try:
            self.digest.set_result(_hash_file(path))
        except BaseException as err:  # noqa: BLE001
            self.digest.set_exception(err)

    def _run_hash(self, raw):
        try:
            self.digest.set_result(_hash(raw))
        except BaseException as err:  # noqa: BLE001
            self.digest.set_exception(err)


def prefetch(paths: Iterable[str], want_digest: bool = True) -> None:
    """Start reading (and inflating) these files in the background; with want_digest also their blake2b
    names.  Idempotent per path."""
    global _cached
    with _lock:
        for p in paths:
            if p in _jobs or not os.path.isfile(p):
                continue
            size = os.path.getsize(p)
            if size >= HASH_ONLY_MIN_BYTES and not _is_gzip(p):
                if want_digest:
                    _jobs[p] = _Job(p, None, hash_only=True)     # holds no bytes: not counted against the cache
                continue
            if _cached + size > _MAX_CACHED_BYTES:
                break                         # the rest is loaded on demand
            _cached += size
            _jobs[p] = _Job(p, _pool_get(), want_digest=want_digest)


def _take(path):
    """(text, digest-or-None).  A prefetched file gives its text as soon as it is read; its digest is
    remembered as a future so that digest() can wait for it separately."""
    global _cached
    with _lock:
        job = _jobs.pop(path, None)
    if job is None:
        return _load(path)
    if job.text is None:                      # name-only job: read now, the digest keeps coming from the job
        with _lock:
            _digest_jobs[_key(path)] = job.digest
        return _text(_read(path)), None
    text = job.text.result()
    with _lock:
        _cached = max(0, _cached - os.path.getsize(path))
        if job.digest is not None:
            _digest_jobs[_key(path)] = job.digest
    return text, None


_digest_jobs: Dict[tuple, Future] = {}


def is_prefetched(path: str) -> bool:
    """True if the file's BYTES are (being) held in memory for fasta_bytes()."""
    with _lock:
        job = _jobs.get(path)
        return job is not None and job.text is not None


def has_digest(path: str) -> bool:
    """True if the file's digest is known or being computed in the background (a streaming reader then
    need not hash the chunks it reads)."""
    try:
        k = _key(path)
    except OSError:
        return False
    with _lock:
        job = _jobs.get(path)
        return (job is not None and job.digest is not None) or k in _digest_jobs or k in _digests


def fasta_bytes(path: str) -> bytes:
    text, dig = _take(path)
    if dig is not None:
        _digests[_key(path)] = dig
    return text


def set_digest(path: str, hexdigest: str) -> None:
    """Record a digest computed elsewhere (the streaming sketch hashes the chunks it reads)."""
    _digests[_key(path)] = hexdigest


def drop(path: str) -> None:
    """Forget the cached bytes of a prefetched file that turned out not to be needed (its sketches
    were already in the database); the digest, once computed, is kept."""
    global _cached
    with _lock:
        job = _jobs.pop(path, None)
        if job is not None:
            if job.text is not None:
                _cached = max(0, _cached - os.path.getsize(path))
            if job.digest is not None:
                _digest_jobs[_key(path)] = job.digest


def drop_all() -> None:
    with _lock:
        paths = list(_jobs)
    for p in paths:
        drop(p)


def digest(path: str) -> str:
    k = _key(path)
    d = _digests.get(k)
    if d is None:
        with _lock:
            job = _jobs.get(path)
            fut = job.digest if job is not None else _digest_jobs.get(k)
        if fut is not None:
            from . import timing
            with timing.span("digest_wait"):
                d = fut.result()              # (the text stays cached for the sketch that follows)
            _digest_jobs.pop(k, None)
        else:
            d = _hash_file(path)
        _digests[k] = d
    return d
