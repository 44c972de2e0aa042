"""Host ingest: once the kernels are fast, `dandd tree` wall-time is reading,
inflating and naming the FASTAs.  A small thread pool does all three per file in the background --
read, gunzip if needed (zlib releases the GIL), blake2b digest (hashlib releases the GIL) -- so that
by the time a leaf is sketched its bytes and its name are already there, and file i+1.. are being
prepared while the GPU works on file i.

    prefetch(paths)          start background work for these files (idempotent)
    fasta_bytes(path)        decompressed bytes (waits for the worker if needed; drops the cache entry)
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
_pool = None
_lock = threading.Lock()
_jobs: Dict[str, Future] = {}
_digests: Dict[tuple, str] = {}
_cached = 0


def _pool_get():
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=max(2, min(16, (os.cpu_count() or 4) // 2)), thread_name_prefix="dd-ingest")
    return _pool


def _key(path):
    st = os.stat(path)
    return (os.path.abspath(path), st.st_size, st.st_mtime_ns)


def _load(path):
    with open(path, "rb") as fh:
        raw = fh.read()
    dig = hashlib.blake2b(raw).hexdigest()
    text = gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw
    return text, dig


def prefetch(paths: Iterable[str]) -> None:
    global _cached
    with _lock:
        for p in paths:
            if p in _jobs or not os.path.isfile(p):
                continue
            size = os.path.getsize(p)
            if _cached + size > _MAX_CACHED_BYTES:
                break                         # the rest is loaded on demand
            _cached += size
            _jobs[p] = _pool_get().submit(_load, p)


def _take(path):
    global _cached
    with _lock:
        job = _jobs.pop(path, None)
    if job is None:
        return _load(path)
    text, dig = job.result()
    with _lock:
        _cached = max(0, _cached - os.path.getsize(path))
    return text, dig


def fasta_bytes(path: str) -> bytes:
    text, dig = _take(path)
    _digests[_key(path)] = dig
    return text


def digest(path: str) -> str:
    k = _key(path)
    d = _digests.get(k)
    if d is None:
        with _lock:
            job = _jobs.get(path)
        if job is not None:
            d = job.result()[1]               # keep the bytes cached for the sketch that follows
        else:
            h = hashlib.blake2b()
            with open(path, "rb") as fh:
                for block in iter(lambda: fh.read(1 << 20), b""):
                    h.update(block)
            d = h.hexdigest()
        _digests[k] = d
    return d
