"""Host ingest: plain gzip, BGZF (bgzip) inflated member-parallel, digests of the file bytes."""
import gzip
import hashlib
import struct
import zlib

import numpy as np
import pytest

from dandd_b200 import ingest
from tests.util import random_bases, to_fasta


def bgzf_compress(data: bytes, block=0xff00) -> bytes:
    """Minimal bgzip writer: independent gzip members with the 'BC' size field + the empty EOF member."""
    out = []
    for at in list(range(0, len(data), block)) + [None]:
        chunk = b"" if at is None else data[at:at + block]
        comp = zlib.compressobj(6, zlib.DEFLATED, -15)
        payload = comp.compress(chunk) + comp.flush()
        bsize = 12 + 6 + len(payload) + 8
        out.append(b"\x1f\x8b\x08\x04" + b"\x00" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
                   + payload + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
    return b"".join(out)


def test_bgzf_is_inflated_in_parallel_and_equals_gzip(tmp_path, monkeypatch):
    rng = np.random.default_rng(5)
    text = to_fasta([(b"chr%d" % i, random_bases(rng, 700_000)) for i in range(3)], width=60)
    raw = bgzf_compress(text)
    assert gzip.decompress(raw) == text                    # a valid multi-member gzip stream
    members = ingest._bgzf_members(raw)
    assert members is not None and len(members) == -(-len(text) // 0xff00) + 1
    monkeypatch.setattr(ingest, "_BGZF_BATCH", 4)          # several tasks even for this small file
    assert ingest.gunzip(raw) == text
    # through the public entry points, with the digest of the FILE bytes (what the reference hashes)
    path = tmp_path / "g.fa.gz"
    path.write_bytes(raw)
    ingest.prefetch([str(path)])
    assert ingest.digest(str(path)) == hashlib.blake2b(raw).hexdigest()
    assert ingest.fasta_bytes(str(path)) == text


def test_plain_gzip_and_near_bgzf_fall_back(tmp_path):
    text = b">r\n" + b"ACGT" * 5000 + b"\n"
    plain = gzip.compress(text)
    assert ingest._bgzf_members(plain) is None and ingest.gunzip(plain) == text
    good = bgzf_compress(text, block=1000)
    broken = good[:-5]                                     # truncated last member: not BGZF end to end
    assert ingest._bgzf_members(broken) is None
    mixed = good + plain                                   # BGZF followed by a plain member
    assert ingest._bgzf_members(mixed) is None and ingest.gunzip(mixed) == text + text
    bad = bytearray(bgzf_compress(text * 40, block=1000))
    assert ingest._bgzf_members(bytes(bad)) is not None
    bad[len(bad) // 2] ^= 0xFF                             # flip a byte somewhere in the middle
    ingest_batch = ingest._BGZF_BATCH
    try:
        ingest._BGZF_BATCH = 4
        with pytest.raises(Exception):
            ingest.gunzip(bytes(bad))                      # bad header, bad deflate data or bad CRC: never silent
    finally:
        ingest._BGZF_BATCH = ingest_batch


def test_prefetch_keeps_every_worker_reading(tmp_path, monkeypatch):
    """N files on an N-worker pool are all read at the same time: no pool task waits for a sibling's
    future (the earlier read/text/hash triples left only a third of the workers doing anything), the
    text of a file is available before its digest, errors reach both futures, and what comes out is
    the file's bytes and hashlib's digest."""
    import threading
    from concurrent.futures import ThreadPoolExecutor
    n = 4
    texts = {}
    for i in range(n):
        texts[str(tmp_path / f"g{i}.fa")] = b">g%d\n" % i + b"ACGT" * (1000 + i) + b"\n"
        (tmp_path / f"g{i}.fa").write_bytes(texts[str(tmp_path / f"g{i}.fa")])
    gz_path = str(tmp_path / "z.fa.gz")
    with open(gz_path, "wb") as fh:
        fh.write(gzip.compress(b">z\nACGTTGCA\n"))
    monkeypatch.setattr(ingest, "_pool", ThreadPoolExecutor(max_workers=n))
    monkeypatch.setattr(ingest, "_jobs", {})
    together = threading.Barrier(n, timeout=20)
    hash_gate = threading.Event()
    real_read, real_hash = ingest._read, ingest._hash

    def read_all_at_once(path):
        if path in texts:
            together.wait()          # breaks (and fails the futures) unless all n reads run concurrently
        return real_read(path)

    def gated_hash(raw):
        hash_gate.wait(20)
        return real_hash(raw)

    monkeypatch.setattr(ingest, "_read", read_all_at_once)
    monkeypatch.setattr(ingest, "_hash", gated_hash)
    ingest.prefetch(list(texts))
    for path, want in texts.items():
        assert ingest.is_prefetched(path)
        assert ingest.fasta_bytes(path) == want          # the text does not wait for the digest
    hash_gate.set()
    for path, want in texts.items():
        assert ingest.digest(path) == hashlib.blake2b(want).hexdigest()
    ingest.prefetch([gz_path])
    assert ingest.fasta_bytes(gz_path) == b">z\nACGTTGCA\n"
    assert ingest.digest(gz_path) == hashlib.blake2b(open(gz_path, "rb").read()).hexdigest()
    # a file that disappears between prefetch() and the read: the error surfaces where the bytes are asked for
    gone = tmp_path / "gone.fa"
    gone.write_bytes(b">x\nAC\n")

    def failing_read(path):
        raise OSError("simulated read failure")
    monkeypatch.setattr(ingest, "_read", failing_read)
    ingest.prefetch([str(gone)])
    with pytest.raises(OSError):
        ingest.fasta_bytes(str(gone))


def test_large_plain_files_get_name_only_jobs(tmp_path, monkeypatch):
    """From HASH_ONLY_MIN_BYTES on, an uncompressed file is not held in memory by prefetch(): only its
    digest is prepared (chunked reads); the bytes are still there for whoever asks, and a gzip file of
    any size keeps the full job (its text only exists after inflation)."""
    monkeypatch.setattr(ingest, "HASH_ONLY_MIN_BYTES", 1000)
    monkeypatch.setattr(ingest, "_jobs", {})
    monkeypatch.setattr(ingest, "_cached", 0)
    rng = np.random.default_rng(9)
    big = to_fasta([(b"big", random_bases(rng, 3_000_000))], width=60)       # several 1 MiB hash blocks
    small = b">s\nACGT\n"
    (tmp_path / "big.fa").write_bytes(big)
    (tmp_path / "small.fa").write_bytes(small)
    (tmp_path / "big.fa.gz").write_bytes(gzip.compress(big, 1))
    paths = [str(tmp_path / n) for n in ("big.fa", "small.fa", "big.fa.gz")]
    ingest.prefetch(paths)
    assert not ingest.is_prefetched(paths[0]) and ingest.is_prefetched(paths[1]) and ingest.is_prefetched(paths[2])
    assert all(ingest.has_digest(p) for p in paths) and not ingest.has_digest(str(tmp_path / "big.fa") + "x")
    assert ingest._cached == len(small) + (tmp_path / "big.fa.gz").stat().st_size      # the name-only job holds no bytes
    assert ingest.digest(paths[0]) == hashlib.blake2b(big).hexdigest()
    assert ingest.fasta_bytes(paths[0]) == big and ingest.fasta_bytes(paths[2]) == big
    assert ingest.digest(paths[2]) == hashlib.blake2b((tmp_path / "big.fa.gz").read_bytes()).hexdigest()
    ingest.drop_all()
    assert ingest.digest(paths[1]) == hashlib.blake2b(small).hexdigest()
    # bytes asked for first, name later: the job's digest future survives the hand-over
    (tmp_path / "big2.fa").write_bytes(big + b">t\nAC\n")
    p2 = str(tmp_path / "big2.fa")
    ingest.prefetch([p2])
    assert ingest.fasta_bytes(p2) == big + b">t\nAC\n"
    assert ingest.has_digest(p2) and ingest.digest(p2) == hashlib.blake2b(big + b">t\nAC\n").hexdigest()


def test_hash_file_pieces_and_empty_file(tmp_path):
    data = np.random.default_rng(1).integers(0, 256, 1_234_567, dtype=np.uint8).tobytes()
    path = tmp_path / "x.bin"
    path.write_bytes(data)
    want = hashlib.blake2b(data).hexdigest()
    assert ingest._hash_file(str(path)) == want
    assert ingest._hash_file(str(path), piece=4096) == want          # many mapped pieces
    path.write_bytes(b"")
    assert ingest._hash_file(str(path)) == hashlib.blake2b(b"").hexdigest()   # cannot be mapped: read path


def test_bytes_only_prefetch_never_hashes(tmp_path, monkeypatch):
    """prefetch(want_digest=False) (helpers/allpairs.py: files that are never named): bytes are read in the
    background, no hashing task is queued, large plain files get no job at all, and a digest asked for
    later is still right."""
    monkeypatch.setattr(ingest, "HASH_ONLY_MIN_BYTES", 1000)
    monkeypatch.setattr(ingest, "_jobs", {})
    monkeypatch.setattr(ingest, "_cached", 0)
    hashed = []
    real = ingest._hash
    monkeypatch.setattr(ingest, "_hash", lambda raw: hashed.append(len(raw)) or real(raw))
    small, big = b">s\nACGTTGCA\n", b">b\n" + b"ACGT" * 1000 + b"\n"
    (tmp_path / "s.fa").write_bytes(small)
    (tmp_path / "b.fa").write_bytes(big)
    (tmp_path / "b.fa.gz").write_bytes(gzip.compress(big))
    paths = [str(tmp_path / n) for n in ("s.fa", "b.fa", "b.fa.gz")]
    ingest.prefetch(paths, want_digest=False)
    assert ingest.is_prefetched(paths[0]) and not ingest.is_prefetched(paths[1]) and ingest.is_prefetched(paths[2])
    assert paths[1] not in ingest._jobs and not any(ingest.has_digest(p) for p in paths)
    assert ingest.fasta_bytes(paths[0]) == small and ingest.fasta_bytes(paths[1]) == big and ingest.fasta_bytes(paths[2]) == big
    assert hashed == [] and ingest._cached == 0
    assert ingest.digest(paths[0]) == hashlib.blake2b(small).hexdigest()
    ingest.prefetch([paths[0]], want_digest=False)
    ingest.drop(paths[0])                                  # a dropped bytes-only job leaves no digest future behind
    assert ingest._cached == 0 and ingest.digest(paths[2]) == hashlib.blake2b((tmp_path / "b.fa.gz").read_bytes()).hexdigest()
