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
