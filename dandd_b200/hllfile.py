"""Dashing-v1 style `.hll` sketch files: a (possibly gzip-compressed) stream of
    uint32[N] flags (is_calculated, clamp, estimation method, joint estimation method[, nthreads])
    uint32    p
    float64   cached estimate
    2^p bytes registers
DandD itself never opens these files -- it only tests that they exist and are non-empty
(reference lib/sketch_classes.py:323-334).  The layout above is a RECOLLECTION of dnbaker/sketch
hll_t::write (no Dashing source or binary is available here), so interoperability
with real Dashing is a goal, not a tested fact.  What is uncertain is kept in one place:
    HEADER_FLAG_WORDS   how many uint32 flags precede p: 4 as the survey recalls, or 5 if the
                        writer also stores its thread count.  The WRITER uses this constant; the
                        READER accepts both widths (the file size, 28 or 32 + 2^p bytes, says which).
    JESTIM_*            the estimator enum values written into the flags (informational: this
                        package always re-estimates with the Ertl MLE from the registers).
tools/crosscheck_dashing.py diffs this module against a real `dashing` when one is on PATH."""
import gzip
import os
import struct
import zlib

import numpy as np

HEADER_FLAG_WORDS = int(os.environ.get("DANDD_B200_HLL_FLAG_WORDS", "4"))   # 4 (28-byte header) or 5 (32-byte)
if HEADER_FLAG_WORDS not in (4, 5):
    raise ValueError("DANDD_B200_HLL_FLAG_WORDS must be 4 or 5")
HEADERS = {4: struct.Struct("<4I I d"), 5: struct.Struct("<5I I d")}
HEADER = HEADERS[HEADER_FLAG_WORDS]
JESTIM_ERTL_MLE = 2          # value recalled for hll::EstimationMethod::ERTL_MLE (unconfirmed)
JESTIM_ERTL_JOINT_MLE = 3    # value recalled for hll::JointEstimationMethod::ERTL_JOINT_MLE (unconfirmed)
ERTL_MLE = ERTL_JOINT_MLE = JESTIM_ERTL_MLE   # older names, kept for callers


def _pack_header(known: int, p: int, card: float) -> bytes:
    """Two readings of the four-word header are in circulation (both recollections):
        (is_calculated, clamp, estimation method, joint estimation method)      the survey's
        (is_calculated, estimation method, joint estimation method, <unused>)   later dnbaker/sketch
    The words written here, (known, 2, 2, 3), name the Ertl MLE under BOTH: clamp = 2 is just "true" (it
    only matters to the original estimator), a joint method of 2 (ERTL_MLE) or 3 (ERTL_JOINT_MLE) is a
    valid choice either way -- so a real Dashing that re-estimates one of these files (after a `union`,
    say) does not silently fall back to estimator 0, the original HyperLogLog formula.  The five-word
    form is unambiguous: (is_calculated, clamp, method, joint method, threads)."""
    if HEADER_FLAG_WORDS == 5:
        flags = [known, 0, JESTIM_ERTL_MLE, JESTIM_ERTL_JOINT_MLE, 1]
    else:
        flags = [known, JESTIM_ERTL_MLE, JESTIM_ERTL_MLE, JESTIM_ERTL_JOINT_MLE]
    return HEADER.pack(*flags, p, card)


def _unpack_header(raw: bytes):
    """-> (header size, known, p, value) or None.  Both header widths are tried; the one whose p is
    consistent with the file size (header + 2^p register bytes) wins."""
    for words in (HEADER_FLAG_WORDS, 9 - HEADER_FLAG_WORDS):
        h = HEADERS[words]
        if len(raw) < h.size:
            continue
        fields = h.unpack_from(raw)
        known, p, value = fields[0], fields[words], fields[words + 1]
        if 4 <= p <= 32 and len(raw) == h.size + (1 << p):
            return h.size, known, p, value
    return None


def write_hll(path: str, regs: np.ndarray, p: int, card: float = 0.0, compresslevel: int = 0) -> None:
    """compresslevel 0 writes the raw stream (fast; still readable through gzread), 1-9 gzip."""
    regs = np.ascontiguousarray(regs, dtype=np.uint8)
    if regs.size != 1 << p:
        raise ValueError(f"expected {1 << p} registers, got {regs.size}")
    known = 1 if card and np.isfinite(card) else 0
    head = _pack_header(known, p, float(card) if known else 0.0)
    tmp = f"{path}.tmp{os.getpid()}"
    if compresslevel > 0:
        with gzip.open(tmp, "wb", compresslevel=compresslevel) as f:
            f.write(head)
            f.write(regs.tobytes())
    else:
        with open(tmp, "wb") as f:
            f.write(head)
            f.write(regs.tobytes())
    os.replace(tmp, path)   # a reader never sees a half-written sketch


STUB_MAGIC = b"\nDANDD-B200-UNION-OF\n"


class StubSketch(Exception):
    """The file is a union marker (header + member list, no registers)."""

    def __init__(self, path, p, card, members):
        super().__init__(f"{path} is a union stub of {len(members)} sketches")
        self.path, self.p, self.card, self.members = path, p, card, members


def write_stub(path: str, p: int, card: float, members) -> None:
    """DANDD_B200_UNION_FILES=stub: a non-empty marker instead of a full union sketch -- the header
    (so `p` and the cached cardinality survive) followed by the member sketch paths, from which the
    registers can be rebuilt on demand."""
    # one write of a few hundred bytes: no temporary + rename (a `kij` run leaves 10^5 of these behind and
    # the two extra system calls per file were a third of its wall time); a torn marker fails to parse and
    # is rebuilt like any unreadable sketch (SketchObj.individual_card)
    with open(path, "wb") as f:
        f.write(_pack_header(1, p, float(card)) + STUB_MAGIC + "\n".join(members).encode())


def read_hll(path: str):
    """-> (registers uint8[2^p], p, cached estimate or None); raises StubSketch for a union marker.
    Accepts 28- and 32-byte headers, gzip-compressed or raw."""
    with open(path, "rb") as f:
        raw = f.read()
    for h in HEADERS.values():
        if raw[h.size:h.size + len(STUB_MAGIC)] == STUB_MAGIC:
            fields = h.unpack_from(raw)
            words = len(fields) - 2
            raise StubSketch(path, fields[words], fields[words + 1], raw[h.size + len(STUB_MAGIC):].decode().split("\n"))
    if raw[:2] == b"\x1f\x8b":
        raw = zlib.decompress(raw, 16 + zlib.MAX_WBITS)
    got = _unpack_header(raw)
    if got is None:
        raise ValueError(f"{path}: not a 2^p-register sketch ({len(raw)} bytes: neither a 28- nor a 32-byte header fits)")
    size, known, p, value = got
    regs = np.frombuffer(raw, dtype=np.uint8, offset=size).copy()
    return regs, p, (value if known else None)
