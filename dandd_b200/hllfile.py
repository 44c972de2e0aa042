"""Dashing-v1 style `.hll` sketch files: a (possibly gzip-compressed) stream of
    uint32[4] flags (is_calculated, clamp, estimation method, joint estimation method)
    uint32    p
    float64   cached estimate
    2^p bytes registers
DandD itself never opens these files -- it only tests that they exist and are non-empty
(reference lib/sketch_classes.py:323-334) -- but writing the real layout lets a sketchdb made here
be read by Dashing's own `card`/`union` (its reader goes through zlib, which passes uncompressed
files through unchanged) and lets sketches cached by real Dashing be re-used here."""
import gzip
import os
import struct
import zlib

import numpy as np

HEADER = struct.Struct("<4I I d")
ERTL_MLE, ERTL_JOINT_MLE = 2, 2


def write_hll(path: str, regs: np.ndarray, p: int, card: float = 0.0, compresslevel: int = 0) -> None:
    """compresslevel 0 writes the raw stream (fast; still readable through gzread), 1-9 gzip."""
    regs = np.ascontiguousarray(regs, dtype=np.uint8)
    if regs.size != 1 << p:
        raise ValueError(f"expected {1 << p} registers, got {regs.size}")
    known = 1 if card and np.isfinite(card) else 0
    head = HEADER.pack(known, 0, ERTL_MLE, ERTL_JOINT_MLE, p, float(card) if known else 0.0)
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
    tmp = f"{path}.tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(HEADER.pack(1, 0, ERTL_MLE, ERTL_JOINT_MLE, p, float(card)))
        f.write(STUB_MAGIC + "\n".join(members).encode())
    os.replace(tmp, path)


def read_hll(path: str):
    """-> (registers uint8[2^p], p, cached estimate or None); raises StubSketch for a union marker."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[HEADER.size:HEADER.size + len(STUB_MAGIC)] == STUB_MAGIC:
        known, _c, _e, _j, p, value = HEADER.unpack_from(raw)
        raise StubSketch(path, p, value, raw[HEADER.size + len(STUB_MAGIC):].decode().split("\n"))
    if raw[:2] == b"\x1f\x8b":
        raw = zlib.decompress(raw, 16 + zlib.MAX_WBITS)
    if len(raw) < HEADER.size:
        raise ValueError(f"{path}: too short for a sketch header")
    known, _clamp, _est, _jest, p, value = HEADER.unpack_from(raw)
    if not 4 <= p <= 32 or len(raw) != HEADER.size + (1 << p):
        raise ValueError(f"{path}: not a 2^p-register sketch (p={p}, {len(raw)} bytes)")
    regs = np.frombuffer(raw, dtype=np.uint8, offset=HEADER.size).copy()
    return regs, p, (value if known else None)
