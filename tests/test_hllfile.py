"""Sketch files (dandd_b200/hllfile.py): Dashing-v1 layout round trips, union stubs, bad input."""
import gzip
import struct

import numpy as np
import pytest

from dandd_b200 import hllfile


@pytest.mark.parametrize("p", [5, 10, 14])
@pytest.mark.parametrize("level", [0, 1, 6])
def test_roundtrip(tmp_path, p, level):
    rng = np.random.default_rng(p * 10 + level)
    regs = rng.integers(0, 64 - p + 2, 1 << p).astype(np.uint8)
    path = str(tmp_path / "a.hll")
    hllfile.write_hll(path, regs, p, 12345.678, compresslevel=level)
    got, gp, card = hllfile.read_hll(path)
    assert np.array_equal(got, regs) and gp == p and card == 12345.678
    raw = open(path, "rb").read()
    body = gzip.decompress(raw) if level else raw
    # the documented layout: uint32[4] flags, uint32 p, float64 estimate, 2^p register bytes
    flags = struct.unpack_from("<4I", body)
    assert flags[0] == 1 and struct.unpack_from("<I", body, 16)[0] == p and len(body) == 28 + (1 << p)
    assert body[28:] == regs.tobytes()
    hllfile.write_hll(path, regs, p)                       # no cached estimate
    assert hllfile.read_hll(path)[2] is None


def test_stub_and_bad_files(tmp_path):
    path = str(tmp_path / "u.hll")
    hllfile.write_stub(path, 12, 777.0, ["/x/a.hll", "/x/b.hll"])
    with pytest.raises(hllfile.StubSketch) as e:
        hllfile.read_hll(path)
    assert e.value.p == 12 and e.value.card == 777.0 and e.value.members == ["/x/a.hll", "/x/b.hll"]
    with pytest.raises(ValueError):
        hllfile.write_hll(path, np.zeros(100, dtype=np.uint8), 8)      # wrong register count
    short = tmp_path / "s.hll"
    short.write_bytes(b"abc")
    with pytest.raises(ValueError):
        hllfile.read_hll(str(short))
    trunc = tmp_path / "t.hll"
    hllfile.write_hll(str(trunc), np.zeros(1 << 8, dtype=np.uint8), 8, 1.0)
    trunc.write_bytes(trunc.read_bytes()[:-5])
    with pytest.raises(ValueError):
        hllfile.read_hll(str(trunc))


def test_reader_accepts_both_header_widths(tmp_path):
    """The recalled dnbaker/sketch layout has four or five uint32 flags before p (28- or 32-byte
    header); the writer uses HEADER_FLAG_WORDS, the reader takes whichever fits the file size."""
    p = 10
    regs = np.arange(1 << p, dtype=np.uint32).astype(np.uint8) % 50
    for words in (4, 5):
        raw = struct.pack(f"<{words}I I d", 1, 0, 2, 2, *([8] if words == 5 else []), p, 4242.5) + regs.tobytes()
        for packer in (lambda b: b, gzip.compress):
            path = tmp_path / f"w{words}.hll"
            path.write_bytes(packer(raw))
            got, gp, card = hllfile.read_hll(str(path))
            assert np.array_equal(got, regs) and gp == p and card == 4242.5
    assert hllfile.HEADER.size == 12 + 4 * hllfile.HEADER_FLAG_WORDS


def test_flag_words_name_the_ertl_mle_under_both_recalled_layouts(tmp_path):
    """(is_calculated, clamp, method, joint) or (is_calculated, method, joint, unused): either way the
    method word is 2 (ERTL_MLE) and the joint word is a valid joint method (2 or 3), never 0 (original)."""
    path = str(tmp_path / "f.hll")
    hllfile.write_hll(path, np.zeros(1 << 6, dtype=np.uint8), 6, 5.0)
    w = struct.unpack_from("<4I", open(path, "rb").read())
    assert w[0] == 1
    assert w[2] == hllfile.JESTIM_ERTL_MLE and w[3] == hllfile.JESTIM_ERTL_JOINT_MLE       # survey's order
    assert w[1] == hllfile.JESTIM_ERTL_MLE and w[2] in (hllfile.JESTIM_ERTL_MLE, hllfile.JESTIM_ERTL_JOINT_MLE)   # later order
