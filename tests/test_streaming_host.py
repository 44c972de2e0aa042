"""Host side of dandd_b200/streaming.py that needs no GPU: the positional readers fill ring slots out of
order, the feeder sees the chunks in file order, and no failure strands a slot or a thread."""
import os
import queue
import threading

import numpy as np

from dandd_b200 import streaming


def _run_readers(path, chunk, readers, consume):
    ring = streaming._Ring(chunk, pin=False)
    size = os.path.getsize(path)
    cursor = streaming._Cursor(size, chunk)
    out_q, stop = queue.Queue(), threading.Event()
    fd = os.open(path, os.O_RDONLY)
    threads = [threading.Thread(target=streaming._reader_at, args=(fd, ring, out_q, cursor, stop), daemon=True)
               for _ in range(readers)]
    for t in threads:
        t.start()
    try:
        for kind, slot, n in streaming._in_order(out_q, readers):
            consume(kind, slot, n, ring, stop)
    finally:
        for t in threads:
            t.join(20)
        os.close(fd)
    assert not any(t.is_alive() for t in threads)
    assert ring.free.qsize() == streaming.RING_SLOTS       # every slot is back
    return ring


def test_positional_readers_deliver_the_file_in_order(tmp_path):
    rng = np.random.default_rng(3)
    for size in (1, (1 << 20) - 1, 1 << 20, (1 << 20) + 1, 7 * (1 << 20) + 12345, 20 << 20):
        data = rng.integers(0, 256, size, dtype=np.uint8).tobytes()
        path = tmp_path / f"f{size}"
        path.write_bytes(data)
        for readers in (1, 3, 5):
            parts = []

            def consume(kind, slot, n, ring, stop):
                assert kind == "chunk"
                parts.append(bytes(ring.views[slot][:n]))
                ring.release(slot)
            _run_readers(str(path), 1 << 20, readers, consume)
            assert b"".join(parts) == data, (size, readers)
            assert len(parts) == -(-size // (1 << 20))


def test_feeder_failure_stops_the_readers(tmp_path):
    path = tmp_path / "big"
    path.write_bytes(os.urandom(40 << 20))
    seen = []

    def consume(kind, slot, n, ring, stop):
        seen.append(kind)
        if len(seen) == 3:
            stop.set()                 # what sketch_file does when a device call fails
        ring.release(slot)
    _run_readers(str(path), 1 << 20, 3, consume)
    assert len(seen) < 40              # the rest of the file was not read


def test_reader_error_reaches_the_feeder_and_orphans_are_returned(tmp_path, monkeypatch):
    path = tmp_path / "big"
    path.write_bytes(os.urandom(12 << 20))
    real = os.preadv
    lock = threading.Lock()
    calls = {"n": 0}

    def flaky(fd, bufs, off):
        with lock:
            calls["n"] += 1
            fail = off == 2 << 20          # the third chunk never arrives
        if fail:
            raise OSError("simulated read error")
        return real(fd, bufs, off)
    monkeypatch.setattr(os, "preadv", flaky)
    kinds, parts = [], []
    data = path.read_bytes()

    def consume(kind, slot, n, ring, stop):
        kinds.append(kind)
        if kind == "error":
            assert isinstance(slot, OSError)
            return
        if kind == "chunk":
            parts.append(bytes(ring.views[slot][:n]))
        ring.release(slot)
    _run_readers(str(path), 1 << 20, 3, consume)
    assert kinds.count("error") == 1 and kinds.count("chunk") <= 2      # nothing behind the gap is fed
    assert b"".join(parts) == data[:len(parts) << 20]
    assert all(k in ("chunk", "error", "orphan") for k in kinds)


def test_in_order_reorders_and_flushes_orphans():
    q = queue.Queue()
    for item in ((2, 12, 5), (0, 10, 5), (1, 11, 5), None, (4, 14, 3), RuntimeError("x"), None):
        q.put(item)
    got = list(streaming._in_order(q, 2))
    assert [g[0] for g in got] == ["chunk", "chunk", "chunk", "error", "orphan"]
    assert [g[1] for g in got if g[0] != "error"] == [10, 11, 12, 14]


# ---- the feeder of sketch_file on a CPU double --------------------------------------------------------
class _FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 0.0


class _FakeStream:
    cuda_stream = 0

    def __init__(self, device=None):
        pass

    def wait_event(self, ev):
        pass

    def wait_stream(self, s):
        pass


class _FakeCuda:
    Event = _FakeEvent
    Stream = _FakeStream

    @staticmethod
    def current_stream(dev=None):
        return _FakeStream()

    class stream:                      # `with torch.cuda.stream(s):`
        def __init__(self, s):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False


class _FakeLib:
    """Stands in for the C ABI: records the text every dd_pack_fasta call is given (a synchronous CPU
    'device'), answers the host-only helper with the real library."""

    def __init__(self, real):
        self.real = real
        self.fed = []
        self.updates = []
        self.fail_at = None

    def dd_pack_codes_bytes(self, n):
        return 64

    dd_pack_invalid_bytes = dd_pack_codes_bytes

    def dd_pack_workspace_bytes(self, n):
        return 64

    def dd_sketch_workspace_bytes(self, nk, p):
        return 64

    def dd_pack_reset(self, codes, cb, invalid, ib, state, st):
        import ctypes
        ctypes.memset(state, 0, 32)
        return 0

    def dd_sketch_begin(self, *a):
        return 0

    def dd_fasta_first_record_host(self, ptr, n):
        return self.real.dd_fasta_first_record_host(ptr, n)

    def dd_pack_fasta(self, d_text, ln, *rest):
        import ctypes
        if self.fail_at is not None and len(self.fed) == self.fail_at:
            return -1
        self.fed.append(ctypes.string_at(d_text, ln))
        return 0

    def dd_sketch_update_sched(self, codes, invalid, state, a, b, ln, seen, *rest):
        self.updates.append((ln, seen))
        return 0

    def dd_sketch_end(self, *a):
        return 0

    def dd_last_error(self):
        return b"injected"


class _FakeEngine:
    polyt_sentinel = False

    def __init__(self, lib):
        import torch
        self.lib = lib
        self.device = torch.device("cpu")
        self._ws = {}

    def _buf(self, nbytes, tag):
        import torch
        t = self._ws.get(tag)
        if t is None or t.numel() < nbytes:
            t = self._ws[tag] = torch.empty(max(int(nbytes), 256), dtype=torch.uint8)
        return t


def _fake_setup(monkeypatch):
    import types
    import torch
    from dandd_b200 import _lib, build
    build.build()
    fake_torch = types.SimpleNamespace(cuda=_FakeCuda, empty=torch.empty, Tensor=torch.Tensor, uint8=torch.uint8,
                                       int32=torch.int32, float64=torch.float64)
    monkeypatch.setattr(streaming, "torch", fake_torch)
    monkeypatch.setattr(streaming, "_Ring", lambda chunk, pin=True, _R=streaming._Ring: _R(chunk, pin=False))
    monkeypatch.setattr(streaming, "_rings", {})
    monkeypatch.setattr(_lib, "load", lambda real=_lib.load(): real)
    lib = _FakeLib(_lib.load())
    return lib, _FakeEngine(lib)


def test_sketch_file_feeds_every_byte_once_and_in_order(tmp_path, monkeypatch):
    """sketch_file on a CPU double of the device: whatever the source (file read by one thread with the
    digest, file read by several threads without it, bytes already in memory), the packer is given
    the text from the first record marker to the end, in order, and the sketch schedule sees the
    running count of bytes."""
    import hashlib
    lib, eng = _fake_setup(monkeypatch)
    rng = np.random.default_rng(11)
    body = b">r1\n" + rng.choice(np.frombuffer(b"ACGT\n", dtype=np.uint8), 9_500_000).tobytes()
    for preamble in (b"", b"junk before the first record\n" * 50_000):      # the second spans more than a chunk
        text = preamble + body
        path = tmp_path / "t.fa"
        path.write_bytes(text)
        for kw in (dict(want_digest=True), dict(want_digest=False), dict(text=text)):
            lib.fed.clear()
            lib.updates.clear()
            regs, cards, digest, stats = streaming.sketch_file(eng, str(path), [21, 31], p=10, chunk_bytes=1 << 20, **kw)
            assert b"".join(lib.fed) == body, (len(preamble), kw.keys())
            assert [u[1] for u in lib.updates] == list(np.cumsum([0] + [u[0] for u in lib.updates[:-1]]))
            assert sum(u[0] for u in lib.updates) == len(body)
            assert digest == (hashlib.blake2b(text).hexdigest() if kw.get("want_digest") else None)
            assert stats.get("readers", 1) == (3 if kw == dict(want_digest=False) else 1)
            ring = streaming._ring_for(eng.device, 1 << 20)
            assert ring.free.qsize() == streaming.RING_SLOTS


def test_sketch_file_device_error_leaves_the_ring_whole(tmp_path, monkeypatch):
    from dandd_b200._lib import DandDError
    import pytest
    lib, eng = _fake_setup(monkeypatch)
    text = b">r\n" + b"ACGT" * 3_000_000
    path = tmp_path / "t.fa"
    path.write_bytes(text)
    before = threading.active_count()
    for kw in (dict(want_digest=True), dict(want_digest=False)):
        lib.fed.clear()
        lib.fail_at = 4
        with pytest.raises(DandDError):
            streaming.sketch_file(eng, str(path), [21], p=10, chunk_bytes=1 << 20, **kw)
        assert threading.active_count() == before
        ring = streaming._ring_for(eng.device, 1 << 20)
        assert ring.free.qsize() == streaming.RING_SLOTS
        lib.fail_at = None
        lib.fed.clear()
        streaming.sketch_file(eng, str(path), [21], p=10, chunk_bytes=1 << 20, **kw)
        assert b"".join(lib.fed) == text
