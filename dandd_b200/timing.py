"""Stage timings of a command-line run (machine-readable perf record; the reference only has
`/usr/bin/time -v` around whole stages, helpers/benchmark.sh:37-51).

    with timing.span("union"): ...        accumulate wall seconds under a name
    timing.add("blake2b", seconds)        same, for times measured elsewhere (worker threads)
    timing.dump()                         append one JSON line {rank, argv, wall_s, stages} to the
                                          file named by DANDD_B200_TIMING (no-op when unset)

Pure Python, no torch: importable before the interpreter has paid for CUDA start-up."""
import contextlib
import json
import os
import sys
import threading
import time

T0 = time.perf_counter()
_stages = {}
_lock = threading.Lock()


def add(name: str, seconds: float) -> None:
    with _lock:
        _stages[name] = _stages.get(name, 0.0) + float(seconds)


def mark(name: str) -> None:
    """Seconds since process start at which `name` happened (first call wins)."""
    with _lock:
        _stages.setdefault("at_" + name, time.perf_counter() - T0)


@contextlib.contextmanager
def span(name: str):
    t0 = time.perf_counter()
    try:
        yield
    finally:
        add(name, time.perf_counter() - t0)


def snapshot() -> dict:
    with _lock:
        return dict(_stages)


def dump(extra: dict = None) -> None:
    path = os.environ.get("DANDD_B200_TIMING")
    if not path:
        return
    rec = {"rank": int(os.environ.get("RANK", "0")), "world": int(os.environ.get("WORLD_SIZE", "1")), "argv": sys.argv[1:],
           "wall_s": time.perf_counter() - T0, "stages": snapshot()}
    if extra:
        rec.update(extra)
    with open(path, "a") as fh:
        fh.write(json.dumps(rec) + "\n")
