"""Which rank owns which genome: pure Python (no torch), so the command line can start reading its
files before the interpreter has paid for importing torch and starting CUDA."""
from typing import List, Sequence


def shard_by_size(sizes: Sequence[int], nranks: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of genomes (by byte size) to ranks; deterministic,
    every rank computes the same table.  Returns the genome indices of each rank, ascending."""
    order = sorted(range(len(sizes)), key=lambda i: (-int(sizes[i]), i))
    load = [0] * nranks
    out = [[] for _ in range(nranks)]
    for i in order:
        r = min(range(nranks), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += int(sizes[i])
    return [sorted(x) for x in out]
