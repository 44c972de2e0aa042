"""Multi-GPU sharding of the hot path: one process per GPU (torchrun), genomes
sharded across ranks, no collective on the sketching itself.  The path has exactly one exchange
step -- merging register arrays -- and HLL merge is an exact, associative, commutative, idempotent
max, so the results are bit-identical to a single-GPU run:

    union_over_ranks   all_reduce(MAX, uint8)  over [nk][2^p]      (the union sketch of the job)
    gather_registers   all_gather              of [n_local][nk][2^p] (every rank gets every leaf:
                                               progressive orderings / pair tiles are then split
                                               across ranks with no further communication)
    sum_counts         all_reduce(SUM, int64)  for key-range-sharded exact counts

Backend: NCCL over NVLink on GPUs, gloo in the CPU tests (tests/test_dist.py, world_size 2).
The payloads are tiny next to NVLink bandwidth (<= 31 MiB per genome), so plain collectives on the
compute stream are the right tool; there is no compute kernel to fuse them into."""
import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def init(backend: str = None) -> (int, int):
    """Initialise the default process group from the torchrun environment (no-op if world is 1)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kwargs = {}
        if backend == "nccl":
            kwargs["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend, **kwargs)
    return rank, world


def world() -> (int, int):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


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


def union_over_ranks(regs: torch.Tensor) -> torch.Tensor:
    """In-place register-wise max over all ranks (uint8 tensor of any shape)."""
    _, n = world()
    if n > 1:
        dist.all_reduce(regs, op=dist.ReduceOp.MAX)
    return regs


def gather_registers(local: torch.Tensor, owners: List[List[int]]) -> torch.Tensor:
    """local: [n_local, nk, m] registers of this rank's genomes (in the order of owners[rank]).
    Returns [n_total, nk, m] in global genome order on every rank."""
    rank, n = world()
    total = sum(len(o) for o in owners)
    if n == 1:
        return local
    nk, m = local.shape[1], local.shape[2]
    width = max(len(o) for o in owners)
    padded = torch.zeros((width, nk, m), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    bucket = [torch.empty_like(padded) for _ in range(n)]
    dist.all_gather(bucket, padded)
    out = torch.empty((total, nk, m), dtype=local.dtype, device=local.device)
    for r, idxs in enumerate(owners):
        for j, g in enumerate(idxs):
            out[g] = bucket[r][j]
    return out


def gather_cards(local: torch.Tensor, owners: List[List[int]]) -> torch.Tensor:
    """Same for the [n_local, nk] float64 cardinalities."""
    return gather_registers(local.unsqueeze(-1), owners).squeeze(-1)


def sum_counts(counts: torch.Tensor) -> torch.Tensor:
    _, n = world()
    if n > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def split_work(n_items: int) -> range:
    """Contiguous slice of n_items (orderings, pair tiles) owned by this rank."""
    rank, n = world()
    lo = (n_items * rank) // n
    hi = (n_items * (rank + 1)) // n
    return range(lo, hi)
