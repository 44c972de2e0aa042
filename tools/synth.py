"""Synthetic human-scale FASTA, generated on the GPU: chromosome-like
records, ~1 % of the bases in N runs of 1000, 50 % soft-masked (lower case), 80-column lines;
mutated copies with a given substitution rate.  Deterministic in (size, records, seed)."""
import torch


def synth_fasta(n_bases: int, n_records: int, seed: int, device) -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lens = torch.full((n_records,), n_bases // n_records, dtype=torch.int64)
    lens[-1] += n_bases - int(lens.sum())
    pieces = []
    for r, ln in enumerate(lens.tolist()):
        ln80 = (ln // 80) * 80
        c = torch.randint(0, 4, (ln80,), dtype=torch.uint8, device=device, generator=g)
        b = 65 + 2 * (c == 1).to(torch.uint8) + 6 * (c == 2).to(torch.uint8) + 19 * (c == 3).to(torch.uint8)
        del c
        low = torch.randint(0, 2, (ln80,), dtype=torch.uint8, device=device, generator=g)
        b |= low * 32
        del low
        nruns = max(1, ln80 // 100000)               # ~1 % of the bases in N runs of 1000
        starts = torch.randint(0, max(1, ln80 - 1000), (nruns,), device=device, generator=g)
        idx = (starts[:, None] + torch.arange(1000, device=device)[None, :]).reshape(-1)
        b[idx] = 78
        body = torch.cat([b.view(-1, 80), torch.full((ln80 // 80, 1), 10, dtype=torch.uint8, device=device)], dim=1).reshape(-1)
        head = torch.tensor(list(f">chr{r + 1} synthetic length={ln80}\n".encode()), dtype=torch.uint8, device=device)
        pieces += [head, body]
    return torch.cat(pieces)


def mutate_text(text: torch.Tensor, rate: float, seed: int) -> torch.Tensor:
    """Substitute `rate` of the sequence letters (never touches headers, newlines or N)."""
    g = torch.Generator(device=text.device)
    g.manual_seed(seed)
    out = text.clone()
    n = out.numel()
    nsub = int(n * rate)
    # one position per stride-sized window: distinct by construction (duplicates would race in the
    # scatter below) and no sort kernel needed
    stride = max(1, n // max(1, nsub))
    nsub = n // stride
    pos = torch.arange(nsub, device=text.device) * stride + torch.randint(0, stride, (nsub,), device=text.device, generator=g)
    cur = out[pos]
    up = cur & 0xDF
    is_base = (up == 65) | (up == 67) | (up == 71) | (up == 84)
    new = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=text.device)[
        torch.randint(0, 4, (nsub,), device=text.device, generator=g)]
    out[pos[is_base]] = new[is_base] | (cur[is_base] & 0x20)
    return out


