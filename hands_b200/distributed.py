"""Multi-GPU plumbing for the geometry path: one process per GPU, batch sharded, no data-path collective.

The path is embarrassingly parallel over hands: every rank runs the same kernels on a
contiguous slice of the batch with its own copy of the (2.9 MB) MANO constants.  The only exchanges around
it are the ones of the training step the reference gets from Lightning DDP (`scripts_method/train.py:61,72`):
  * an all-reduce (mean) of the upstream parameter gradients after backward, and
  * one all-reduce of a packed fp32 vector of loss / metric partial sums and counts
    (the reference logs rank-0 values only, `common/abstract_pl.py:83,107,161`; with world_size 1 the packed
    reduction returns exactly those).
Both go through `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from typing import Dict, Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of a batch of n units for `rank` (first n % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world) or n < 0:
        raise ValueError(f"bad shard request n={n} rank={rank} world={world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, rank: int, world: int, units_per_row: int = 1) -> torch.Tensor:
    """Rows of `t` belonging to `rank` when the leading dimension holds `units_per_row` rows per unit
    (e.g. 2 crops per sample)."""
    n = t.shape[0] // units_per_row
    lo, hi = shard_bounds(n, rank, world)
    return t[lo * units_per_row : hi * units_per_row]


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def allreduce_gradients(params: Iterable[torch.Tensor], bucket_bytes: int = 64 << 20) -> int:
    """Average `.grad` of `params` across ranks in flat buckets (one collective per <= bucket_bytes).
    Parameters without a gradient contribute zeros, like DDP's find_unused_parameters mode
    (`train.py:72`).  Returns the number of collectives issued."""
    world = _world()
    params = [p for p in params if p.requires_grad]
    if world == 1 or not params:
        return 0
    calls = 0
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size, calls
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat)
        flat /= world
        off = 0
        for p in bucket:
            n = p.numel()
            g = flat[off : off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        calls += 1
        bucket, size = [], 0

    for p in params:
        nbytes = p.numel() * p.element_size()
        if bucket and (size + nbytes > bucket_bytes or p.dtype != bucket[0].dtype):
            flush()
        bucket.append(p)
        size += nbytes
    flush()
    return calls


class PackedMetrics:
    """Loss / metric scalars reduced with ONE collective: each entry is kept as (sum, count) in a flat fp32
    vector; `reduce()` all-reduces the vector and returns the global means."""

    def __init__(self, names: List[str], device):
        self.names = list(names)
        self.index = {n: i for i, n in enumerate(self.names)}
        self.buf = torch.zeros(2 * len(self.names), dtype=torch.float32, device=device)

    def add(self, name: str, value_sum, count) -> None:
        i = self.index[name]
        self.buf[2 * i] += value_sum
        self.buf[2 * i + 1] += count

    def reduce(self) -> Dict[str, float]:
        if _world() > 1:
            dist.all_reduce(self.buf)
        host = self.buf.detach().cpu().tolist()
        out = {}
        for n, i in self.index.items():
            s, c = host[2 * i], host[2 * i + 1]
            out[n] = s / c if c > 0 else float("nan")
        self.buf.zero_()
        return out
