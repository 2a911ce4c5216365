"""Stream sharding across GPUs: the inter-frame dependency makes one stream sequential, so the box is
partitioned by independent streams -- rank r owns a contiguous block of streams and there is NO collective on the
data path.  The only cross-rank traffic is the timing reduction (max over ranks) used by bench.py."""
from __future__ import annotations


def stream_range(total_streams: int, world: int, rank: int) -> range:
    """Contiguous block of streams owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world) or total_streams < 0:
        raise ValueError("bad rank/world")
    base, extra = divmod(total_streams, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def owner_of(stream: int, total_streams: int, world: int) -> int:
    base, extra = divmod(total_streams, world)
    edge = extra * (base + 1)
    return stream // (base + 1) if stream < edge else extra + (stream - edge) // max(base, 1)


def max_over_ranks(values, dist=None, device=None):
    """Element-wise max of a list of floats over all ranks (the contract's 'max over ranks' timing)."""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]
