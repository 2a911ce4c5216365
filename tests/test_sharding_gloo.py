"""The N>1 path on CPU: world_size-2 `gloo` processes partition the streams with no data-path collective; the only
cross-rank operation is the max-over-ranks timing reduction bench.py uses."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from p64_b200 import shard


def test_stream_ranges_partition_exactly():
    for total in (0, 1, 7, 256, 2048, 2049):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                rg = shard.stream_range(total, world, r)
                seen += list(rg)
                assert all(shard.owner_of(s, total, world) == r for s in rg)
            assert seen == list(range(total))
            sizes = [len(shard.stream_range(total, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.stream_range(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard.stream_range(11, world, rank)
        # each rank "encodes" its own streams independently (host-side stand-in: a checksum per stream)
        local = {s: (s * 2654435761) & 0xffff for s in mine}
        ms = shard.max_over_ranks([10.0 + rank, 5.0 - rank], dist)
        gathered = [None] * world
        dist.all_gather_object(gathered, local)       # test-only: prove the shards were disjoint and complete
        q.put((rank, ms, gathered))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, gathered in out:
        assert ms == [11.0, 5.0]
        merged = {}
        for d in gathered:
            assert not (set(d) & set(merged))
            merged.update(d)
        assert sorted(merged) == list(range(11))


def test_link_proportional_split_is_a_partition():
    """bench.py's link-balanced end-to-end partition (and encoder.cpp's balance_links): integer shares proportional to the
    measured link rates, summing to the job's stream count, none below the floor"""
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    sys.modules["bench_mod"] = bench
    spec.loader.exec_module(bench)
    sp = bench.split_proportional
    assert sp(2048, [23.5] * 4 + [36.0] * 4) == [202, 202, 202, 202, 310, 310, 310, 310]
    assert sp(512, [55.2, 55.2]) == [256, 256]
    for total, w in [(2048, [1, 2, 3, 4, 5, 6, 7, 8]), (1024, [29.0, 29.1, 28.9, 29.0]), (300, [0.0, 10.0, 10.0]), (64, [1e-9, 1.0])]:
        s = sp(total, w)
        assert sum(s) == total and min(s) >= 8 and len(s) == len(w)
        big = [i for i, x in enumerate(w) if x > 1e-6]
        for i in big:                                   # proportional to within rounding and the floor of the starved ranks
            assert abs(s[i] - total * w[i] / sum(w)) <= 8 * len(w)
