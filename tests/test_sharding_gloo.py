"""World-size-2 gloo test of the replica-sharding bookkeeping used by bench.py at N > 1 (CPU only)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xpoint_b200.sharding import aggregate_throughput, max_over_ranks, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 129):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(129, rank, world)
    # rank r "processes" its pairs in (r + 1) * 100 ms
    tput = aggregate_throughput(hi - lo, 100.0 * (rank + 1))
    worst = max_over_ranks(100.0 * (rank + 1))
    # identical weights on every rank: same seed -> same tensor
    torch.manual_seed(0)
    w = torch.randn(16)
    ws = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(ws, w)
    q.put((rank, lo, hi, tput, worst, all(torch.equal(w, o) for o in ws)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replicas_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, w0, same0), (r1, lo1, hi1, t1, w1, same1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 65, 65, 129)
    assert w0 == w1 == 200.0                       # max over ranks
    assert abs(t0 - 129 / 0.2) < 1e-9 and t0 == t1  # all pairs / slowest rank
    assert same0 and same1


# ------------------------------------------------------------------------------------------ ShardedPairPipeline (f4)
class _FakeStream:
    """Stands in for PairStream on CPU: 'keypoints' are a deterministic function of the images, so the gather can be checked."""

    def __init__(self, device):
        self.device = device

    def run(self, batches, keep=True):
        out = []
        for o, t in batches:
            s = (o.flatten(1).sum(1) * 1000).round().to(torch.int32)
            out.append({"n_optical": s, "n_thermal": (t.flatten(1).sum(1) * 1000).round().to(torch.int32),
                        "kp_optical": s[:, None, None].expand(-1, 3, 2).contiguous()})
        return out


def _expected(o, t):
    s = (o.flatten(1).sum(1) * 1000).round().to(torch.int32)
    return s, (t.flatten(1).sum(1) * 1000).round().to(torch.int32)


def test_sharded_pipeline_in_process_devices_and_padding():
    """Two fake devices, 13 pairs in batches of 4: shards [0,7) and [7,13), the padded tail batches are trimmed, results
    come back in input order."""
    from xpoint_b200.pipeline import ShardedPairPipeline
    g = torch.Generator().manual_seed(0)
    o, t = torch.rand(13, 1, 8, 8, generator=g), torch.rand(13, 1, 8, 8, generator=g)
    sp = ShardedPairPipeline(None, devices=["fake0", "fake1"], batch=4, stream_factory=_FakeStream)
    res = sp.run(o, t)
    so, st = _expected(o, t)
    assert torch.equal(res["n_optical"], so) and torch.equal(res["n_thermal"], st)
    assert res["kp_optical"].shape == (13, 3, 2) and torch.equal(res["kp_optical"][:, 0, 0], so)


def _sharded_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xpoint_b200.pipeline import ShardedPairPipeline
    g = torch.Generator().manual_seed(1)                       # every rank holds the same host list of pairs
    o, t = torch.rand(11, 1, 8, 8, generator=g), torch.rand(11, 1, 8, 8, generator=g)
    sp = ShardedPairPipeline(None, devices=[f"fake{rank}"], batch=4, distributed=True, stream_factory=_FakeStream)
    res = sp.run(o, t)
    so, st = _expected(o, t)
    q.put((rank, bool(torch.equal(res["n_optical"], so) and torch.equal(res["n_thermal"], st)), tuple(res["kp_optical"].shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_pipeline_distributed_gather_gloo():
    """World size 2 on gloo: each rank runs its shard_range of the pairs, all_gather_object reassembles them in input order on
    every rank."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (11, 3, 2)), (1, True, (11, 3, 2))]
