"""Multi-GPU path = batch sharding with no data-path collective; only the benchmark's timing protocol
communicates.  Covered here with world_size-2 gloo processes on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fsr_b200 import sharding


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_single_process_helpers_are_identity():
    assert sharding.max_over_ranks(3.5, torch.device("cpu")) == 3.5
    assert sharding.sum_over_ranks(2.0, torch.device("cpu")) == 2.0
    sharding.barrier()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = sharding.shard_range(65536, rank, world)
        sharding.barrier()
        t = sharding.max_over_ranks(10.0 + rank, torch.device("cpu"))
        n = sharding.sum_over_ranks(float(e - b), torch.device("cpu"))
        # data-parallel gradient exchange of the training step: sum, then / world
        from fsr_b200 import training
        g = torch.full((5,), float(rank + 1)) * torch.arange(1, 6)
        training.allreduce_mean_(g)
        # the bucketed exchange: slices of the flat gradient arrive in the order the backward completes them
        # (descending addresses: tail, groups last to first, head) and are averaged bucket by bucket
        total = 1000
        flat = torch.arange(total, dtype=torch.float32) * (rank + 1)
        ex = training.BucketedAllReduce(total, n_buckets=4)
        sent = []
        orig_flush = ex._flush
        def spy(gr):
            if ex._lo is not None:
                sent.append((ex._lo, ex._hi))
            orig_flush(gr)
        ex._flush = spy
        for begin, count in [(900, 100), (750, 150), (600, 150), (450, 150), (300, 150), (150, 150), (10, 140), (0, 10)]:
            ex.on_stage(flat, begin, count)
        ex.finish(flat)
        q.put((rank, b, e, t, n, g.tolist(), flat.tolist() == (torch.arange(total, dtype=torch.float32) * 1.5).tolist(), sent))
    finally:
        dist.destroy_process_group()


def test_two_rank_timing_protocol_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 32768), (32768, 65536)]
    assert all(r[3] == 11.0 for r in res)      # max over ranks of the per-rank time
    assert all(r[4] == 65536.0 for r in res)   # every image processed exactly once
    assert all(r[5] == [1.5, 3.0, 4.5, 6.0, 7.5] for r in res)   # mean of the two ranks' gradients, on both
    assert all(r[6] for r in res)              # bucketed exchange: every element averaged exactly once
    # contiguous buckets, tail first; once what is still to come is small the pending bucket goes out at once, so that
    # only the small last piece (the head of the network, which finishes last) is exposed
    assert res[0][7] == res[1][7] == [(750, 1000), (450, 750), (150, 450), (10, 150), (0, 10)]


def test_allreduce_mean_is_identity_without_process_group():
    from fsr_b200 import training
    g = torch.arange(4, dtype=torch.float32)
    assert torch.equal(training.allreduce_mean_(g.clone()), g)
