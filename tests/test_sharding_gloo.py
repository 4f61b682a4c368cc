"""The N > 1 path on CPU: world_size 2 over gloo (127.0.0.1).  Streams are sharded by stream id
with no data-path collective; each rank resamples its shard (the oracle stands in for the GPU
here), and the per-rank results / timings are combined exactly as bench.py combines them."""
import os
import socket

import numpy as np
import pytest

import oracle_lib as O
from resampler_b200.sharding import all_reduce_scalar, job_throughput, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1024, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            got = [shard_range(n, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            for a, b in zip(got, got[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in got]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_streams, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_streams, world, rank)
    produced = 0
    checksum = 0.0
    for s in range(lo, hi):
        rng = np.random.default_rng(1000 + s)          # data depends on the GLOBAL stream id
        x = rng.uniform(-1, 1, 2 * 3000).astype(np.float32)
        r = O.OracleFir(2, 44100, 48000, 3, 1).process(x, 1024)
        produced += len(r["out"])
        checksum += float(np.sum(r["out"].astype(np.float64)))
    dist.barrier()
    seconds = 0.5 + rank                               # pretend rank 1 is the slow one
    total = all_reduce_scalar(dist, produced, "sum")
    t_max = all_reduce_scalar(dist, seconds, "max")
    thr = job_throughput(dist, produced, seconds)
    csum = all_reduce_scalar(dist, checksum, "sum")
    q.put((rank, lo, hi, produced, total, t_max, thr, csum))
    dist.destroy_process_group()


def test_two_ranks_over_gloo_cover_all_streams_once():
    import torch.multiprocessing as mp
    n_streams, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth
    produced = 0
    checksum = 0.0
    for s in range(n_streams):
        rng = np.random.default_rng(1000 + s)
        x = rng.uniform(-1, 1, 2 * 3000).astype(np.float32)
        r = O.OracleFir(2, 44100, 48000, 3, 1).process(x, 1024)
        produced += len(r["out"])
        checksum += float(np.sum(r["out"].astype(np.float64)))
    assert [(r[1], r[2]) for r in res] == [shard_range(n_streams, world, k) for k in range(world)]
    assert sum(r[3] for r in res) == produced
    for r in res:
        assert r[4] == produced                      # all-reduced total, same on every rank
        assert r[5] == 1.5                           # max over ranks of the time
        assert abs(r[6] - produced / 1.5) < 1e-9     # whole-job throughput
        assert abs(r[7] - checksum) < 1e-6 * max(1.0, abs(checksum))
