"""CPU, world_size 2 over gloo: the N>1 host logic (stream sharding, whole-job rate from the
slowest rank) behaves as bench.py relies on."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cbinfer_b200 import streams
    mine = streams.shard_streams(13, world, rank)
    cyc = streams.shard_streams(13, world, rank, policy="cyclic")
    rate, worst = streams.whole_job_rate(len(mine) * 10, 100.0 * (rank + 1))
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, cyc))
    dist.barrier()
    q.put((rank, mine, cyc, rate, worst, gathered))
    dist.destroy_process_group()


def test_stream_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, c0, rate0, worst0, g0), (r1, m1, c1, rate1, worst1, g1) = res
    assert sorted(m0 + m1) == list(range(13)) and not set(m0) & set(m1)       # a partition
    assert sorted(c0 + c1) == list(range(13)) and c0 == list(range(0, 13, 2))
    assert abs(len(m0) - len(m1)) <= 1
    assert worst0 == worst1 == 200.0                                           # slowest rank
    assert rate0 == rate1 == 130 / 0.2                                         # all units / max time
    assert g0 == g1


def test_single_process_identity():
    from cbinfer_b200 import streams
    assert streams.shard_streams(8, 1, 0) == list(range(8))
    rate, worst = streams.whole_job_rate(80, 40.0)
    assert rate == 2000.0 and worst == 40.0
