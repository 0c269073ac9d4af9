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


def _spatial_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cbinfer_b200 import models, spatial
    torch.manual_seed(0)
    base = models.sceneLabelingBaseline().eval()          # dense CPU stand-in for the CB model
    g = torch.Generator().manual_seed(1)
    frame = torch.rand(1, 3, 96, 40, generator=g)
    sp = spatial.SpatialSplit(base, 96, world, rank, halo=24, stride=4)
    lo, hi = sp.band
    with torch.no_grad():
        full = base(frame)
        got = sp(frame[:, :, lo:hi].contiguous())
    q.put((rank, sp.band, sp.slab, float((got - full).abs().max()), tuple(got.shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_spatial_split_world2_matches_full_frame():
    """row-band split with a 24-row input halo reproduces the full-frame result exactly (dense CPU
    model as the stand-in; the GPU run of the CB model is benchmarks/split_4k.py)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_spatial_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == (0, 48) and res[1][1] == (48, 96)
    assert res[0][2] == (0, 72) and res[1][2] == (24, 96)
    for r in res:
        assert r[3] < 1e-5 and r[4] == (1, 8, 24, 10)


def test_band_partition_properties():
    from cbinfer_b200 import spatial
    for H in (2160, 1080, 96):
        for world in (1, 2, 4, 8):
            bands = [spatial.band_rows(H, world, r) for r in range(world)]
            assert bands[0][0] == 0 and bands[-1][1] == H
            assert all(b[1] == c[0] for b, c in zip(bands[:-1], bands[1:]))
            assert all(b[0] % 4 == 0 for b in bands)
