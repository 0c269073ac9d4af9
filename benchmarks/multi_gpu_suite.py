#!/usr/bin/env python
"""One launch, several multi-GPU measurements (each `torchrun` start costs ~1 min of box time on every
GPU of the call, so the N-GPU rows of BASELINE config 5 and of the metric's "vs frame-change rate"
axis are taken in ONE process group):

  rates     frames/s of the bench workload (8 x 640x480 streams per GPU, weak scaling) at change
            rates 1 / 5 / 20 / 100 %
  streams   BASELINE configs[4]: 64 x 1080p streams in total, sharded 64/N per GPU (strong scaling)
  split4k   one 3840x2160 stream in N row bands (cbinfer_b200/spatial.py) at 20 / 100 % change,
            checked bit for bit against the full-frame model on rank 0
  e2e       the host-fed pipeline (runtime.FramePipeline) with fp32 and uint8 pinned host frames,
            next to the COPY-ONLY rate of the same buffers on all ranks at once (the host-side
            ceiling of the box: PCIe + host memory), and the host facts that explain it

Launch: python -m torch.distributed.run --nproc-per-node N benchmarks/multi_gpu_suite.py [--sections ...]
        (N = 1: plain `python benchmarks/multi_gpu_suite.py`).  Rank 0 prints one JSON line per row.
All rates are device-timed (CUDA events, barrier + synchronize on both sides), max over ranks.
"""
import argparse
import json
import os
import sys
import time
import types

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch
import torch.distributed as dist

import bench as B
import cbinfer_b200 as cb
from cbinfer_b200 import models, video, streams, spatial, runtime


def emit(rank, row):
    if rank == 0:
        print(json.dumps(row), flush=True)


def barrier(world):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def bench_args(**kw):
    a = dict(gemm="auto", dense_scan=False, threshold_factor=0.02)
    a.update(kw)
    return types.SimpleNamespace(**a)


def timed_rate(step, S, K, world, dev, min_s=0.1, warm=5):
    """median K-step block (device time, max over ranks) -> whole-job frames/s, ms per step"""
    t = 2
    for _ in range(warm):
        step(t)
        t += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    blocks = []
    while True:
        barrier(world)
        e0.record()
        for _ in range(K):
            step(t)
            t += 1
        e1.record()
        barrier(world)
        fps, ms = streams.whole_job_rate(S * K, e0.elapsed_time(e1), dev)
        blocks.append((ms, fps))
        if sum(b[0] for b in blocks) >= min_s * 1e3 or len(blocks) >= 100:
            break
    blocks.sort()
    ms, fps = blocks[len(blocks) // 2]
    return fps, ms / K, len(blocks)


def scene_run(S, H, W, rate, K, world, rank, dev, nframes=12, tag=""):
    """the bench's step (one graph per frame slot: first-layer detection in place + the rest)"""
    base = models.sceneLabelingBaseline().to(dev)
    my = streams.shard_streams(S * world, world, rank)
    frames = [f.to(dev) for f in video.sequence(S, H, W, nframes, rate, "block", seed=my[0])]
    args = bench_args()
    model, thr = B.build_model(args, base, frames[0])
    so = B.SceneStep(model, frames[0], frames[1])
    graphs = [so.capture_slot(f) for f in frames]
    period = 2 * (nframes - 1)

    def fidx(t):
        r = t % period
        return r if r < nframes else period - r

    def step(t):
        graphs[fidx(t)].replay()

    fps, ms, nb = timed_rate(step, S, K, world, dev)
    counts = [int(m._scratch["count"].item()) for m in model.modules() if type(m) is cb.CBConv2d]
    row = {"section": tag, "n_gpus": world, "streams_per_gpu": S, "height": H, "width": W, "rate": rate,
           "frames_per_s": round(fps, 1), "ms_per_step": round(ms, 4), "timed_blocks": nb, "steps_per_block": K,
           "changed_pixels_last_frame": counts}
    del graphs, so
    cb.clearMemory(model)
    torch.cuda.empty_cache()
    return row


def section_rates(a, world, rank, dev):
    for rate in a.rates:
        row = scene_run(8, 480, 640, rate, 100 if rate <= 0.2 else 40, world, rank, dev, tag="rates")
        row["scaling"] = "weak"
        emit(rank, row)


def section_streams(a, world, rank, dev):
    S = max(1, 64 // world)
    row = scene_run(S, 1080, 1920, 0.05, 20, world, rank, dev, nframes=6, tag="streams1080p")
    row["scaling"] = "strong"
    row["streams_total"] = S * world
    emit(rank, row)


def section_split4k(a, world, rank, dev):
    H, W = 2160, 3840
    base = models.sceneLabelingBaseline().to(dev)
    for rate in a.split_rates:
        frames = video.sequence(1, H, W, 7, rate)            # the same frames on every rank
        model = models.sceneLabelingCBinfer(base, experimentIdx=6, threshold=0.02, clonePoolOutput=False,
                                            candidateDetect=True)
        models.calibrateThresholds(base, model, frames[0][:, :, :480, :640].contiguous().to(dev), factor=0.02)
        sp = spatial.SpatialSplit(model, H, world, rank, halo=24, stride=4)
        lo, hi = sp.band
        bands = [f[:, :, lo:hi].contiguous().to(dev) for f in frames]
        with torch.no_grad():
            slab = sp.exchange_halo(bands[0])
            sp.forward_slab(slab)
            slab.copy_(sp.exchange_halo(bands[1]))
            sp.forward_slab(slab)
        g = runtime.FrameGraph(model, slab)
        o0 = (lo - sp.slab[0]) // 4
        ms, out = [], None
        for t in range(2, len(frames)):
            barrier(world)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            slab.copy_(sp.exchange_halo(bands[t]))
            y = g.replay()
            out = sp.gather(y[:, :, o0:o0 + (hi - lo) // 4])
            e1.record()
            torch.cuda.synchronize()
            ms.append(streams.max_over_ranks(e0.elapsed_time(e1), dev))
        ms = sorted(ms[1:])
        row = {"section": "split4k", "n_gpus": world, "rate": rate, "band_rows": hi - lo,
               "slab_rows": sp.slab[1] - sp.slab[0], "ms_per_frame_median": round(ms[len(ms) // 2], 3),
               "frames_per_s": round(1000.0 / ms[len(ms) // 2], 1)}
        if rank == 0 and a.check:
            full = models.sceneLabelingCBinfer(base, experimentIdx=6, threshold=0.02, clonePoolOutput=False,
                                               candidateDetect=True)
            for c, c2 in zip([m for m in full.modules() if type(m) is cb.CBConv2d],
                             [m for m in model.modules() if type(m) is cb.CBConv2d]):
                c.threshold = c2.threshold
            with torch.no_grad():
                for f in frames:
                    ref = full(f.to(dev))
            row["max_abs_diff_vs_full_frame"] = float((out - ref).abs().max())
            cb.clearMemory(full)
            del full
        emit(rank, row)
        del g, bands, slab
        cb.clearMemory(model)
        torch.cuda.empty_cache()


def host_facts(local):
    facts = {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    try:
        facts["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
    except OSError:
        pass
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        facts["gpu_numa_node"] = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
    except Exception as e:                                   # noqa: BLE001 - diagnostics only
        facts["gpu_numa_node"] = "n/a (%s)" % type(e).__name__
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemTotal"):
                facts["host_mem_gb"] = round(int(line.split()[1]) / 1e6, 1)
    except OSError:
        pass
    return facts


def section_e2e(a, world, rank, local, dev):
    S, H, W, K, nframes = 8, 480, 640, 100, 12
    base = models.sceneLabelingBaseline().to(dev)
    my = streams.shard_streams(S * world, world, rank)
    frames_cpu = video.sequence(S, H, W, nframes, 0.05, "block", seed=my[0])
    period = 2 * (nframes - 1)

    def fidx(t):
        r = t % period
        return r if r < nframes else period - r

    emit(rank, dict(section="e2e_host", n_gpus=world, **host_facts(local)))
    for kind in ("f32", "u8"):
        if kind == "f32":
            pin = [f.pin_memory() for f in frames_cpu]
        else:
            pin = [(f * 255.0).round().clamp(0, 255).to(torch.uint8).pin_memory() for f in frames_cpu]
        first_dev = pin[0].to(dev)
        # ---- copy-only ceiling: the same pinned buffers, H2D on every rank at once, no compute ----
        dst = torch.empty_like(first_dev)
        for rep in range(2):
            barrier(world)
            t0 = time.perf_counter()
            n = 0
            while n < 3 * K:
                dst.copy_(pin[fidx(n)], non_blocking=True)
                n += 1
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            barrier(world)
        bytes_step = pin[0].numel() * pin[0].element_size()
        worst = streams.max_over_ranks(wall, dev)
        copy_gbs = bytes_step * 3 * K / worst / 1e9
        args = bench_args()
        model, _ = B.build_model(args, base, frames_cpu[0].to(dev))
        if kind == "u8":
            first = [m for m in model.modules() if type(m) is cb.CBConv2d][0]
            first.inputNorm = (255.0, 0.0)
        pipe = runtime.FramePipeline(model, first_dev, depth=a.depth)
        for i in range(1, 6):
            pipe.submit(pin[fidx(i)])
        pipe.drain()
        i0, walls = 6, []
        for rep in range(4):
            barrier(world)
            t0 = time.perf_counter()
            last = None
            for i in range(i0, i0 + K):
                last = pipe.submit(pin[fidx(i)])
            pipe.wait(last)
            pipe.drain()
            w = time.perf_counter() - t0
            barrier(world)
            walls.append(streams.max_over_ranks(w, dev))
            i0 += K
        walls.sort()
        w = walls[len(walls) // 2]
        # host-only cost of one submit (python + CUDA API calls): the same loop on a 1-stream 8x8 frame
        emit(rank, {"section": "e2e", "n_gpus": world, "host_frames": kind, "frames_per_s": round(S * K * world / w, 1),
                    "ms_per_step": round(w / K * 1e3, 4), "h2d_bytes_per_step": bytes_step,
                    "h2d_gbs_per_gpu": round(bytes_step * K / w / 1e9, 2),
                    "copy_only_gbs_per_gpu": round(copy_gbs, 2), "copy_only_gbs_all_gpus": round(copy_gbs * world, 1),
                    "pipeline_depth": a.depth})
        del pipe
        cb.clearMemory(model)
        torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sections", default="rates,streams,split4k,e2e")
    ap.add_argument("--rates", default="0.01,0.05,0.2,1.0")
    ap.add_argument("--split-rates", default="0.2,1.0")
    ap.add_argument("--depth", type=int, default=3)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--bind", action="store_true", help="pin each rank to its own slice of the host cores")
    a = ap.parse_args()
    a.rates = [float(r) for r in a.rates.split(",") if r]
    a.split_rates = [float(r) for r in a.split_rates.split(",") if r]
    a.check = not a.no_check
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    if a.bind:
        runtime.bind_host_cores(local, world)
    torch.set_num_threads(max(1, min(4, (os.cpu_count() or 4) // max(world, 1))))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    with torch.no_grad():
        for sec in a.sections.split(","):
            if sec == "rates":
                section_rates(a, world, rank, dev)
            elif sec == "streams":
                section_streams(a, world, rank, dev)
            elif sec == "split4k":
                section_split4k(a, world, rank, dev)
            elif sec == "e2e":
                section_e2e(a, world, rank, local, dev)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
