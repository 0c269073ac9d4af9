#!/usr/bin/env python
"""BASELINE config 5 (first half): N independent 1080p streams of the scene-labeling CBinfer model
on this rank's GPU (launch under torchrun for several GPUs: streams are sharded per rank, no
collective).  One JSON line from rank 0."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import cbinfer_b200 as cb
from cbinfer_b200 import models, video, streams
from cbinfer_b200.benchtools import time_frames, median


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=64, help="total streams over all ranks")
    ap.add_argument("--chunk", type=int, default=8, help="streams batched per model call")
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--rate", type=float, default=0.05)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mine = streams.shard_streams(args.streams, world, rank)
    base = models.sceneLabelingBaseline().to(dev)
    total_ms = 0.0
    for c0 in range(0, len(mine), args.chunk):
        ids = mine[c0:c0 + args.chunk]
        m = models.sceneLabelingCBinfer(base, experimentIdx=6, threshold=0.02, clonePoolOutput=False,
                                        candidateDetect=True)
        fr = [f.to(dev) for f in video.sequence(len(ids), 1080, 1920, args.frames, args.rate, seed=ids[0])]
        models.calibrateThresholds(base, m, fr[0], factor=0.02)
        ms, _ = time_frames(m, fr, warm=3)
        total_ms += median(ms)
        del m, fr
        torch.cuda.empty_cache()
    fps, worst = streams.whole_job_rate(len(mine), total_ms, dev)
    if rank == 0:
        print(json.dumps({"workload": "%d x 1080p scene CBinfer streams, %d%% change" % (args.streams, args.rate * 100),
                          "n_gpus": world, "streams_per_gpu": len(mine), "chunk": args.chunk,
                          "ms_per_frame_of_all_local_streams": round(worst, 3), "frames_per_s": round(fps, 1)}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
