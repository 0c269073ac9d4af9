#!/usr/bin/env python
"""BASELINE config 4: OpenPose-style CPM (VGG-19 stem + T stages, random-init) CBinfer bf16 at
368x368 synthetic video; dense cuDNN beside it.  One JSON line."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cbinfer_b200 as cb
from cbinfer_b200 import models, video
from cbinfer_b200.benchtools import time_frames, median


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=6)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--rate", type=float, default=0.05)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--threshold-factor", type=float, default=0.02)
    args = ap.parse_args()
    dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[args.dtype]
    torch.backends.cudnn.benchmark = True
    pose = models.PoseModel(T=args.T).cuda().to(dt).to(memory_format=torch.channels_last)
    fr = [f.cuda().to(dt) for f in video.sequence(args.batch, 368, 368, args.frames, args.rate, lo=-0.5, hi=0.5)]
    res = {"model": "CPM T=%d" % args.T, "dtype": args.dtype, "batch": args.batch, "rate": args.rate,
           "convs": sum(1 for m in pose.modules() if isinstance(m, torch.nn.Conv2d))}
    ms, dense_out = time_frames(pose, [f.contiguous(memory_format=torch.channels_last) for f in fr], warm=3)
    res["dense_cudnn_ms"] = round(median(ms), 4)
    for cand, par in ((False, False), (True, False), (True, True)):
        m = models.poseModelCBinfer(pose, threshold=0.0)
        m.parallelBranches = par               # branch 2 of every stage on a side stream
        if cand:
            models.enableCandidateDetection(m)
        # fixed per-layer thresholds: factor * input range measured on frame 0 (dense hooks)
        feeds = {}
        convs = [mm for mm in pose.modules() if isinstance(mm, torch.nn.Conv2d)]
        hooks = [c.register_forward_hook(lambda mod, inp, out, i=i: feeds.__setitem__(i, float(inp[0].max() - inp[0].min())))
                 for i, c in enumerate(convs)]
        with torch.no_grad():
            pose(fr[0])
        for h in hooks:
            h.remove()
        for i, c in enumerate([mm for mm in m.modules() if type(mm) is cb.CBConv2d]):
            c.threshold = args.threshold_factor * feeds[i]
        ms, out = time_frames(m, fr, warm=3)
        key = ("cb_candidates_parallel_branches" if par else "cb_candidates") if cand else "cb_dense_scan"
        res[key + "_ms"] = round(median(ms), 4)
        with torch.no_grad():
            ref = pose(fr[-1])
        res[key + "_max_abs_diff_vs_dense"] = float(max((a.float() - b.float()).abs().max() for a, b in zip(out, ref)))
        res[key + "_ref_max_abs"] = float(max(b.float().abs().max() for b in ref))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
