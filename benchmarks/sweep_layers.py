#!/usr/bin/env python
"""BASELINE config 3: single-layer sweep, coarse-grained (tcgen05) vs fine-grained CBConv2d vs dense
cuDNN conv, change rate 0-100 %, fp32 (3xTF32 / TF32) and bf16, on layer shapes of the two nets.
Prints one JSON line per (layer, dtype, mode, rate): ms per frame (median, one graph replay)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import torch.nn.functional as F
import cbinfer_b200 as cb
from cbinfer_b200 import video
from cbinfer_b200.benchtools import time_frames, median

LAYERS = {   # name: (Cin, Cout, k, H, W)
    "scene_L2": (16, 64, 7, 240, 320), "scene_L3": (64, 256, 7, 120, 160),
    "pose_conv1_2": (64, 64, 3, 368, 368), "pose_conv4_2": (512, 512, 3, 46, 46),
    "pose_Mconv2": (128, 128, 7, 46, 46),
}


def frames_for(Cin, H, W, n, rate, dt, B):
    g = torch.Generator().manual_seed(0)
    f = [torch.rand(B, Cin, H, W, generator=g)]
    import math
    for t in range(1, n):
        x = f[-1].clone()
        if rate > 0:
            area = rate * H * W
            bh = min(H, max(1, int(round(math.sqrt(area * 3 / 4)))))
            bw = min(W, max(1, int(round(area / bh))))
            for b in range(B):
                y0 = int(torch.randint(0, H - bh + 1, (1,), generator=g))
                x0 = int(torch.randint(0, W - bw + 1, (1,), generator=g))
                x[b, :, y0:y0 + bh, x0:x0 + bw] = torch.rand(Cin, bh, bw, generator=g)
        f.append(x)
    return [t.to(dt).cuda() for t in f]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", default=",".join(LAYERS))
    ap.add_argument("--rates", default="0,0.01,0.02,0.05,0.1,0.2,0.5,1.0")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=14)
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = True
    for name in args.layers.split(","):
        Cin, Cout, k, H, W = LAYERS[name]
        for dtn, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
            torch.manual_seed(0)
            conv = nn.Conv2d(Cin, Cout, k, padding=k // 2).cuda().to(dt).eval()
            for rate in [float(r) for r in args.rates.split(",")]:
                fr = frames_for(Cin, H, W, args.frames, rate, dt, args.batch)
                row = {"layer": name, "shape": [Cin, Cout, k, H, W], "dtype": dtn, "rate": rate, "batch": args.batch}
                modes = [("cg", "auto")] + ([("cg_tf32", "tc"), ("fg", None)] if dtn == "f32" else [])
                for label, gm in modes:
                    m = cb.CBConv2d(conv, 0.0)
                    m.withReLU = True
                    if label == "fg":
                        m.finegrained = True
                    else:
                        m.feedbackLoop = True
                        m.gemmMode = gm
                    ms, _ = time_frames(m, fr, warm=3, graph=True)
                    row[label + "_ms"] = round(median(ms), 4)
                for tf32 in ((False, True) if dtn == "f32" else (True,)):
                    torch.backends.cudnn.allow_tf32 = tf32
                    cl = conv.to(memory_format=torch.channels_last)
                    xs = [x.contiguous(memory_format=torch.channels_last) for x in fr]
                    dense = lambda x: F.relu(cl(x))
                    ms, _ = time_frames(dense, xs, warm=3, graph=True)
                    row["cudnn_%s_ms" % ("tf32" if tf32 and dtn == "f32" else "fp32" if dtn == "f32" else "bf16")] = round(median(ms), 4)
                print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
