"""Timing helpers shared by bench.py and benchmarks/*: CUDA-event timing of per-frame model calls,
eager or as one CUDA-graph replay per frame (mirrors the role of the reference's
`poseDetection/evalTools.py:7-35` inferFramesetBenchmark, with device-side timing)."""
import torch

from .runtime import FrameGraph


def time_frames(model, frames, warm=3, graph=True, flush_l2=False):
    """Run `frames` (list of device tensors) through `model` in order; returns (ms_per_frame list
    for frames[warm:], last output).  With graph=True frames[0..warm) run eagerly first (state
    allocation, full first frame), then one graph replay per frame."""
    assert len(frames) > warm >= 1
    with torch.no_grad():
        x = frames[0].clone()
        out = model(x)
        for i in range(1, warm):
            x.copy_(frames[i])
            out = model(x)
        torch.cuda.synchronize()
        g = FrameGraph(model, x) if graph else None
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=x.device) if flush_l2 else None
        ms = []
        for i in range(warm, len(frames)):
            x.copy_(frames[i])
            if flush is not None:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = g.replay() if g is not None else model(x)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
    return ms, out


def median(v):
    s = sorted(v)
    return s[len(s) // 2]
