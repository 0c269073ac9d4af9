"""Evaluation helpers with the interface of the reference's ``poseDetection/evalTools.py`` (symlinked
into ``sceneLabeling/``): frame loops, the min-of-N benchmark, layer listing, CSV tables.  Torch-2
idioms (``torch.no_grad`` instead of ``Variable(volatile=True)``), CUDA-event timing beside the
reference's wall clock, otherwise the same call signatures and return values.
"""
import csv
import os
import timeit

import torch

from . import CBConv2d, clearMemory


def inferFrameset(model, frames, cuda=True, preprocessor=None):
    """reference evalTools.py:37-49: clear state, run every frame, return the last output."""
    clearMemory(model)
    out = None
    with torch.no_grad():
        for frame in frames:
            x = preprocessor(frame) if preprocessor is not None else frame
            if cuda:
                x = x.cuda(non_blocking=True)
            out = model(x)
    if cuda:
        torch.cuda.synchronize()
    return out


def inferFramesetBenchmark(model, frames, cuda=True, preprocessor=None, repeat=3):
    """reference evalTools.py:7-35: setup = clear + all frames but the last (+ sync), stmt = the last
    frame (+ sync), result = min over `repeat` runs in seconds.  Also returns the CUDA-event time."""
    xs = [preprocessor(f) if preprocessor is not None else f for f in frames]
    if cuda:
        xs = [x.cuda() for x in xs]
    state = {}

    def setup():
        clearMemory(model)
        with torch.no_grad():
            for x in xs[:-1]:
                model(x)
        if cuda:
            torch.cuda.synchronize()

    def stmt():
        with torch.no_grad():
            if cuda:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                model(xs[-1])
                b.record()
                torch.cuda.synchronize()
                state.setdefault("ev", []).append(a.elapsed_time(b) * 1e-3)
            else:
                model(xs[-1])

    times = []
    for _ in range(repeat):
        setup()
        times.append(timeit.timeit(stmt, number=1))
    return (min(times), min(state["ev"])) if cuda else (min(times), None)


def inferNextFrameBenchmark(model, frame, numIter=3):
    """reference poseDetection/eval03.py:87-106: time ONE next frame from the model's current state
    - snapshot getStateTensors(), and before every repetition restore the snapshot with copy_.  The
    modules notice state written from outside their kernels (tensor version) and fall back to the
    dense scan / rebuild their operand planes for that frame, so every repetition does the same work
    and gives the same result.  Returns (min wall seconds, min CUDA-event seconds)."""
    from . import getStateTensors
    snapshot = [t.clone() for t in getStateTensors(model)]
    frame = frame.cuda()
    torch.cuda.synchronize()
    wall, dev = [], []
    for _ in range(numIter):
        for now, prev in zip(getStateTensors(model), snapshot):
            now.copy_(prev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = timeit.default_timer()
        a.record()
        with torch.no_grad():
            model(frame)
        b.record()
        torch.cuda.synchronize()
        wall.append(timeit.default_timer() - t0)
        dev.append(a.elapsed_time(b) * 1e-3)
    return min(wall), min(dev)


def getCBconvLayers(model):
    """reference evalTools.py:85-103 in spirit: the CBConv2d modules in forward order."""
    return [m for m in model.modules() if type(m) is CBConv2d]


def changeStatistics(model):
    """number of changed output pixels of every CB conv on the last frame (one host sync each)."""
    return [int(m._scratch["count"].item()) if m._scratch is not None else 0 for m in getCBconvLayers(model)]


def writeTable(filename, rows, header=None, directory="results"):
    """reference evalTools.py:105-125: dump rows to <directory>/<filename>.csv."""
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, filename if filename.endswith(".csv") else filename + ".csv")
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        if header:
            w.writerow(header)
        w.writerows(rows)
    return path
