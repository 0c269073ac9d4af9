"""GPU parity at the BASELINE configurations against the reference FLOW: the unmodified reference
CUDA kernels (oracle/_ref, compiled from pycbinfer/cbconv2d_cg[_half]_backend.cu where they lie)
driven by tests/ref_flow.py -- a torch-2 restatement of pycbinfer/conv2d.py:178-259 and :49-78 --
with fp32 / fp16 torch.matmul (TF32 off) for the GEMM.  Thresholds > 0 throughout.

Two kinds of comparison:
  * TEACHER-FORCED, layer by layer: every CB layer of this repo is fed the reference's input of
    that layer.  Change index lists and the feedback state must then be bit-identical on every
    frame (same inputs, same state history), pooled maps bit-identical, conv outputs within the
    contraction tolerance (1e-4 fp32, 2e-3 fp16, relative to max |ref|).
  * FREE-RUNNING: the benchmarked configuration itself (candidate detection, pool-fused detection,
    masked 1x1 contraction, tiled contraction, detectInput + CUDA graph; bench.parity_check).
    Layer-1 index lists are bit-identical (same frames).  Deeper layers see inputs that differ
    from the reference's by the contraction tolerance, so a pixel whose change magnitude lies
    within that distance of the threshold may be classified differently ("threshold flip"); the
    effect of a flip on the output is bounded by threshold * sum|w| (the approximation CBinfer
    makes anyway), not by the contraction tolerance.  The bar: final outputs within 1e-4 when no
    flip occurred, else within 5e-3, and at most a handful of differing mask bits.
"""
import argparse

import pytest
import torch
import torch.nn as nn

from tests import ref_flow

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _setup():
    if ref_flow.libs() is None:
        pytest.skip("oracle/_ref not built (reference checkout absent at build time)")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _ours_layers(base, thresholds, tdt):
    """this repo's conversion of a plain nn.Sequential, plain reference semantics (dense scan per
    layer, no candidates), as a list aligned with ref_flow.convert_sequential"""
    import cbinfer_b200 as cb
    m = cb.convertPools(cb.convert(base, threshold=0.1))
    convs = [c for c in m.modules() if type(c) is cb.CBConv2d]
    for c, th in zip(convs, thresholds):
        c.threshold = th
        c.feedbackLoop = True
    return list(m.children())


def _teacher_forced(base, thresholds, frames, tol, tdt=torch.float32):
    """run the reference flow; feed each of our layers the reference's input of that layer"""
    import cbinfer_b200 as cb
    refs = ref_flow.convert_sequential(base, thresholds)
    ours = _ours_layers(base, thresholds, tdt)
    assert len(refs) == len(ours)
    worst = 0.0
    for t, f in enumerate(frames):
        x_ref = f
        for li, (r, o) in enumerate(zip(refs, ours)):
            if isinstance(r, ref_flow.RefCBConv2d):
                inp = x_ref[1] if type(x_ref) == tuple else x_ref
                y_ref = r.forward(inp)
                y = o(inp)
                y_t = y[1] if type(y) == tuple else y
                yr_t = y_ref[1] if type(y_ref) == tuple else y_ref
                n = int(o._scratch["count"])
                assert torch.equal(o.lastChangeIndexes(), r.changeIndexes), (t, li, n, r.changeIndexes.numel())
                assert torch.equal(o.prevInput, r.prevInput), (t, li)
                scale = float(yr_t.float().abs().max()) + 1e-30
                err = float((y_t.float() - yr_t.float()).abs().max()) / scale
                worst = max(worst, err)
                assert err <= tol, (t, li, err)
            elif isinstance(r, ref_flow.RefCBPoolMax2d):
                y_ref = r.forward(x_ref)
                y = o(('changeIndexes', x_ref[1], x_ref[2]))
                y_t = y[1] if type(y) == tuple else y
                assert torch.equal(y_t, y_ref), (t, li)
            else:
                y_ref = r(x_ref)
            x_ref = y_ref
    return worst


def _scene(dev, tdt=torch.float32):
    from cbinfer_b200 import models
    return models.sceneLabelingBaseline().to(dev).to(tdt)


def _thresholds(base, frame, factor=0.02):
    feeds, hooks = {}, []
    convs = [m for m in base.modules() if type(m) is nn.Conv2d]
    for i, c in enumerate(convs):
        hooks.append(c.register_forward_hook(
            lambda mod, inp, out, i=i: feeds.__setitem__(i, float(inp[0].float().max() - inp[0].float().min()))))
    with torch.no_grad():
        base(frame)
    for h in hooks:
        h.remove()
    return [factor * feeds[i] for i in range(len(convs))]


def test_scene_640x480_teacher_forced_vs_reference_flow():
    """BASELINE configs[1] size, 5 % block change, calibrated thresholds, 7 frames."""
    from cbinfer_b200 import video
    dev = torch.device("cuda")
    base = _scene(dev)
    frames = [f.to(dev) for f in video.sequence(1, 480, 640, 7, 0.05, "block", seed=11)]
    thr = _thresholds(base, frames[0])
    worst = _teacher_forced(base, thr, frames, tol=1e-4)
    print("scene 640x480 teacher-forced: worst conv rel err %.2e" % worst)


def test_scene_1080p_teacher_forced_vs_reference_flow():
    """one 1080p stream (BASELINE configs[4] frame size), 4 frames."""
    from cbinfer_b200 import video
    dev = torch.device("cuda")
    base = _scene(dev)
    frames = [f.to(dev) for f in video.sequence(1, 1080, 1920, 4, 0.05, "block", seed=12)]
    thr = _thresholds(base, frames[0][:, :, :480, :640].contiguous())
    worst = _teacher_forced(base, thr, frames, tol=1e-4)
    print("scene 1080p teacher-forced: worst conv rel err %.2e" % worst)


def test_bench_configuration_free_running_vs_reference_flow():
    """the exact bench model (640x480, 5 % block, calibrateThresholds, candidate + masked + tiled +
    detectInput + graph), 2 streams, 8 frames, against the reference flow (bench.parity_check)."""
    import bench
    from cbinfer_b200 import models, video
    dev = torch.device("cuda")
    args = argparse.Namespace(dense_scan=False, gemm="auto", threshold_factor=0.02, height=480, width=640)
    base = _scene(dev)
    frames_cpu = video.sequence(2, 480, 640, 10, 0.05, "block", seed=0)
    _, thresholds = bench.build_model(args, base, frames_cpu[0].to(dev))
    res = bench.parity_check(args, base, thresholds, frames_cpu, dev, nstreams=2, nframes=10)
    print("bench configuration vs reference flow:", res)
    assert res["layer1_index_lists_bit_exact"]
    flips = res["deeper_layer_mask_bits_differing_last_frame"]
    assert flips <= 64, flips
    assert res["parity_max_rel"] <= (1e-4 if flips == 0 else 5e-3), res


def test_cpm_368_fp16_teacher_forced_vs_reference_flow():
    """OpenPose-style CPM (BASELINE configs[3] topology and 368x368 size) in fp16 against the
    reference's HALF backend, T=2: the VGG trunk (12 convs, 3 pools), both branches of stage 1 and of
    stage 2 (fed by the reference's concatenated features), thresholds > 0, 3 frames.  bf16 has no
    reference kernel (same recipe as fp16, tests/test_gpu_ops.py); its CPM run is checked against
    dense fp32 in test_gpu_modules.py."""
    from cbinfer_b200 import models
    dev = torch.device("cuda")
    tdt = torch.float16
    pose = models.PoseModel(T=2).to(dev).to(tdt)
    g = torch.Generator().manual_seed(5)
    f0 = (torch.rand(1, 3, 368, 368, generator=g) - 0.5)
    frames = [f0]
    for t in range(1, 3):
        f = frames[-1].clone()
        y0, x0 = 40 * t, 60 * t
        f[:, :, y0:y0 + 80, x0:x0 + 100] = torch.rand(1, 3, 80, 100, generator=g) - 0.5
        frames.append(f)
    frames = [f.to(dev).to(tdt) for f in frames]
    blocks = ["model0", "model1_1", "model1_2", "model2_1", "model2_2"]
    thr = {}
    # thresholds: 2 % of each layer's dense input range on the first frame, block by block
    with torch.no_grad():
        feat = pose.model0(frames[0])
        L1, S1 = pose.model1_1(feat), pose.model1_2(feat)
        cat = torch.cat([L1, S1, feat], 1)
    feeds = {"model0": frames[0], "model1_1": feat, "model1_2": feat, "model2_1": cat, "model2_2": cat}
    for b in blocks:
        thr[b] = _thresholds(getattr(pose, b), feeds[b])
    refs = {b: ref_flow.convert_sequential(getattr(pose, b), thr[b]) for b in blocks}
    ours = {b: _ours_layers(getattr(pose, b), thr[b], tdt) for b in blocks}

    def run_block(b, x, t):
        xr = x
        for li, (r, o) in enumerate(zip(refs[b], ours[b])):
            if isinstance(r, ref_flow.RefCBConv2d):
                inp = xr[1] if type(xr) == tuple else xr
                yr = r.forward(inp)
                y = o(inp)
                y_t = y[1] if type(y) == tuple else y
                yr_t = yr[1] if type(yr) == tuple else yr
                n = int(o._scratch["count"])
                assert torch.equal(o.lastChangeIndexes(), r.changeIndexes), (b, t, li)
                assert torch.equal(o.prevInput, r.prevInput), (b, t, li)
                scale = float(yr_t.float().abs().max()) + 1e-30
                assert float((y_t.float() - yr_t.float()).abs().max()) / scale <= 2e-3, (b, t, li)
            elif isinstance(r, ref_flow.RefCBPoolMax2d):
                yr = r.forward(xr)
                y = o(('changeIndexes', xr[1], xr[2]))
                assert torch.equal(y[1] if type(y) == tuple else y, yr), (b, t, li)
            else:
                yr = r(xr)
            xr = yr
        return xr[1] if type(xr) == tuple else xr

    for t, f in enumerate(frames):
        feat = run_block("model0", f, t)
        L1 = run_block("model1_1", feat, t)
        S1 = run_block("model1_2", feat, t)
        cat = torch.cat([L1, S1, feat], 1)
        run_block("model2_1", cat, t)
        run_block("model2_2", cat, t)
