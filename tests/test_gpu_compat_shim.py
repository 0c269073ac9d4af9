"""The reference's own FFI symbols, bound the way the reference binds them (cdef prototypes of
pycbinfer/conv2d_cg.py:6-38 and conv2d_fg.py:13-29, launch geometry of its wrappers), against TWO
sets of libraries that export them: the compatibility shim built from cbinfer_b200/csrc/compat_shim.cu
on top of libcbinfer_sm100.so (cbinfer_b200/compat/), and the UNMODIFIED reference libraries compiled
from /root/reference where they lie (oracle/_ref).  Identical call sequences, identical inputs:
change maps, feedback state, im2col matrix, scattered output and pooled map must be bit-identical,
the fine-grained update within 1e-4 (atomic summation order).  This is the evidence that the
unmodified reference python files can run on the sm_100a kernels by swapping the three .so files."""
import ctypes
import os
import platform

import numpy as np
import pytest
import torch

from tests.util import REPO, ref_lib, rand_tensor, perturb, vp

pytestmark = pytest.mark.gpu


def shim_lib(name):
    p = os.path.join(REPO, "cbinfer_b200", "compat", "%s_%s.so" % (name, platform.machine()))
    if not os.path.exists(p):
        from cbinfer_b200 import build
        build.build_compat()
    return ctypes.CDLL(p)


def _pair(name):
    ref = ref_lib(name)
    if ref is None:
        pytest.skip("oracle/_ref not built (reference checkout absent at build time)")
    return shim_lib(name), ref


def _change_detection(lib, x, prev, k, thr, update):
    """conv2d_cg.py:100-122"""
    H, W = x.shape[-2:]
    cmap = torch.zeros(H, W, dtype=torch.int8, device="cuda")
    st = prev.clone()
    lib.changeDetection(1, 1, (H * W - 1) // 128 + 1, 1, 1, 128, vp(x), vp(st), vp(cmap), W, H, x.shape[1],
                        (k - 1) // 2, (k - 1) // 2, ctypes.c_float(thr), ctypes.c_bool(update))
    torch.cuda.synchronize()
    return cmap, st


@pytest.mark.parametrize("half", [False, True])
def test_shim_change_detection_equals_reference_library(half):
    shim, ref = _pair("cbconv2d_cg_half_backend" if half else "cbconv2d_cg_backend")
    dt = "f16" if half else "f32"
    for (C, H, W, k, thr) in ((16, 40, 30, 3, 0.1), (3, 15, 20, 7, 0.4), (64, 33, 65, 1, 0.6), (5, 9, 130, 5, 0.5)):
        prev = rand_tensor((1, C, H, W), dt, C + H)
        x = perturb(prev, 0.08, C + 3)
        for update in (False, True):
            m1, s1 = _change_detection(shim, x, prev, k, thr, update)
            m2, s2 = _change_detection(ref, x, prev, k, thr, update)
            assert torch.equal(m1, m2), (C, H, W, k, update)
            assert torch.equal(s1.view(torch.int16 if half else torch.int32),
                               s2.view(torch.int16 if half else torch.int32)), (C, H, W, k, update)
            assert int(m1.sum()) > 0


def test_shim_reference_kat_15_pixels():
    """genTestData + changeDetection_test1 (conv2d_cg.py:84-98,136-142): two of the four perturbed points
    exceed the threshold; with a 3x3 filter the dilated map holds 6 + 9 pixels (one point sits on the border)"""
    shim = shim_lib("cbconv2d_cg_backend")
    torch.manual_seed(10)
    x = torch.randn(1, 16, 400, 300, device="cuda")
    prev = x.clone()
    for c, y, xx, d in ((0, 0, 4, 1.00), (1, 6, 9, 0.05), (2, 10, 4, -11.00), (1, 6, 19, -0.05)):
        prev[0, c, y, xx] += d
    cmap, _ = _change_detection(shim, x, prev, 3, 0.1, False)
    assert int(cmap.sum()) == 15
    assert sorted(torch.nonzero(cmap.view(-1)).view(-1).tolist())[:3] == [3, 4, 5]


def test_shim_propagation_gather_scatter_pool_equal_reference_library():
    shim, ref = _pair("cbconv2d_cg_backend")
    g = torch.Generator().manual_seed(5)
    C, Cout, H, W, k = 6, 10, 37, 45, 5
    # changePropagation (conv2d_cg.py:159-177)
    raw = (torch.rand(H, W, generator=g) < 0.03).to(torch.int8).cuda()
    outs = []
    for lib in (shim, ref):
        o = torch.zeros(H, W, dtype=torch.int8, device="cuda")
        lib.changePropagation(1, 1, (H * W - 1) // 128 + 1, 1, 1, 128, vp(raw), vp(o), W, H,
                              (k - 1) // 2, (k - 1) // 2)                          # geometry of conv2d_cg.py:166-170
        torch.cuda.synchronize()
        outs.append(o)
    assert torch.equal(outs[0], outs[1]) and int(outs[0].sum()) > int(raw.sum())
    # genXMatrix (conv2d_cg.py:239-261), updateOutput (:292-313), maxPool2d (:58-82)
    x = rand_tensor((1, C, H, W), "f32", 1)
    idx = torch.nonzero(outs[0].view(-1)).view(-1).int().cuda()
    n = idx.numel()
    tz = 128 // (k * k)
    Yt = rand_tensor((Cout, n), "f32", 3)
    po0 = rand_tensor((1, Cout, H, W), "f32", 4)
    xe = rand_tensor((1, C, 36, 52), "f32", 5)
    idx2 = torch.nonzero(torch.rand(36 * 52, generator=g) < 0.2).view(-1).int().cuda()
    res = []
    for lib in (shim, ref):
        X = torch.zeros(n, C * k * k, device="cuda")
        lib.genXMatrix(1, 1, (n - 1) // tz + 1, k, tz, k, vp(X), vp(x), vp(idx), k, k, C, W, H, n)
        po = po0.clone()
        lib.updateOutput(1, 1, (n * Cout - 1) // 1024 + 1, 1, 1, 1024, vp(Yt), vp(po), vp(idx), H * W, n, Cout,
                         ctypes.c_bool(True))
        st = torch.full((1, C, 18, 26), float("inf"), device="cuda")
        lib.maxPool2d((idx2.numel() - 1) // 64 + 1, 64, vp(xe), vp(st), vp(idx2), idx2.numel(), C, 36, 52, 18, 26, 2, 2)
        torch.cuda.synchronize()
        res.append((X, po, st))
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a, b)


def test_shim_half_gather_scatter_pool_equal_reference_library():
    shim, ref = _pair("cbconv2d_cg_half_backend")
    g = torch.Generator().manual_seed(6)
    C, Cout, H, W, k = 8, 12, 30, 44, 3
    x = rand_tensor((1, C, H, W), "f16", 1)
    idx = torch.nonzero(torch.rand(H * W, generator=g) < 0.2).view(-1).int().cuda()
    n = idx.numel()
    tz = 128 // (k * k)
    Yt = rand_tensor((Cout, n), "f16", 3)
    po0 = rand_tensor((1, Cout, H, W), "f16", 4)
    idx2 = torch.nonzero(torch.rand(H * W, generator=g) < 0.2).view(-1).int().cuda()
    res = []
    for lib in (shim, ref):
        X = torch.zeros(n, C * k * k, device="cuda", dtype=torch.float16)
        lib.genXMatrix(1, 1, (n - 1) // tz + 1, k, tz, k, vp(X), vp(x), vp(idx), k, k, C, W, H, n)
        po = po0.clone()
        lib.updateOutput(1, 1, (n * Cout - 1) // 1024 + 1, 1, 1, 1024, vp(Yt), vp(po), vp(idx), H * W, n, Cout,
                         ctypes.c_bool(False))
        st = torch.full((1, C, H // 2, W // 2), float("inf"), device="cuda", dtype=torch.float16)
        lib.maxPool2d((idx2.numel() - 1) // 64 + 1, 64, vp(x), vp(st), vp(idx2), idx2.numel(), C, H, W, H // 2, W // 2, 2, 2)
        torch.cuda.synchronize()
        res.append((X, po, st))
    for a, b in zip(res[0], res[1]):
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))


def test_shim_fine_grained_equals_reference_library():
    shim, ref = _pair("cbconv2d_fg_backend")
    Cin, Cout, H, W, k = 5, 7, 21, 26, 3
    prev = rand_tensor((1, Cin, H, W), "f32", 1)
    x = perturb(prev, 0.1, 2)
    w = rand_tensor((Cout, Cin, k, k), "f32", 3, scale=0.3)
    out0 = rand_tensor((1, Cout, H, W), "f32", 4)
    res = []
    for lib in (shim, ref):
        diffs = torch.zeros_like(x)
        cmap = torch.zeros(x.shape, dtype=torch.int8, device="cuda")
        lib.changeDetectionFG(vp(x), vp(prev), vp(diffs), vp(cmap), x.numel(), ctypes.c_float(0.2))
        torch.cuda.synchronize()
        coords = torch.nonzero(cmap.view(-1)).view(-1).contiguous()          # conv2d_fg.py:82
        out = out0.clone()
        lib.updateOutputFG(1, 1, (coords.numel() - 1) // 128 + 1, 1, 1, 128, vp(diffs), vp(w), vp(out), vp(coords),
                           Cout, Cin, H, W, k, k, coords.numel())
        torch.cuda.synchronize()
        res.append((cmap, diffs, out))
    assert torch.equal(res[0][0], res[1][0]) and int(res[0][0].sum()) > 0
    assert torch.equal(res[0][1], res[1][1])
    np.testing.assert_allclose(res[0][2].cpu().numpy(), res[1][2].cpu().numpy(), rtol=1e-5, atol=1e-4)
