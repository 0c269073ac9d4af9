"""Frame ingest: the resizers of the reference's readers.  CPU part: the oracle against
the golden vectors made with PIL (and against the live PIL when importable).  GPU part: the CUDA kernels
behind the C ABI against the oracle and the golden vectors -- bit-exact for the 8-bit bicubic path."""
import os

import numpy as np
import pytest
import torch

from oracle import ingest_oracle as io
from tests.golden.make_ingest_golden import CASES

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ingest_golden.npz"))


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_bicubic_matches_pil_golden(i):
    H, W, oh, ow = CASES[i]
    got = io.pil_bicubic_u8(GOLD["in%d" % i], oh, ow)
    assert got.dtype == np.uint8 and np.array_equal(got, GOLD["bicubic%d" % i])


def test_oracle_bicubic_matches_live_pil():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(7)
    for H, W, oh, ow in [(61, 47, 23, 90), (33, 200, 33, 77), (90, 90, 368, 12)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        exp = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        assert np.array_equal(io.pil_bicubic_u8(img, oh, ow), exp)


def test_oracle_bilinear_against_scipy_interior():
    """same interpolation rule as scipy.ndimage.map_coordinates(order=1) wherever no neighbour lies outside the
    image (down-scaling: everywhere); identity for equal sizes."""
    ndimage = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(3)
    img = rng.random((50, 70, 3))
    assert np.array_equal(io.skimage_resize_bilinear(img, 50, 70), img)
    for oh, ow in [(25, 31), (49, 7)]:
        rr, cc = np.meshgrid(50 / oh * (np.arange(oh) + 0.5) - 0.5, 70 / ow * (np.arange(ow) + 0.5) - 0.5, indexing="ij")
        exp = np.stack([ndimage.map_coordinates(img[..., k], [rr, cc], order=1, mode="constant") for k in range(3)], -1)
        assert np.abs(io.skimage_resize_bilinear(img, oh, ow, clip=False) - exp).max() < 1e-12


def test_pose_normalisation_equals_divide_by_256():
    """ToTensor()/255 then mul_(255/256).add_(-0.5) (PoseDetector.py:68-72) is, for all 256 byte values,
    bit-identical to u8 / 256 - 0.5: what CBConv2d.inputNorm = (256, -0.5) computes inside the detection."""
    u = np.arange(256, dtype=np.float32)
    a = (u / np.float32(255)) * np.float32(255.0 / 256.0) + np.float32(-0.5)
    assert np.array_equal(a, u / np.float32(256) + np.float32(-0.5))


# ---- GPU ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_gpu_bicubic_bit_exact(i):
    from cbinfer_b200 import ingest
    H, W, oh, ow = CASES[i]
    img = torch.from_numpy(GOLD["in%d" % i]).cuda()
    exp = torch.from_numpy(GOLD["bicubic%d" % i]).cuda()
    assert torch.equal(ingest.resize_bicubic_u8(img, oh, ow), exp)
    # planar output (ToTensor's layout), planar / pitched input views, repeated call on the cached plan
    assert torch.equal(ingest.resize_bicubic_u8(img, oh, ow, planar=True), exp.permute(2, 0, 1))
    planar_in = img.permute(2, 0, 1).contiguous().permute(1, 2, 0)
    assert torch.equal(ingest.resize_bicubic_u8(planar_in, oh, ow), exp)
    wide = torch.zeros(H, W + 5, 4, dtype=torch.uint8, device="cuda")
    wide[:, :W, :3] = img
    assert torch.equal(ingest.resize_bicubic_u8(wide[:, :W, :3], oh, ow), exp)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(480, 640, 3, 368, 490), (720, 1280, 3, 368, 654), (97, 131, 1, 200, 50),
                                   (64, 48, 4, 64, 96)])
def test_gpu_bicubic_vs_oracle_full_size(shape):
    from cbinfer_b200 import ingest
    H, W, Cc, oh, ow = shape
    rng = np.random.default_rng(H + W)
    img = rng.integers(0, 256, (H, W, Cc), dtype=np.uint8)
    img[H // 3:H // 2, W // 5:W // 2] = 255
    exp = np.concatenate([io.pil_bicubic_u8(img[..., c:c + 1], oh, ow) for c in range(Cc)], -1)
    got = ingest.resize_bicubic_u8(torch.from_numpy(img).cuda(), oh, ow).cpu().numpy()
    assert np.array_equal(got, exp)


@pytest.mark.gpu
def test_gpu_preprocess_pose_equals_reference_flow():
    """PoseDetector.preprocess (:66-72): both the fp32 tensor and the uint8-ingest route (first layer with
    inputNorm = (256, -0.5)) reproduce the reference's input tensor bit for bit."""
    import torch.nn as nn
    import cbinfer_b200 as cb
    from cbinfer_b200 import ingest
    img = torch.from_numpy(GOLD["in6"]).cuda()
    exp = torch.from_numpy(GOLD["pose_pre"]).cuda()
    got = ingest.preprocessPose(img, boxsize=92)
    assert got.shape == exp.shape and torch.equal(got, exp)
    u8 = ingest.preprocessPose(img, boxsize=92, asUint8=True)
    assert u8.dtype == torch.uint8 and u8.shape == (1,) + tuple(exp.shape)
    torch.manual_seed(0)
    conv = nn.Conv2d(3, 8, 3, padding=1).cuda()
    a, b = cb.CBConv2d(conv, 0.01), cb.CBConv2d(conv, 0.01)
    b.inputNorm = (256.0, -0.5)
    ya, yb = a(exp.unsqueeze(0)), b(u8)
    assert torch.equal(a.prevInput, b.prevInput) and torch.equal(ya, yb)


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_gpu_bilinear_vs_oracle(i):
    from cbinfer_b200 import ingest
    H, W, oh, ow = CASES[i]
    img = torch.from_numpy(GOLD["in%d" % i]).cuda()
    got = ingest.resize_bilinear_u8(img, oh, ow)
    exp = torch.from_numpy(GOLD["bilinear%d" % i]).permute(2, 0, 1).unsqueeze(0).cuda()
    assert got.shape == exp.shape and got.dtype == torch.float32
    assert float((got - exp).abs().max()) <= 1e-6          # float64 interpolation, fp32 result: <= 1 ulp of 1.0
    if (oh, ow) == (H, W):
        assert torch.equal(got, exp)


@pytest.mark.gpu
def test_gpu_read_scene_frame_shape_and_range():
    from cbinfer_b200 import ingest
    rng = np.random.default_rng(5)
    img = rng.integers(10, 200, (97, 130, 3), dtype=np.uint8)
    got = ingest.readSceneFrame(torch.from_numpy(img).cuda(), size=(776, 1040))
    assert got.shape == (1, 3, 776, 1040)
    exp = io.skimage_resize_bilinear(img.astype(np.float64) / 255, 776, 1040).astype(np.float32)
    assert float((got[0].permute(1, 2, 0).cpu() - torch.from_numpy(exp)).abs().max()) <= 1e-6
    assert float(got.min()) >= 10 / 255 - 1e-7 and float(got.max()) <= 199 / 255 + 1e-7    # clip=True
    with pytest.raises(Exception):
        ingest.readSceneFrame(torch.from_numpy(img))                                         # no CPU path
