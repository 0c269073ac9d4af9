"""Frame ingest on the device: what the reference's readers do to a decoded
uint8 frame on the host before the first layer sees it, as CUDA kernels behind the C ABI
(``cb_resize_bicubic_u8``, ``cb_resize_bilinear_u8``, csrc/ingest.cuh), so a decoder surface goes
resize -> first-layer change detection without visiting the host.

* :func:`preprocessPose` mirrors ``PoseDetector.preprocess`` (poseDetection/openPose/PoseDetector.py:45-73):
  scale the frame to ``boxsize`` rows with PIL's bicubic filter (bit-exact), then ``/255 * 255/256 - 0.5``.
* :func:`readSceneFrame` mirrors the per-frame part of ``getDataFrames``
  (sceneLabeling/videoSequenceReader.py:64-67): ``/255``, skimage's bilinear ``resize`` to 776 x 1040, NCHW.

There is no CPU path: tensors must live on the GPU.
"""
import torch

from . import _lib
from ._lib import C, check, require_cuda, stream_ptr

_plans = {}


def _strides_hwc(t):
    """(row, pixel, channel) strides of an [H, W, C] view, in elements."""
    return t.stride(0), t.stride(1), t.stride(2)


def resize_bicubic_u8(img, out_h, out_w, planar=False, out=None):
    """``PIL.Image.fromarray(img).resize((out_w, out_h), PIL.Image.BICUBIC)`` for a uint8 CUDA tensor
    ``[H, W, C]`` (any strides, C <= 4), bit-exact.  Returns uint8 ``[out_h, out_w, C]``, or ``[C, out_h,
    out_w]`` with ``planar=True`` (the layout ``ToTensor`` produces)."""
    require_cuda(img)
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] > 4:
        raise _lib.CBinferError("resize_bicubic_u8: uint8 [H, W, C <= 4] expected")
    sH, sW, Cc = img.shape
    key = (sH, sW, int(out_h), int(out_w), Cc, str(img.device))
    ws = _plans.get(key)
    if ws is None:
        ws = torch.empty(C.cb_resize_ws_bytes(sH, sW, out_h, out_w, Cc), dtype=torch.uint8, device=img.device)
        check(C.cb_resize_bicubic_u8_init(stream_ptr(img.device), ws.data_ptr(), sH, sW, out_h, out_w, Cc))
        if len(_plans) > 64:
            _plans.clear()
        _plans[key] = ws
    if out is None:
        out = torch.empty((Cc, out_h, out_w) if planar else (out_h, out_w, Cc), dtype=torch.uint8,
                          device=img.device)
    view = out.permute(1, 2, 0) if planar else out
    if tuple(view.shape) != (out_h, out_w, Cc) or out.dtype != torch.uint8:
        raise _lib.CBinferError("resize_bicubic_u8: bad output tensor")
    check(C.cb_resize_bicubic_u8(stream_ptr(img.device), img.data_ptr(), *_strides_hwc(img), view.data_ptr(),
                                 *_strides_hwc(view), ws.data_ptr(), sH, sW, out_h, out_w, Cc))
    return out


def resize_bilinear_u8(img, out_h, out_w, divisor=255.0, cval=0.0, clip=True, out=None):
    """``skimage.transform.resize(img / divisor, [out_h, out_w], mode='constant', cval=cval, clip=clip)``
    (order 1) of a uint8 CUDA tensor ``[H, W, C]``; returns fp32 ``[1, C, out_h, out_w]``."""
    require_cuda(img)
    if img.dtype != torch.uint8 or img.dim() != 3:
        raise _lib.CBinferError("resize_bilinear_u8: uint8 [H, W, C] expected")
    sH, sW, Cc = img.shape
    lo, hi = -3.0e38, 3.0e38
    if clip:                                    # skimage clips to the input's value range
        mn, mx = torch.aminmax(img)
        lo, hi = float(mn) / divisor, float(mx) / divisor
        lo, hi = min(lo, hi), max(lo, hi)
    if out is None:
        out = torch.empty(1, Cc, out_h, out_w, dtype=torch.float32, device=img.device)
    view = out[0].permute(1, 2, 0)
    check(C.cb_resize_bilinear_u8(stream_ptr(img.device), img.data_ptr(), *_strides_hwc(img), sH, sW,
                                  view.data_ptr(), *_strides_hwc(view), out_h, out_w, Cc, float(divisor),
                                  float(cval), lo, hi))
    return out


def preprocessPose(oriImg, boxsize=368, asUint8=False):
    """``PoseDetector.preprocess`` (openPose/PoseDetector.py:45-73) for a uint8 CUDA frame ``[H, W, 3]``:
    ``scale = boxsize / H``; bicubic resize to ``(int(H*scale), int(W*scale))``; ``ToTensor`` (/255);
    ``padBottomRight`` (returns its input, util.py:96); ``mul_(255/256).add_(-0.5)``.

    Returns the fp32 tensor ``[3, H', W']`` the reference feeds (bit-identical), or with ``asUint8=True`` the
    resized uint8 planes ``[1, 3, H', W']`` for a first layer with ``inputNorm = (256.0, -0.5)`` -- the
    normalisation then happens inside the change detection (cb_change_detect_u8), bit-identical too."""
    H, W = oriImg.shape[0], oriImg.shape[1]
    scale = boxsize / float(H)
    oh, ow = int(H * scale), int(W * scale)
    planes = resize_bicubic_u8(oriImg, oh, ow, planar=True)
    if asUint8:
        return planes.unsqueeze(0)
    # (u8 / 255) * (255 / 256) - 0.5 in fp32 equals u8 / 256 - 0.5 bit for bit for all 256 byte values
    # (tests/test_ingest.py::test_pose_normalisation_equals_divide_by_256); 1/256 is a power of two: exact
    return planes.float().mul_(1.0 / 256.0).add_(-0.5)


def readSceneFrame(img, size=(776, 1040)):
    """per-frame part of ``getDataFrames`` (sceneLabeling/videoSequenceReader.py:64-67) for a uint8 CUDA
    frame ``[H, W, 3]``: ``/255``, bilinear ``resize`` to ``size``, ``permute(2,0,1).unsqueeze(0).float()``."""
    return resize_bilinear_u8(img, size[0], size[1], divisor=255.0, cval=0.0, clip=True)
