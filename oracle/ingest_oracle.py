"""ingest_oracle.py -- TEST INFRASTRUCTURE ONLY: CPU restatement of the image resizing the reference's
readers apply before the first layer.  Only ``tests/`` may import this.

Both algorithms live in third-party dependencies of the reference that are NOT part of its checkout
(``conda-env-cbinfer.yml``: ``pillow=5.0.0``, ``scikit-image=0.13.1``); they are restated here from
their published sources and anchored on the reference's call sites:

* ``pil_bicubic_u8``: ``torchvision.transforms.Scale(boxsize, interpolation=3)`` in
  ``poseDetection/openPose/PoseDetector.py:67`` = ``PIL.Image.resize(size, BICUBIC)`` on an 8-bit RGB
  image: Pillow ``src/libImaging/Resample.c`` -- ``precompute_coeffs`` (filter support scaled by the
  down-scaling factor, window bounds rounded with +0.5, weights normalised in double),
  ``normalize_coeffs_8bpc`` (fixed point, 22 fractional bits, round half away from zero),
  ``ImagingResampleHorizontal_8bpc`` then ``ImagingResampleVertical_8bpc`` (accumulator starts at
  1 << 21, arithmetic shift, saturate to 0..255; the horizontally resized image is rounded to 8 bits before
  the vertical pass).  PINNED: bit-exact against the Pillow of this image (12.2; the resampling core is
  unchanged since 4.x) in ``tests/test_ingest.py`` and through ``tests/golden/ingest_golden.npz``.
* ``skimage_resize_bilinear``: ``skimage.transform.resize(img, [776, 1040], mode='constant')`` in
  ``sceneLabeling/videoSequenceReader.py:66`` (order 1, no anti-aliasing in 0.13.1, ``clip=True``):
  ``skimage/transform/_warps.py`` ``resize`` -> ``warp`` -> ``_warps_cy._warp_fast`` /
  ``interpolation.pxd::bilinear_interpolation``: source coordinate ``scale * (i + 0.5) - 0.5`` per axis,
  the four neighbours fetched with constant padding (``cval`` outside the image), result clipped to the
  input's value range.  scikit-image is not in this image: PARITY UNPINNED for this function (checked
  against ``scipy.ndimage.map_coordinates(order=1, mode='constant')``, the same interpolation rule).
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x, a=-0.5):
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size, out_size, support=2.0, filt=_bicubic):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc: (ksize, bounds[out,2], kk[out,ksize] int32)."""
    scale = float(in_size) / out_size                  # (in1 - in0) / outSize with the full box
    filterscale = max(scale, 1.0)
    sup = support * filterscale
    ksize = int(math.ceil(sup)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - sup + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + sup + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _resample_axis0(img, out_size):
    """one 8bpc pass along axis 0 of a [n, m, C] uint8 array."""
    n = img.shape[0]
    if out_size == n:
        return img
    ksize, bounds, kk = pil_coeffs(n, out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for yy in range(out_size):
        lo, cnt = bounds[yy]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        acc += np.tensordot(kk[yy, :cnt].astype(np.int64), src[lo:lo + cnt], axes=(0, 0))
        out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_bicubic_u8(img, out_h, out_w):
    """PIL.Image.fromarray(img).resize((out_w, out_h), BICUBIC) for an [H, W, C] uint8 array."""
    img = np.ascontiguousarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3
    tmp = _resample_axis0(img.transpose(1, 0, 2), out_w).transpose(1, 0, 2)     # horizontal pass first
    return np.ascontiguousarray(_resample_axis0(np.ascontiguousarray(tmp), out_h))


def skimage_resize_bilinear(img, out_h, out_w, cval=0.0, clip=True):
    """skimage.transform.resize(img, [out_h, out_w], order=1, mode='constant', cval=cval, clip=clip) of a
    float [H, W, C] array (float64 arithmetic, as skimage's img_as_float path)."""
    img = np.asarray(img, np.float64)
    H, W, C = img.shape
    rs, cs = float(H) / out_h, float(W) / out_w
    r = rs * (np.arange(out_h) + 0.5) - 0.5
    c = cs * (np.arange(out_w) + 0.5) - 0.5
    pad = np.full((H + 2, W + 2, C), cval, np.float64)         # constant padding, one pixel is enough:
    pad[1:-1, 1:-1] = img                                      # |coordinate| never leaves (-1, size)
    minr, minc = np.floor(r).astype(int), np.floor(c).astype(int)
    maxr, maxc = np.ceil(r).astype(int), np.ceil(c).astype(int)
    dr, dc = (r - minr)[:, None, None], (c - minc)[None, :, None]

    def get(rr, cc):
        return pad[np.clip(rr + 1, 0, H + 1)[:, None], np.clip(cc + 1, 0, W + 1)[None, :]]
    top = (1 - dc) * get(minr, minc) + dc * get(minr, maxc)
    bottom = (1 - dc) * get(maxr, minc) + dc * get(maxr, maxc)
    out = (1 - dr) * top + dr * bottom
    if clip:
        out = np.clip(out, img.min(), img.max())
    return out
