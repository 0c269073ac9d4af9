#!/usr/bin/env python
"""Golden vectors for the frame-ingest resizers (tests/golden/ingest_golden.npz), produced in the build
container with the third-party code the reference calls: PIL.Image.resize(size, BICUBIC) (Pillow 12.2 here;
the reference pins pillow=5.0.0, same resampling core) through the very torchvision-style call of
PoseDetector.preprocess (poseDetection/openPose/PoseDetector.py:66-72).  scikit-image is not available in
this image, so the bilinear vectors come from the restatement in oracle/ingest_oracle.py (parity unpinned
for that function, see its header).

    python tests/golden/make_ingest_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import ingest_oracle as io   # noqa: E402

CASES = [(48, 64, 37, 49), (37, 53, 64, 80), (100, 60, 33, 21), (16, 16, 16, 40), (5, 7, 11, 3), (64, 64, 8, 8),
         (120, 160, 92, 122)]


def main():
    from PIL import Image
    rng = np.random.default_rng(2024)
    out = {}
    for i, (H, W, oh, ow) in enumerate(CASES):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        img[H // 4:H // 2, W // 4:W // 2] = 255           # saturated regions: overshoot must clip
        img[H // 2:, :W // 8] = 0
        out["in%d" % i] = img
        out["bicubic%d" % i] = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
        out["bilinear%d" % i] = io.skimage_resize_bilinear(img.astype(np.float64) / 255, oh, ow).astype(np.float32)
    # the whole preprocess of PoseDetector.py:66-72 on one frame (ToTensor = /255 in fp32)
    img = out["in6"]
    scale = 92 / float(img.shape[0])
    box = (int(img.shape[0] * scale), int(img.shape[1] * scale))
    pil = np.asarray(Image.fromarray(img).resize(box[::-1], Image.BICUBIC))
    t = pil.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    out["pose_pre"] = t * np.float32(255.0 / 256.0) + np.float32(-0.5)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ingest_golden.npz"), **out)
    print("wrote ingest_golden.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
