#!/usr/bin/env python
"""The two ingest resizers on a 720p uint8 frame (one call each after warm-up): target of the ncu capture in
tools/gpu_profile_ingest.sh (profiles/r02_ncu_ingest.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cbinfer_b200 import ingest
img = torch.randint(0, 256, (720, 1280, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ingest.resize_bicubic_u8(img, 368, 654, planar=True)
    ingest.resize_bilinear_u8(img, 776, 1040, clip=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ingest.resize_bicubic_u8(img, 368, 654, planar=True)
ingest.resize_bilinear_u8(img, 776, 1040, clip=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("algorithmic bytes: bicubic horizontal %d + vertical %d, bilinear %d"
      % (720 * 1280 * 3 + 720 * 654 * 3, 720 * 654 * 3 + 368 * 654 * 3, 720 * 1280 * 3 + 776 * 1040 * 3 * 4))
