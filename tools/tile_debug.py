#!/usr/bin/env python
"""Debug aid for the tiled contraction: one layer shape, all / block changed, error map vs dense
F.conv2d and vs the index-list kernel, grouped by tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
import cbinfer_b200 as cb
from cbinfer_b200 import conv2d_cg as cg, _lib

torch.backends.cudnn.allow_tf32 = False


def run(B, Cin, Cout, H, W, k, frac, mode="bf16x3", dt=torch.float32, reps=2):
    gemm = cb.CBConv2d.GEMM_MODES[mode]
    g = torch.Generator().manual_seed(1)
    state, sbuf = cg.pixel_major((B, Cin, H, W), dt, "cuda", 0)
    state.copy_((torch.rand(B, Cin, H, W, generator=g) - 0.5).to(dt))
    w = ((torch.rand(Cout, Cin, k, k, generator=g) - 0.5) * 2 * (Cin * k * k) ** -0.5).to(dt).cuda()
    bias = (torch.rand(Cout, generator=g) - 0.5).to(dt).cuda()
    raw = (torch.rand(B, H, W, generator=g) < frac).to(torch.int8).cuda()
    raw_bits, shape = cg._map_to_bits(raw)
    s = cg.alloc_scratch(shape, "cuda")
    tws = cg.alloc_tile_ws(shape, "cuda")
    packed = cg.pack_weights(w, gemm)
    ref = F.conv2d(state.float(), w.float(), bias.float(), padding=k // 2)
    scale = float(ref.abs().max())
    for rep in range(reps):
        cg.dilate_compact(raw_bits, shape, (k, k), s["idx"], s["count"], s["ws"], dil_bits=s["dil_bits"], tile_ws=tws)
        out, obuf = cg.pixel_major((B, Cout, H, W), dt, "cuda", 0)
        out.fill_(2.0)
        cg.conv_update_tiled(sbuf, tws, s["dil_bits"], packed, bias.float().contiguous(), obuf, Cin, Cout, (k, k), False, gemm)
        torch.cuda.synchronize()
        ntl = int(tws[1])
        NT = (tws.numel() - 4) // 2
        lst = tws[4 + NT: 4 + NT + ntl].cpu().numpy()
        o = out.float()
        err = (o - ref).abs().amax(dim=1) / scale              # [B,H,W]
        touched = torch.from_numpy(np.unpackbits(s["dil_bits"].cpu().numpy().view(np.uint8), bitorder="little")
                                   .reshape(B * H, -1)[:, :W].reshape(B, H, W).astype(bool)).cuda()
        e = torch.where(touched, err, torch.zeros_like(err))
        nan = int(torch.isnan(o).sum())
        bad = (e > 1e-4) | torch.isnan(e)
        print("B%d %d->%d %dx%d k%d frac %.2f %s rep %d: tiles %d  max err %.3e  bad px %d / %d  nan %d" % (
            B, Cin, Cout, H, W, k, frac, mode, rep, ntl, float(torch.nan_to_num(e, nan=9.0).max()), int(bad.sum()),
            int(touched.sum()), nan), flush=True)
        if int(bad.sum()):
            TY, TXp = (H + 15) // 16, ((W + 31) // 32) * 4
            pos = {int(t): i for i, t in enumerate(lst)}
            bb, yy, xx = torch.nonzero(bad, as_tuple=True)
            tiles = ((bb * TY + yy // 16) * TXp + xx // 8).cpu().numpy()
            ut, cnt = np.unique(tiles, return_counts=True)
            grid = min(148 * 2, ntl)
            its = sorted(set(pos[int(t)] // grid for t in ut))
            print("   bad tiles %d; pixels per bad tile min/max %d/%d; list positions // grid (=iteration): %s" % (
                len(ut), cnt.min(), cnt.max(), its[:20]))
            print("   first bad tiles (list pos, b, ty, tx, bad px):", [(pos[int(t)], int(t) // (TY * TXp), (int(t) % (TY * TXp)) // TXp, int(t) % TXp, int(c)) for t, c in list(zip(ut, cnt))[:8]])
            t0 = int(ut[0]); b0, ty0, tx0 = t0 // (TY * TXp), (t0 % (TY * TXp)) // TXp, t0 % TXp
            sub = e[b0, ty0 * 16:ty0 * 16 + 16, tx0 * 8:tx0 * 8 + 8]
            print("   error map of the first bad tile (x1e4):\n", (sub * 1e4).round().int().cpu().numpy())


if __name__ == "__main__":
    cfgs = [(1, 3, 16, 480, 640, 7, 1.0), (1, 16, 64, 240, 320, 7, 1.0), (1, 16, 64, 64, 64, 7, 1.0),
            (1, 3, 16, 64, 64, 7, 1.0), (1, 16, 64, 240, 320, 7, 0.002), (2, 64, 64, 120, 160, 3, 1.0)]
    for c in cfgs:
        run(*c)
