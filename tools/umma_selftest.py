#!/usr/bin/env python
"""Diagnostics for the tcgen05 conv kernel (run on a GPU box): compares the tensor-core modes with
the SIMT fp32 kernel on a few shapes and, when they disagree, probes with one-hot weights to show
which K index each UMMA K slot really reads (swizzle / descriptor bugs show up as a permutation)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import cbinfer_b200 as cb
from cbinfer_b200 import _lib, conv2d_cg as cg


def run(mode, dt, B, Cin, Cout, H, W, k, w=None, frac=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    state, sbuf = cg.pixel_major((B, Cin, H, W), dt, "cuda", 0)
    state.copy_(torch.randn(B, Cin, H, W, generator=g).to(dt).cuda())
    out, obuf = cg.pixel_major((B, Cout, H, W), dt, "cuda", 0)
    if w is None:
        w = (torch.randn(Cout, Cin, k, k, generator=g) * (Cin * k * k) ** -0.5).to(dt).cuda()
    bias = torch.zeros(Cout, device="cuda")
    sel = torch.nonzero(torch.rand(B * H * W, generator=g) < frac).view(-1).int().cuda()
    ci = cg.ChangeIndexes.from_tensor(sel, (B, H, W))
    gemm = cb.CBConv2d.GEMM_MODES[mode]
    cg.conv_update(sbuf, ci, cg.pack_weights(w, gemm), bias, obuf, Cin, Cout, (k, k), False, gemm)
    torch.cuda.synchronize()
    return out.float().clone(), state.float().clone(), sel


def main():
    torch.manual_seed(0)
    shapes = [(1, 16, 16, 8, 16, 1), (1, 32, 64, 8, 16, 1), (1, 16, 64, 12, 20, 3), (1, 3, 16, 20, 30, 7),
              (2, 64, 256, 9, 12, 7), (1, 256, 64, 10, 16, 1)]
    bad = False
    for dt, modes in ((torch.float32, ("tc", "tc3x", "bf16x3")), (torch.bfloat16, ("tc",)), (torch.float16, ("tc",))):
        for (B, Cin, Cout, H, W, k) in shapes:
            ref, _, _ = run("simt", dt, B, Cin, Cout, H, W, k)
            for mode in modes:
                try:
                    got, _, _ = run(mode, dt, B, Cin, Cout, H, W, k)
                    err = float((got - ref).abs().max() / (ref.abs().max() + 1e-30))
                except Exception as e:
                    err = float("nan")
                    print("EXC", mode, dt, e)
                flag = "" if err < 2e-2 else "   <-- MISMATCH"
                bad |= not (err < 2e-2)
                print("%-5s %-14s B%d Cin%-3d Cout%-3d %2dx%-2d k%d  rel err vs simt %.3e%s"
                      % (mode, str(dt), B, Cin, Cout, H, W, k, err, flag), flush=True)
    if bad:
        # one-hot probe, fp32 single pass, 1x1 conv, Cin = 32 (one 128-byte K row)
        Cin, Cout, H, W = 32, 16, 8, 16
        for kprobe in range(0, Cin):
            w = torch.zeros(Cout, Cin, 1, 1, device="cuda")
            w[0, kprobe] = 1.0
            got, state, sel = run("tc", torch.float32, 1, Cin, Cout, H, W, 1, w=w)
            col = got[0, 0].reshape(-1)            # should equal state[0, kprobe]
            src = state[0].reshape(Cin, -1)
            match = [(c, float((src[c] - col).abs().max())) for c in range(Cin)]
            best = min(match, key=lambda t: t[1])
            print("one-hot k=%2d -> output matches input channel %2d (err %.2e)" % (kprobe, best[0], best[1]))
    # how does kind::tf32 convert fp32 operands?  v = 1 + 2^-11 + 2^-12: truncation -> 1.0,
    # round-to-nearest -> 1 + 2^-10.  The 3xTF32 path assumes truncation (hi = raw fp32 state).
    Cin, Cout, H, W = 32, 16, 8, 16
    w = torch.zeros(Cout, Cin, 1, 1, device="cuda")
    w[0, 0] = 1.0
    state, sbuf = cg.pixel_major((1, Cin, H, W), torch.float32, "cuda", 0)
    v = 1.0 + 2.0 ** -11 + 2.0 ** -12
    state.fill_(v)
    out, obuf = cg.pixel_major((1, Cout, H, W), torch.float32, "cuda", 0)
    sel = torch.arange(H * W, dtype=torch.int32, device="cuda")
    ci = cg.ChangeIndexes.from_tensor(sel, (1, H, W))
    cg.conv_update(sbuf, ci, cg.pack_weights(w, _lib.GEMM_TC), torch.zeros(Cout, device="cuda"), obuf,
                   Cin, Cout, (1, 1), False, _lib.GEMM_TC)
    torch.cuda.synchronize()
    r = float(out[0, 0, 0, 0])
    print("tf32 operand conversion probe: in %.10f -> out %.10f  (%s)" %
          (v, r, "TRUNCATION" if r == 1.0 else "ROUNDING" if r == 1.0 + 2.0 ** -10 else "??"))
    bad |= r != 1.0
    print("SELFTEST", "FAILED" if bad else "PASSED")


if __name__ == "__main__":
    main()
