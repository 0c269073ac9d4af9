"""Synthetic static-camera video: frame t is frame t-1 with a clustered
change region re-drawn, so the per-frame change rate is controlled exactly.

    block mode: one axis-aligned rectangle of area r*H*W (aspect 4:3) at a position drawn from
                Generator(seed=1000+t)
    iid   mode: Bernoulli(r) per pixel (stress: a 7x7 dilation turns 5 % into ~92 % dirty)
"""
import math

import torch


def base_frame(B, H, W, seed=0, lo=0.0, hi=1.0, device="cpu", dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    f = torch.rand(B, 3, H, W, generator=g) * (hi - lo) + lo
    return f.to(device=device, dtype=dtype)


def next_frame(prev, t, rate, mode="block", lo=0.0, hi=1.0):
    """Return frame t given frame t-1 (same device/dtype); `rate` in [0,1]."""
    B, C, H, W = prev.shape
    f = prev.clone()
    if rate <= 0:
        return f
    g = torch.Generator(device="cpu").manual_seed(1000 + t)
    if mode == "block":
        area = rate * H * W
        bh = min(H, max(1, int(round(math.sqrt(area * 3.0 / 4.0)))))
        bw = min(W, max(1, int(round(area / bh))))
        for b in range(B):
            y0 = int(torch.randint(0, H - bh + 1, (1,), generator=g))
            x0 = int(torch.randint(0, W - bw + 1, (1,), generator=g))
            patch = torch.rand(C, bh, bw, generator=g) * (hi - lo) + lo
            f[b, :, y0:y0 + bh, x0:x0 + bw] = patch.to(device=f.device, dtype=f.dtype)
    elif mode == "iid":
        m = (torch.rand(B, 1, H, W, generator=g) < rate).to(f.device)
        noise = (torch.rand(B, C, H, W, generator=g) * (hi - lo) + lo).to(device=f.device, dtype=f.dtype)
        f = torch.where(m, noise, f)
    else:
        raise ValueError(mode)
    return f


def sequence(B, H, W, n, rate, mode="block", seed=0, lo=0.0, hi=1.0, device="cpu",
             dtype=torch.float32):
    """List of n frames [B,3,H,W]."""
    frames = [base_frame(B, H, W, seed, lo, hi, device, dtype)]
    for t in range(1, n):
        frames.append(next_frame(frames[-1], t, rate, mode, lo, hi))
    return frames
