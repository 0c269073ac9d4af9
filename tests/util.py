"""helpers shared by the GPU parity tests: torch <-> oracle (numpy) plumbing."""
import ctypes
import os
import platform

import numpy as np
import torch

from oracle import oracle as orc

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TORCH_DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
ORC_DT = {"f32": orc.F32, "f16": orc.F16, "bf16": orc.BF16}


def to_np(t):
    """torch tensor (any device / layout) -> oracle storage array in planar NCHW order."""
    return orc.from_torch(t.detach().cpu().contiguous())[0]


def to_val(t):
    return t.detach().cpu().contiguous().double().numpy()


def bits_to_map(bits, B, H, W):
    """int32 bitmap tensor -> uint8 [B,H,W] numpy map."""
    Wd = (W + 31) // 32
    words = bits.detach().cpu().numpy().view(np.uint32)[: B * H * Wd].reshape(B * H, Wd)
    m = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).astype(np.uint8)
    return m.reshape(B * H, Wd * 32)[:, :W].reshape(B, H, W)


def rand_tensor(shape, dt, seed, scale=1.0, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(TORCH_DT[dt]).to(device)


def perturb(x, frac, seed, scale=1.0):
    """copy of x with ~frac of the *pixels* changed in a random subset of channels."""
    g = torch.Generator().manual_seed(seed)
    B, C, H, W = x.shape
    pm = (torch.rand(B, 1, H, W, generator=g) < frac)
    cm = (torch.rand(B, C, H, W, generator=g) < 0.5)
    d = torch.randn(B, C, H, W, generator=g) * scale
    y = x.detach().cpu().float() + (pm & cm).float() * d
    return y.to(x.dtype).to(x.device)


def ref_lib(name):
    """ctypes handle of an UNMODIFIED reference library compiled into oracle/_ref (or None)."""
    p = os.path.join(REPO, "oracle", "_ref", "%s_%s.so" % (name, platform.machine()))
    if not os.path.exists(p):
        return None
    return ctypes.CDLL(p)


def vp(t):
    return ctypes.c_void_p(t.data_ptr())
