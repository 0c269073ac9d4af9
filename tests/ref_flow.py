"""GPU-side reference-flow oracle (test infrastructure, never imported by the product).

A torch-2 restatement of the reference's per-frame module flows -- ``CBConv2d.forward_normal``
(pycbinfer/conv2d.py:178-259) and ``CBPoolMax2d.forward`` (pycbinfer/conv2d.py:49-78) -- that
drives the UNMODIFIED reference CUDA kernels (``oracle/_ref/*.so``, built from
pycbinfer/cbconv2d_cg_backend.cu and cbconv2d_cg_half_backend.cu where they lie, compute_52/61 PTX
JIT-compiled on the box) with the launch geometry of the reference's own wrappers
(pycbinfer/conv2d_cg.py:58-82,100-122,239-261,292-313) and ``torch.matmul`` for the GEMM
(conv2d_cg.py:342-349; TF32 must be off).  The reference's python package itself cannot be imported
under torch 2.x; what differs here is only the torch idiom (no Variable, no
legacy tensor constructors), not a single kernel or launch parameter.

Planar NCHW, batch 1 -- one instance per video stream, exactly like the reference.
"""
import ctypes

import torch

from tests.util import ref_lib, vp


def libs():
    """(fp32 library, fp16 library) or None when oracle/_ref was not built."""
    a, b = ref_lib("cbconv2d_cg_backend"), ref_lib("cbconv2d_cg_half_backend")
    if a is None or b is None:
        return None
    return a, b


class RefCBConv2d(object):
    """conv2d.py:87-259 (coarse-grained path)."""

    def __init__(self, weight, bias, threshold, withReLU=False, feedbackLoop=False,
                 propChangeIndexes=False):
        self.weight = weight.detach().contiguous()
        self.bias = bias.detach().contiguous()
        self.threshold = float(threshold)
        self.kernel_size = tuple(weight.shape[2:])
        self.in_channels, self.out_channels = weight.shape[1], weight.shape[0]
        self.withReLU, self.feedbackLoop, self.propChangeIndexes = withReLU, feedbackLoop, propChangeIndexes
        self.prevInput = weight.new_empty(0)
        self.prevOutput = weight.new_empty(0)
        self.changeMap = None
        self.changeIndexes = None
        self.half = weight.dtype == torch.float16
        self.lib = libs()[1 if self.half else 0]

    def forward(self, inp):
        changeIndexes = None
        if type(inp) == tuple:                                            # conv2d.py:180-187
            assert inp[0] == 'changeIndexes'
            input, changeIndexes = inp[1].contiguous(), inp[2].contiguous()
        else:
            input = inp.contiguous()
        assert input.dim() == 4 and input.size(0) == 1 and input.size(1) == self.in_channels
        _, C, H, W = input.shape
        if self.prevInput.size() != input.size():                         # :192-194
            self.prevInput = torch.full_like(input, float("inf"))
        outpSize = (1, self.out_channels, H, W)
        if tuple(self.prevOutput.size()) != outpSize:                     # :195-199
            self.prevOutput = input.new_full(outpSize, float("inf"))
        kH, kW = self.kernel_size
        if changeIndexes is None:
            # changeDetection (conv2d_cg.py:100-122): zeroed CharTensor map, block 128
            changeMap = torch.zeros(H, W, dtype=torch.int8, device=input.device)
            grid = (H * W - 1) // 128 + 1
            self.lib.changeDetection(1, 1, grid, 1, 1, 128, vp(input), vp(self.prevInput), vp(changeMap),
                                     W, H, C, int((kH - 1) / 2), int((kW - 1) / 2),
                                     ctypes.c_float(self.threshold), ctypes.c_bool(self.feedbackLoop))
            self.changeMap = changeMap
            changeIndexes = torch.nonzero(changeMap.view(-1)).int().view(-1)   # conv2d_cg.py:200-203
        if not self.feedbackLoop:                                         # :234-238
            self.prevInput.copy_(input)
        n = changeIndexes.numel()
        if n != 0:                                                        # :240-253
            # genXMatrix (conv2d_cg.py:239-261): block (kW, 128//(kH*kW), kH)
            threadZ = 128 // (kH * kW)
            X = self.prevInput.new_empty(n, C * kH * kW)
            self.lib.genXMatrix(1, 1, (n - 1) // threadZ + 1, kH, threadZ, kW, vp(X), vp(self.prevInput),
                                vp(changeIndexes), kW, kH, C, W, H, n)
            # matrixMult_python (conv2d_cg.py:342-349)
            Y = X.matmul(self.weight.view(self.out_channels, -1).transpose(0, 1)).add_(self.bias)
            Yt = Y.transpose(0, 1).contiguous()                           # conv2d.py:247, conv2d_cg.py:305
            # updateOutput (conv2d_cg.py:292-313): block 1024
            self.lib.updateOutput(1, 1, (n * self.out_channels - 1) // 1024 + 1, 1, 1, 1024, vp(Yt),
                                  vp(self.prevOutput), vp(changeIndexes), H * W, n, self.out_channels,
                                  ctypes.c_bool(self.withReLU))
        self.changeIndexes = changeIndexes
        if self.propChangeIndexes:
            return 'changeIndexes', self.prevOutput, changeIndexes
        return self.prevOutput


class RefCBPoolMax2d(object):
    """conv2d.py:24-84 (2x2 / stride 2, ceil_mode honoured for the size only)."""

    def __init__(self, ceil_mode=False, propChangeIndexes=False):
        self.ceil_mode, self.propChangeIndexes = ceil_mode, propChangeIndexes
        self.outputState = None

    def forward(self, inp):
        assert type(inp) == tuple and inp[0] == 'changeIndexes'
        input, changeIndexes = inp[1].contiguous(), inp[2].contiguous()
        half = input.dtype == torch.float16
        lib = libs()[1 if half else 0]
        n = changeIndexes.numel()
        _, nc, h, w = input.shape
        oh, ow = ((h - 1) // 2 + 1, (w - 1) // 2 + 1) if self.ceil_mode else (h // 2, w // 2)
        if self.outputState is None or tuple(self.outputState.shape[-3:]) != (nc, oh, ow):
            # (the reference allocates lazily on the first non-empty list, conv2d.py:53-62; the
            #  first frame always has one)
            self.outputState = input.new_full((1, nc, oh, ow), float("inf"))
        if n != 0:
            assert h % 2 == 0 and w % 2 == 0, "the reference kernel has no bounds guard for odd sizes"
            lib.maxPool2d((n - 1) // 64 + 1, 64, vp(input), vp(self.outputState), vp(changeIndexes), n, nc,
                          h, w, oh, ow, 2, 2)
        output = self.outputState.clone()                                 # conv2d.py:73
        if self.propChangeIndexes:
            return 'changeIndexes', output, changeIndexes
        return output


def convert_sequential(base, thresholds, feedbackLoop=True, pools=True):
    """The reference's conversion recipe for a plain nn.Sequential (pycbinfer/__init__.py:10-66,
    sceneLabeling/modelLoader.py:41-87 with experimentIdx 5/6): Conv2d -> RefCBConv2d with the
    following ReLU merged, MaxPool2d -> RefCBPoolMax2d fed by the preceding conv's indices.
    Returns the list of layers; run them with :func:`run`."""
    import torch.nn as nn
    kids = list(base.children())
    layers, ci = [], 0
    i = 0
    while i < len(kids):
        k = kids[i]
        if isinstance(k, nn.Conv2d):
            relu = i + 1 < len(kids) and isinstance(kids[i + 1], nn.ReLU)
            nxt = kids[i + (2 if relu else 1)] if i + (2 if relu else 1) < len(kids) else None
            layers.append(RefCBConv2d(k.weight, k.bias, thresholds[ci], withReLU=relu,
                                      feedbackLoop=feedbackLoop,
                                      propChangeIndexes=pools and isinstance(nxt, nn.MaxPool2d)))
            ci += 1
            i += 2 if relu else 1
        elif isinstance(k, nn.MaxPool2d) and pools:
            layers.append(RefCBPoolMax2d(ceil_mode=k.ceil_mode))
            i += 1
        elif isinstance(k, (nn.Dropout, nn.Dropout2d)):
            i += 1
        else:
            layers.append(k)
            i += 1
    return layers


def run(layers, x):
    """one frame through the layer list; returns (final output, per-layer outputs)."""
    outs = []
    for l in layers:
        x = l.forward(x) if hasattr(l, "forward") and not isinstance(l, torch.nn.Module) else l(x)
        outs.append(x)
    return (x[1] if type(x) == tuple else x), outs
