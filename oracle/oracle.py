"""oracle.py -- TEST INFRASTRUCTURE ONLY: python face of the CPU parity checker.

Loads ``oracle/_build/libcbinfer_oracle.so`` (the plain-C restatement in ``cbinfer_oracle.c``)
and restates the *host-side flow* of the reference modules on numpy arrays in the reference's
layout (planar NCHW, batch 1).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product package
``cbinfer_b200`` never does.

Reference anchors (relative to the reference checkout):
  * op semantics ......... pycbinfer/cbconv2d_cg_backend.cu, cbconv2d_cg_half_backend.cu,
                           cbconv2d_fg_backend.cu (cited per function in cbinfer_oracle.c)
  * CBConv2d flow ........ pycbinfer/conv2d.py:178-259 (forward_normal), :160-176 (forward_fg)
  * CBPoolMax2d flow ..... pycbinfer/conv2d.py:49-78

Parity pinning: tests/test_oracle.py checks this oracle against tests/golden/*.npz, which were
produced by the reference's own python twins / native conv2d_fg_cpu (tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcbinfer_oracle.so")

F32, F16, BF16 = 0, 1, 2


def build(force=False):
    """Compile the C restatement (gcc only).  Building the checker is not using it."""
    src = os.path.join(_HERE, "cbinfer_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "_build/libcbinfer_oracle.so"])
    return _LIB_PATH


def build_ref(reference="/root/reference"):
    """Compile the unmodified reference .cu files into oracle/_ref (only where the reference
    checkout exists, i.e. in the build container)."""
    if os.path.isdir(os.path.join(reference, "pycbinfer")):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref", "REFERENCE=" + reference])
        return True
    return False


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i32, f32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_long
        L.orc_change_detection.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, i32, i32]
        L.orc_change_propagation.argtypes = [vp, vp, i32, i32, i32, i32]
        L.orc_change_indexes.argtypes = [vp, i64, vp]
        L.orc_change_indexes.restype = i32
        L.orc_gen_xmatrix.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32]
        L.orc_matrix_mult.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32]
        L.orc_update_output.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32]
        L.orc_maxpool2d.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32]
        L.orc_fg_detect.argtypes = [vp, vp, vp, vp, i64, f32]
        L.orc_fg_update.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i64]
        L.orc_cbconv_frame.argtypes = [vp] * 7 + [i32] * 6 + [f32, i32, i32, i32]
        L.orc_cbconv_frame.restype = i32
        L.orc_cbconv_frame_fast_f32.argtypes = [vp] * 7 + [i32] * 6 + [f32, i32, i32]
        L.orc_cbconv_frame_fast_f32.restype = i32
        L.orc_f64_to_f16.argtypes = [ctypes.c_double]
        L.orc_f64_to_f16.restype = ctypes.c_uint16
        L.orc_f64_to_bf16.argtypes = [ctypes.c_double]
        L.orc_f64_to_bf16.restype = ctypes.c_uint16
        _lib = L
    return _lib


# ------------------------------------------------------------------------------------------
# dtype plumbing: fp32 arrays are np.float32; fp16 / bf16 arrays are np.uint16 bit patterns
# ------------------------------------------------------------------------------------------

def _p(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def _np_dtype(dtype):
    return np.float32 if dtype == F32 else np.uint16


def to_bits(x64, dtype):
    """float64 array -> storage array of the given dtype code (RNE)."""
    x64 = np.ascontiguousarray(x64, dtype=np.float64)
    if dtype == F32:
        return x64.astype(np.float32)
    if dtype == F16:
        return x64.astype(np.float16).view(np.uint16)
    # bf16 RNE from float64 through the C helper semantics, vectorised: round float32 bits
    f = x64.astype(np.float32)
    # double->float->bf16 double rounding is avoided by only using this on values that are
    # exactly representable in float32 (test data is generated in float32).
    b = f.view(np.uint32).astype(np.uint64)
    nan = np.isnan(f)
    r = ((b + 0x7FFF + ((b >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = 0x7FC0
    return r


def from_bits(a, dtype):
    """storage array -> float64 values."""
    if dtype == F32:
        return a.astype(np.float64)
    if dtype == F16:
        return a.view(np.float16).astype(np.float64)
    return (a.astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def from_torch(t):
    """torch CPU tensor -> (storage ndarray, dtype code)."""
    import torch
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.float32:
        return t.numpy().copy(), F32
    if t.dtype == torch.float16:
        return t.view(torch.int16).numpy().view(np.uint16).copy(), F16
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16).copy(), BF16
    raise TypeError(t.dtype)


def to_torch(a, dtype):
    import torch
    if dtype == F32:
        return torch.from_numpy(np.ascontiguousarray(a))
    td = torch.float16 if dtype == F16 else torch.bfloat16
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int16)).view(td)


def inf_like(shape, dtype):
    """state initial value: +inf in every cell (conv2d.py:193,199 fill_(1e1000))."""
    if dtype == F32:
        return np.full(shape, np.inf, dtype=np.float32)
    return np.full(shape, 0x7C00 if dtype == F16 else 0x7F80, dtype=np.uint16)


# ------------------------------------------------------------------------------------------
# ops, one per reference wrapper in pycbinfer/conv2d_cg.py / conv2d_fg.py
# ------------------------------------------------------------------------------------------

def changeDetection(inp, prevInput, filtSize, threshold, updateInputState=False, dtype=F32,
                    return_raw=False):
    """conv2d_cg.py:100-122 -> cbconv2d_cg_backend.cu:40-100.  inp/prevInput: [1,C,H,W]
    (prevInput is updated in place when updateInputState).  Returns int8 map [H,W]."""
    _, C, H, W = inp.shape
    cmap = np.zeros((H, W), dtype=np.uint8)
    raw = np.zeros((H, W), dtype=np.uint8)
    lib().orc_change_detection(_p(inp), _p(prevInput), _p(cmap), _p(raw), W, H, C,
                               (filtSize[0] - 1) // 2, (filtSize[1] - 1) // 2,
                               float(threshold), int(updateInputState), dtype)
    return (cmap, raw) if return_raw else cmap


def changePropagation(changeMap, filtSize):
    """conv2d_cg.py:159-177 -> cbconv2d_cg_backend.cu:101-136."""
    H, W = changeMap.shape[-2:]
    src = np.ascontiguousarray(changeMap.reshape(H, W).astype(np.uint8))
    out = np.zeros_like(src)
    lib().orc_change_propagation(_p(src), _p(out), W, H, (filtSize[0] - 1) // 2,
                                 (filtSize[1] - 1) // 2)
    return out


def changeIndexesExtr(changeMap):
    """conv2d_cg.py:200-213: ascending int32 indices of non-zero cells."""
    flat = np.ascontiguousarray(changeMap.reshape(-1).astype(np.uint8))
    idx = np.zeros(flat.size, dtype=np.int32)
    n = lib().orc_change_indexes(_p(flat), flat.size, _p(idx))
    return idx[:n].copy()


def genXMatrix(inp, changeIndexes, filtSize, dtype=F32):
    """conv2d_cg.py:239-261 -> cbconv2d_cg_backend.cu:138-173."""
    _, C, H, W = inp.shape
    kH, kW = filtSize
    n = int(changeIndexes.size)
    X = np.zeros((n, C * kH * kW), dtype=_np_dtype(dtype))
    if n:
        lib().orc_gen_xmatrix(_p(X), _p(inp), _p(np.ascontiguousarray(changeIndexes, np.int32)),
                              kW, kH, C, W, H, n, dtype)
    return X


def matrixMult(X, weights, bias, dtype=F32):
    """conv2d_cg.py:342-349 (double accumulation, rounded once)."""
    n, K = X.shape
    Cout = weights.shape[0]
    Y = np.zeros((n, Cout), dtype=_np_dtype(dtype))
    if n:
        lib().orc_matrix_mult(_p(X), _p(np.ascontiguousarray(weights).reshape(Cout, K)), _p(bias),
                              _p(Y), n, K, Cout, dtype)
    return Y


def updateOutput(Yt, changeIndexes, prevOutput, withReLU=False, dtype=F32):
    """conv2d_cg.py:292-313 -> cbconv2d_cg_backend.cu:175-197.  Yt is [Cout, n]."""
    Cout, H, W = prevOutput.shape[-3:]
    n = int(changeIndexes.size)
    if n:
        lib().orc_update_output(_p(np.ascontiguousarray(Yt)), _p(prevOutput),
                                _p(np.ascontiguousarray(changeIndexes, np.int32)), H * W, n, Cout,
                                int(withReLU), dtype)
    return prevOutput


def maxPool2d(inp, outputState, changeIndexes, kernelSize=(2, 2), stride=(2, 2), dtype=F32):
    """conv2d_cg.py:58-82 -> cbconv2d_cg_backend.cu:199-240."""
    C, H, W = inp.shape[-3:]
    oH, oW = outputState.shape[-2:]
    n = int(changeIndexes.size)
    if n:
        lib().orc_maxpool2d(_p(inp), _p(outputState),
                            _p(np.ascontiguousarray(changeIndexes, np.int32)), n, C, H, W, oH, oW,
                            stride[0], stride[1], dtype)
    return outputState


def changeDetectionFG(inp, prevInput, threshold):
    """conv2d_fg.py:34-46 -> cbconv2d_fg_backend.cu:7-35 (fp32 only)."""
    diffs = np.zeros_like(inp)
    cmap = np.zeros(inp.shape, dtype=np.int8)
    lib().orc_fg_detect(_p(inp), _p(prevInput), _p(diffs), _p(cmap), inp.size, float(threshold))
    return diffs, cmap


def updateOutputFG(diffs, weight, output, changeCoords):
    """conv2d_fg.py:48-72 -> cbconv2d_fg_backend.cu:37-79 (double accumulation)."""
    Cout, Cin, kH, kW = weight.shape
    H, W = output.shape[-2:]
    coords = np.ascontiguousarray(changeCoords.reshape(-1), dtype=np.int64)
    lib().orc_fg_update(_p(diffs), _p(np.ascontiguousarray(weight)), _p(output), _p(coords), Cout,
                        Cin, H, W, kH, kW, coords.size)
    return output


def cbconvFG(inp, prevInput, output, weight, threshold):
    """conv2d_fg.py:75-96, GPU branch semantics (strict '>')."""
    diffs, cmap = changeDetectionFG(inp, prevInput, threshold)
    coords = np.flatnonzero(cmap.reshape(-1))
    if coords.size:
        updateOutputFG(diffs, weight, output, coords)
    return output


# ------------------------------------------------------------------------------------------
# module flow restatements
# ------------------------------------------------------------------------------------------

class OracleCBConv2d:
    """conv2d.py:87-304, coarse-grained path (forward_normal :178-259) and FG (:160-176)."""

    def __init__(self, weight, bias, threshold, dtype=F32, withReLU=False, feedbackLoop=False,
                 propChangeIndexes=False, finegrained=False):
        self.weight = np.ascontiguousarray(weight)          # [Cout,Cin,kH,kW] storage dtype
        self.bias = np.ascontiguousarray(bias)
        self.threshold = threshold
        self.dtype = dtype
        self.withReLU = withReLU
        self.feedbackLoop = feedbackLoop
        self.propChangeIndexes = propChangeIndexes
        self.finegrained = finegrained
        self.kernel_size = tuple(weight.shape[2:])
        self.out_channels, self.in_channels = weight.shape[:2]
        self.clearMemory()

    def clearMemory(self):
        self.prevInput = None
        self.prevOutput = None
        self.changeMap = None
        self.changeIndexes = None

    def forward(self, inp):
        if self.finegrained:
            return self._forward_fg(inp)
        changeIndexes = None
        if isinstance(inp, tuple):
            assert inp[0] == "changeIndexes"
            x, changeIndexes = inp[1], inp[2]
        else:
            x = inp
        x = np.ascontiguousarray(x)
        _, C, H, W = x.shape
        assert C == self.in_channels
        if self.prevInput is None or self.prevInput.shape != x.shape:
            self.prevInput = inf_like(x.shape, self.dtype)                    # :192-194
        oshape = (1, self.out_channels, H, W)
        if self.prevOutput is None or self.prevOutput.shape != oshape:
            self.prevOutput = inf_like(oshape, self.dtype)                    # :195-199
        if changeIndexes is None:
            self.changeMap = changeDetection(x, self.prevInput, self.kernel_size, self.threshold,
                                             updateInputState=self.feedbackLoop,
                                             dtype=self.dtype)                 # :222-224
            changeIndexes = changeIndexesExtr(self.changeMap)                 # :232
        if not self.feedbackLoop:
            self.prevInput = x.copy()                                         # :234-238
        self.changeIndexes = changeIndexes
        if changeIndexes.size:
            X = genXMatrix(self.prevInput, changeIndexes, self.kernel_size, self.dtype)   # :242
            Y = matrixMult(X, self.weight, self.bias, self.dtype)             # :246
            Yt = np.ascontiguousarray(Y.T)                                    # :247
            updateOutput(Yt, changeIndexes, self.prevOutput, self.withReLU, self.dtype)  # :249
        if self.propChangeIndexes:
            return "changeIndexes", self.prevOutput, changeIndexes            # :256-257
        return self.prevOutput                                                # :259

    def _forward_fg(self, x):
        assert self.dtype == F32 and not self.feedbackLoop
        x = np.ascontiguousarray(x)
        if self.prevInput is None or self.prevInput.shape != x.shape:
            self.prevOutput = dense_conv2d(x, self.weight, self.bias)         # :163-167
        else:
            po = self.prevOutput.copy()                                       # :169
            self.prevOutput = cbconvFG(x, self.prevInput, po, self.weight, self.threshold)
        out = self.prevOutput
        if self.withReLU:
            out = np.maximum(out, 0)                                          # :173-174
        self.prevInput = x.copy()                                             # :175
        return out


class OracleCBPoolMax2d:
    """conv2d.py:24-84."""

    def __init__(self, dtype=F32, ceil_mode=False):
        self.dtype = dtype
        self.ceil_mode = ceil_mode
        self.outputState = None

    def clearMemory(self):
        self.outputState = None

    def forward(self, inp):
        assert isinstance(inp, tuple) and inp[0] == "changeIndexes"          # :50
        x, idx = np.ascontiguousarray(inp[1]), inp[2]
        if idx.size:
            _, C, H, W = x.shape
            if self.ceil_mode:
                oh, ow = (H - 1) // 2 + 1, (W - 1) // 2 + 1                   # :57-58
            else:
                oh, ow = H // 2, W // 2                                       # :59-60
            if self.outputState is None or self.outputState.shape != (1, C, oh, ow):
                self.outputState = inf_like((1, C, oh, ow), self.dtype)       # :61-62
            maxPool2d(x, self.outputState, idx, dtype=self.dtype)             # :65
        return self.outputState.copy()                                        # :73


def dense_conv2d(x, weight, bias, relu=False):
    """fp64-accumulated dense 'same' convolution of an fp32 NCHW batch-1 array: the arbiter for
    'threshold 0 == dense nn.Conv2d' (cuDNN's own summation order is unpinned)."""
    _, C, H, W = x.shape
    Cout, Cin, kH, kW = weight.shape
    assert C == Cin
    xp = np.zeros((C, H + kH - 1, W + kW - 1), dtype=np.float64)
    xp[:, (kH - 1) // 2:(kH - 1) // 2 + H, (kW - 1) // 2:(kW - 1) // 2 + W] = x[0]
    out = np.zeros((Cout, H, W), dtype=np.float64)
    w64 = weight.astype(np.float64)
    for ky in range(kH):
        for kx in range(kW):
            out += np.einsum("oc,chw->ohw", w64[:, :, ky, kx], xp[:, ky:ky + H, kx:kx + W])
    out += bias.astype(np.float64)[:, None, None]
    if relu:
        out = np.maximum(out, 0)
    return out.astype(np.float32)[None]
