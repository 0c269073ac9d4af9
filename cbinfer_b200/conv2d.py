"""CBConv2d / CBPoolMax2d -- the nn.Modules of the pycbinfer surface on the sm_100a backend.

Mirrors the reference's ``pycbinfer/conv2d.py``: class ``CBPoolMax2d`` (:24-84) and class
``CBConv2d`` (:87-304) with the same constructor arguments, flags (``withReLU``,
``saveChangeMap``, ``propChangeIndexes``, ``gatherComputationStats``, ``finegrained``,
``copyInput``, ``feedbackLoop``; :112-118), ``threshold`` attribute, state buffers
(``prevInput``, ``prevOutput``, ``outputState``), ``clearMemory()`` / ``getStateTensors()``,
``('changeIndexes', tensor, indices)`` tuple protocol (:180-187, :256-257) and ``__repr__``.

What is different underneath (B200-first, not a translation):
  * state lives pixel-major ([B,H,W,pitch], exposed as channels-last [B,C,H,W] views) so a changed
    pixel's channels are one contiguous run for detection, gather, scatter and pooling;
  * one frame of a layer is three launches -- detect, dilate+compact, fused gather/tcgen05
    contraction/bias/ReLU/scatter -- with the change count kept on the device: no torch.nonzero,
    no host sync, no X / Y / Y^T matrices (reference: ~10 launches + 1 sync, conv2d.py:222-251);
  * batch > 1 = independent video streams (the reference is batch 1);
  * CUDA only: the reference's CPU branches (conv2d.py:225-227,244-245,252-253) do not exist.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import conv2d_cg as cg
from .conv2d_cg import ChangeIndexes
from .conv2d_fg import cbconvFG

_INF = float("inf")


class DetectionDone(object):
    """Marker in the index slot of the ('changeIndexes', x, idx) tuple: the upstream CBPoolMax2d
    already ran this layer's change detection (fused into its pooling kernel); the layer's raw
    bitmap, state and operand planes are up to date."""

    def __init__(self, owner, dense=False):
        self.owner = owner
        self.dense = dense        # True: a dense scan wrote every bitmap word (CBConv2d.detectInput)


class TailDone(object):
    """Marker in the index slot of the tuple: the upstream 1x1 CBConv2d already ran this layer's
    detection AND contraction inside its own kernel (cb_tail_update); this layer's state and output
    are up to date.  `candidates` is the superset list both layers walked."""

    def __init__(self, owner, candidates):
        self.owner = owner
        self.candidates = candidates


def _parse_input(inp):
    """tensor, or ('changeIndexes', tensor, indices) from an upstream CB layer (conv2d.py:180-187)."""
    if type(inp) == tuple:
        assert inp[0] == 'changeIndexes'
        return inp[1].detach(), inp[2]
    return inp.detach(), None


class CBPoolMax2d(nn.Module):
    def __init__(self, m):
        super(CBPoolMax2d, self).__init__()
        ks = m.kernel_size if isinstance(m.kernel_size, tuple) else (m.kernel_size,) * 2
        st = m.stride if isinstance(m.stride, tuple) else (m.stride,) * 2
        assert(ks == (2, 2) and st == (2, 2))
        self.stride = st
        self.kernel_size = ks
        self.ceil_mode = m.ceil_mode
        self.propChangeIndexes = False
        self.propPooledIndexes = False  # extension: hand pooled-resolution change candidates on
        self._fusedNext = []            # extension: [downstream CBConv2d] whose detection is fused in
        self.cloneOutput = True   # reference returns outputState.clone() (conv2d.py:73)
        self.register_buffer('outputState', torch.empty(0))
        self._stateBuf = None
        self.clearMemory()

    def clearMemory(self):
        self.outputState = self.outputState.new_empty(0)
        self._stateBuf = None
        self._scratch = None

    def getStateTensors(self):
        return [self.outputState] if hasattr(self, 'outputState') else []

    def _pooled_size(self, h, w):
        if self.ceil_mode:
            return (h - 1) // 2 + 1, (w - 1) // 2 + 1
        return h // 2, w // 2

    def _fusedPoolTarget(self, input_shape, dtype, device):
        """What an upstream CBConv2d on the tile path needs to pool inside its own epilogue and run the
        next layer's detection there (cb_conv_update_tiled_pool), or None when this pool's state is
        not allocated yet / was touched from outside, or the next layer's candidate-detection
        conditions do not hold (then the ordinary path runs and sets everything up)."""
        nxt = self._fusedNext[0] if getattr(self, '_fusedNext', None) else None
        if nxt is None or self._stateBuf is None:
            return None
        B, nc, h, w = input_shape
        oh, ow = self._pooled_size(h, w)
        st = self.outputState
        if list(st.shape) != [B, nc, oh, ow] or st.dtype != dtype or st.device != device \
                or st.data_ptr() != self._stateBuf.data_ptr() \
                or getattr(self, '_outVersion', None) != st._version:
            return None
        tgt = nxt._fusedDetectTarget(tuple(st.shape), dtype, device)
        if tgt is None:
            return None
        return dict(out=st, next_state=tgt['state'], next_raw_bits=tgt['raw_bits'],
                    threshold=tgt['threshold'], mode=tgt['mode'], aux=tgt['aux'])

    def forward(self, inp):
        assert(type(inp) == tuple and inp[0] == 'changeIndexes')
        input = inp[1].detach()
        changeIndexes = inp[2]
        assert input.dim() == 4
        _lib.require_cuda(input)
        if getattr(changeIndexes, 'pooledBy', None) is self:
            # the producing CBConv2d already pooled its tiles and ran the next layer's detection
            # inside its epilogue (cb_conv_update_tiled_pool)
            changeIndexes.pooledBy = None
            output = self.outputState.clone() if self.cloneOutput else self.outputState
            return 'changeIndexes', output, DetectionDone(self._fusedNext[0])
        B, nc, h, w = input.shape
        if not isinstance(changeIndexes, ChangeIndexes):
            assert(changeIndexes.dim() == 1)
            changeIndexes = ChangeIndexes.from_tensor(changeIndexes, (B, h, w))
        if self.ceil_mode:
            oh, ow = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        else:
            oh, ow = h // 2, w // 2
        if list(self.outputState.shape) != [B, nc, oh, ow] or self.outputState.dtype != input.dtype \
                or self.outputState.device != input.device or self._stateBuf is None \
                or self.outputState.data_ptr() != self._stateBuf.data_ptr():
            # deviation (stated): the reference allocates only once a non-empty index list arrives
            # (conv2d.py:53-62); deciding that needs a host sync, so the state is allocated up front.
            self.outputState, self._stateBuf = cg.pixel_major((B, nc, oh, ow), input.dtype,
                                                              input.device, _INF)
            self._scratch = None
        # candidates handed downstream are complete only if neither the producer's output nor this
        # pool's state was written from outside the kernels (see CBConv2d.forward_normal)
        complete = getattr(changeIndexes, 'complete', True) and \
            getattr(self, '_outVersion', None) == self.outputState._version
        self._outVersion = self.outputState._version
        nxt = self._fusedNext[0] if getattr(self, '_fusedNext', None) else None
        tgt = None
        if nxt is not None and complete and changeIndexes.bits is not None and input.stride(1) == 1:
            tgt = nxt._fusedDetectTarget(tuple(self.outputState.shape), input.dtype, input.device)
        if tgt is not None:
            # pooling + the next layer's detection in one kernel (cb_maxpool2x2_detect)
            cg.maxPool2d_detect(input, self.outputState, changeIndexes, tgt['state'], tgt['raw_bits'],
                                tgt['threshold'], tgt['mode'], aux=tgt['aux'])
            output = self.outputState.clone() if self.cloneOutput else self.outputState
            return 'changeIndexes', output, DetectionDone(nxt)
        cg.maxPool2d(input, self.outputState, changeIndexes, self.kernel_size, self.stride)

        output = self.outputState.clone() if self.cloneOutput else self.outputState
        if getattr(self, 'propPooledIndexes', False) and changeIndexes.bits is not None:
            # extension (no reference counterpart): change candidates at the pooled resolution for
            # a downstream CBConv2d with candidateDetect=True
            if self._scratch is None:
                self._scratch = cg.alloc_scratch((B, oh, ow), input.device)
            s = self._scratch
            cg.pool_compact(changeIndexes.bits, (B, h, w), (B, oh, ow), s["idx"], s["count"], s["ws"],
                            out_bits=s["dil_bits"])
            pooled = ChangeIndexes(s["idx"], s["count"], (B, oh, ow), bits=s["dil_bits"])
            pooled.complete = complete
            return 'changeIndexes', output, pooled
        if self.propChangeIndexes:
            # reference behaviour: the *input-resolution* indices are forwarded (conv2d.py:75-76)
            return 'changeIndexes', output, changeIndexes
        else:
            return output

    def __repr__(self):
        s = ('{name} (k={kernel_size}, s={stride}, ceil_mode={ceil_mode}')
        s += ', propChgIdxs={propChangeIndexes}'
        s += ')'
        return s.format(name=self.__class__.__name__, **self.__dict__)


class CBConv2d(nn.Module):
    #: contraction arithmetic (see include/cbinfer_b200.h).  gemmMode 'auto': fp32 data -> 3xBF16
    #: split on the tensor cores (max rel. error ~2e-5, bar 1e-4), 16-bit data -> kind::f16 with
    #: fp32 accumulation.  'tc3x' = 3xTF32 (~2e-6), 'tc' = single-pass TF32, 'simt' = fp32 FFMA.
    GEMM_MODES = {'simt': _lib.GEMM_SIMT_F32, 'tc': _lib.GEMM_TC, 'tc3x': _lib.GEMM_TC_3X,
                  'bf16x3': _lib.GEMM_TC_BF16X3}

    def __init__(self, m, threshold):
        super(CBConv2d, self).__init__()

        assert(m.groups == 1 and m.transposed == False)
        assert(m.output_padding == (0, 0) and m.padding == (m.kernel_size[-2] // 2, m.kernel_size[-1] // 2))
        assert(m.dilation == (1, 1) and m.stride == (1, 1))
        self.groups = m.groups
        self.transposed = m.transposed
        self.output_padding = m.output_padding
        self.padding = m.padding
        self.dilation = m.dilation
        self.stride = m.stride
        self.kernel_size = m.kernel_size
        self.in_channels = m.in_channels
        self.out_channels = m.out_channels

        assert(m.weight is not None and m.bias is not None)
        self.weight = m.weight
        self.bias = m.bias

        self.threshold = threshold

        self.register_buffer('prevInput', self.weight.data.new_empty(0))
        self.register_buffer('prevOutput', self.weight.data.new_empty(0))
        self.clearMemory()

        self.withReLU = False
        self.saveChangeMap = False
        self.propChangeIndexes = False
        self.gatherComputationStats = False
        self.finegrained = False
        self.copyInput = True
        self.feedbackLoop = False
        self.gemmMode = 'auto'
        # extension: treat indices received from upstream as *candidates* for the thresholded
        # detection instead of as the final change set (exact, see cb_change_detect_sparse)
        self.candidateDetect = False
        self.fuse1x1 = False      # extension: detect+compact in one launch for 1x1 layers
        # extension: (divisor, bias) applied to uint8 input frames inside the detection kernel
        self.inputNorm = None
        # extension: 1x1 layer on the candidate path without any compaction (see forward_normal)
        self.maskedConv = False
        # extension: contraction over dirty 8x16 output tiles (TMA-staged halo, implicit im2col)
        # instead of the per-pixel gather: 'auto' = where the library recommends it
        # (cb_conv_tiled_supported == 1; CBINFER_TILES=0 disables), 'on' = wherever supported, 'off'
        self.tileMode = 'auto'
        self._wsHolder = cg.ConvWorkspace()    # shared per model by pycbinfer.convert()

    # ---- state ---------------------------------------------------------------------------
    def clearMemory(self):
        # reference: fill_(-1e100).resize_(0) (conv2d.py:147-148; raises on torch >= 1.x)
        self.prevInput = self.weight.data.new_empty(0)
        self.prevOutput = self.weight.data.new_empty(0)
        self._inBuf = None        # pixel-major storage behind prevInput / prevOutput
        self._outBuf = None
        self._auxPlanes = None    # fp32 only: operand planes of prevInput for the 3x modes
        self._scratch = None      # bitmaps, index list, count, compaction workspace
        self._packed = None       # (key, packed weights, fp32 bias)
        self._fresh = True        # state holds +inf: the next detection must be a full scan
        self._lastThr = None
        self._lastChanges = None
        self._hintCache = None
        self.changeMap = None
        if getattr(self, '_wsHolder', None) is not None:
            self._wsHolder.clear()
        if hasattr(self, 'compStats'):
            self.compStats = None

    def getStateTensors(self):
        state = []
        if hasattr(self, 'prevInput'):
            state += [self.prevInput]
        if hasattr(self, 'prevOutput'):
            state += [self.prevOutput]
        return state

    def _gemm(self, dtype):
        mode = getattr(self, 'gemmMode', 'auto')
        if mode == 'auto':
            return _lib.GEMM_TC_BF16X3 if dtype == torch.float32 else _lib.GEMM_TC
        g = self.GEMM_MODES[mode]
        if g in (_lib.GEMM_TC_3X, _lib.GEMM_TC_BF16X3) and dtype != torch.float32:
            g = _lib.GEMM_TC
        return g

    def _aux(self, gemm):
        """Auxiliary operand planes of prevInput for the 3x modes (kept in step by the detection
        kernel); (re)built from prevInput when the mode changes or after a plain copy."""
        want = {_lib.GEMM_TC_3X: 'tf32', _lib.GEMM_TC_BF16X3: 'bf16'}.get(gemm) \
            if self.prevInput.dtype == torch.float32 else None
        if want is None:
            self._auxPlanes = None
        elif self._auxPlanes is None or self._auxPlanes[0] != want:
            if want == 'tf32':
                view, buf = cg.pixel_major(self.prevInput.shape, torch.float32, self.prevInput.device, 0)
                fin = torch.isfinite(self.prevInput)
                view.copy_(torch.where(fin, cg.tf32_lo(torch.where(fin, self.prevInput, 0)), 0))
                self._auxPlanes = ('tf32', view, buf)
            else:
                hi, lo = cg.bf16_planes(torch.nan_to_num(self._inBuf, posinf=0.0, neginf=0.0),
                                        self.in_channels)
                self._auxPlanes = ('bf16', hi, lo)
        return self._auxPlanes

    def _fusedDetectTarget(self, shape, dtype, device):
        """What an upstream CBPoolMax2d needs to run this layer's detection inside its own kernel,
        or None when the exactness conditions of candidate detection do not hold (fresh state,
        lowered threshold, different shape) or the layer is not on the candidate path."""
        if not getattr(self, 'candidateDetect', False) or self.finegrained or self._fresh \
                or self._inBuf is None or tuple(self.prevInput.shape) != tuple(shape) \
                or self.prevInput.dtype != dtype or self._lastThr is None \
                or getattr(self, '_inVersion', None) != self.prevInput._version \
                or self.threshold < self._lastThr or self._scratch is None \
                or not self._scratch.get("raw_clear", False):
            return None
        gemm, _, _ = self._weights(dtype, device)
        aux = self._aux(gemm)
        self._lastThr = self.threshold
        return dict(state=self.prevInput, raw_bits=self._scratch["raw_bits"],
                    threshold=self.threshold,
                    mode=_lib.UPDATE_CHANGED if self.feedbackLoop else _lib.UPDATE_ALL,
                    aux=None if aux is None else (aux[:2] if aux[0] == 'tf32' else aux))

    def _weights(self, dtype, device):
        gemm = self._gemm(dtype)
        key = (self.weight.data_ptr(), self.weight._version, self.bias.data_ptr(),
               self.bias._version, dtype, gemm, str(device))
        if self._packed is None or self._packed[0] != key:
            w = self.weight.detach().to(device=device, dtype=dtype)
            self._packed = (key, cg.pack_weights(w, gemm),
                            self.bias.detach().to(device=device, dtype=torch.float32).contiguous())
        return gemm, self._packed[1], self._packed[2]

    # ---- fine-grained path (conv2d.py:160-176) ---------------------------------------------
    def forward_fg(self, inp):
        """Reference flow: the first frame is a dense convolution (:163-167); afterwards every input
        VALUE whose change exceeds the threshold pushes W*delta into (a clone of) prevOutput
        (cbconvFG, :169-170) and prevInput = input (:175).  Here: per-value thresholded delta planes
        (cb_fg_detect) -> touched output pixels (cb_dilate_compact) -> accumulating tcgen05
        contraction over them (cb_conv_accumulate); gemmMode='simt' keeps the scattered
        red.global.add kernel on planar tensors (cb_fg_update, exact fp32 products)."""
        input = inp.detach()
        _lib.require_cuda(input)
        if input.dtype != torch.float32:
            raise _lib.CBinferError("the fine-grained path is fp32 only (as in the reference)")
        if getattr(self, 'gemmMode', 'auto') == 'simt':
            return self._forward_fg_planar(input.contiguous())
        B, _, H, W = input.shape
        dev = input.device
        outpSize = (B, self.out_channels, H, W)
        gemm = _lib.GEMM_TC_BF16X3
        key = (self.weight.data_ptr(), self.weight._version, 'fg', str(dev))
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, cg.pack_weights(self.weight.detach().to(device=dev, dtype=torch.float32), gemm),
                            None)
        fresh = self.prevInput.size() != input.size() or self._inBuf is None or self._outBuf is None \
            or self.prevInput.device != dev or self.prevInput.data_ptr() != self._inBuf.data_ptr() \
            or tuple(self.prevOutput.size()) != outpSize or self.prevOutput.data_ptr() != self._outBuf.data_ptr() \
            or self._auxPlanes is None or self._auxPlanes[0] != 'fg'
        if fresh:
            self.prevOutput, self._outBuf = cg.pixel_major(outpSize, torch.float32, dev, 0)
            self.prevOutput.copy_(F.conv2d(input, self.weight.detach(),
                                           padding=tuple(s // 2 for s in self.weight.size()[2:]),
                                           bias=self.bias.detach()))
            self.prevInput, self._inBuf = cg.pixel_major(input.shape, torch.float32, dev, 0)
            self.prevInput.copy_(input)
            p16 = _lib.C.cb_plane_pitch16(self.in_channels)
            self._auxPlanes = ('fg', torch.zeros(B, H, W, p16, dtype=torch.bfloat16, device=dev),
                               torch.zeros(B, H, W, p16, dtype=torch.bfloat16, device=dev))
            self._scratch = cg.alloc_scratch((B, H, W), dev)
            self._scratch["nvalues"] = torch.zeros(1, dtype=torch.int32, device=dev)
        else:
            s = self._scratch
            planes = self._auxPlanes[1:]
            cg.fg_detect(input, self.prevInput, self._inBuf, planes, s["raw_bits"], self.threshold,
                         count=s["nvalues"])
            cg.dilate_compact(s["raw_bits"], (B, H, W), self.kernel_size, s["idx"], s["count"], s["ws"],
                              dil_bits=s["dil_bits"])
            changes = ChangeIndexes(s["idx"], s["count"], (B, H, W), bits=s["dil_bits"])
            cg.conv_accumulate(planes, changes, self._packed[1], self._outBuf, self.in_channels,
                               self.out_channels, self.kernel_size, gemm, ws=self._workspace(dev))
            self._lastChanges = changes
        # the reference updates a clone (:169), so outputs handed out earlier stay as they were
        return F.relu(self.prevOutput) if self.withReLU else self.prevOutput.clone()

    def _forward_fg_planar(self, input):
        if self.prevInput.size() != input.size() or not self.prevInput.is_contiguous() \
                or not self.prevOutput.is_contiguous():
            # init prevOutput with a dense convolution, as the reference does (conv2d.py:163-167)
            self.prevOutput = F.conv2d(input, self.weight.detach(),
                                       padding=tuple(s // 2 for s in self.weight.size()[2:]),
                                       bias=self.bias.detach()).contiguous()
            self.prevInput = input.clone()
            self._inBuf = self._outBuf = None
        else:
            po = self.prevOutput.clone()                     # conv2d.py:169
            # also performs prevInput = input (conv2d.py:175) in the same pass
            self.prevOutput = cbconvFG(input, self.prevInput, po, self.weight, self.threshold)
        outp = self.prevOutput
        if self.withReLU:
            outp = F.relu(outp)
        return outp

    # ---- coarse-grained path (conv2d.py:178-259) ---------------------------------------------
    def forward_normal(self, inp):
        input, changeIndexes = _parse_input(inp)
        assert(input.dim() == 4)
        assert(input.size(-3) == self.in_channels)
        # a layer fed with plain frames re-scans them densely every frame, which rewrites every word of
        # its raw bitmap: nothing ever ORs bits into it, so its compaction need not zero it afterwards
        plain_input = changeIndexes is None or (isinstance(changeIndexes, DetectionDone)
                                                and getattr(changeIndexes, 'dense', False))
        _lib.require_cuda(input)
        B, _, H, W = input.shape
        dev, dt = input.device, input.dtype
        # uint8 frames (decoder output) are normalised inside the detection kernel:
        # value = u8 / divisor + bias with inputNorm = (divisor, bias); the layer computes in fp32
        u8norm = None
        if dt == torch.uint8:
            u8norm = getattr(self, 'inputNorm', None)
            if u8norm is None:
                raise _lib.CBinferError("uint8 input needs CBConv2d.inputNorm = (divisor, bias), e.g. "
                                        "(255.0, 0.0) for the scene reader's frame/255")
            if (changeIndexes is not None and not isinstance(changeIndexes, DetectionDone)) \
                    or self.gatherComputationStats or self.in_channels > 4:
                input = input.float().div(u8norm[0]).add(u8norm[1])     # rare paths: plain fp32
                u8norm = None
            dt = torch.float32

        # (model.half() / .to(device) on a warm model replace the registered buffers with fresh
        #  tensors that no longer alias the pixel-major storage: treat that as a reset)
        outpSize = (B, self.out_channels, H, W)
        need_out = tuple(self.prevOutput.size()) != outpSize or self.prevOutput.dtype != dt \
            or self._outBuf is None or self.prevOutput.device != dev \
            or self.prevOutput.data_ptr() != self._outBuf.data_ptr()
        need_in = need_out or self.prevInput.size() != input.size() or self.prevInput.dtype != dt \
            or self._inBuf is None or self.prevInput.device != dev \
            or self.prevInput.data_ptr() != self._inBuf.data_ptr()
        if need_in:       # (a fresh output map needs every pixel recomputed: the input state goes too)
            self.prevInput, self._inBuf = cg.pixel_major(input.shape, dt, dev, _INF)   # :192-194
            self._auxPlanes = None
            self._scratch = None
            self._fresh = True
        if need_out:
            self.prevOutput, self._outBuf = cg.pixel_major(outpSize, dt, dev, _INF)    # :195-199
        # State tensors written from outside this module's kernels (the reference's eval scripts
        # snapshot and restore getStateTensors() with copy_, poseDetection/eval03.py:87-95): torch
        # bumps the tensor version, our kernels do not.  A touched prevInput is treated as fresh
        # (dense scan, operand planes rebuilt); a touched prevOutput makes this frame's index list
        # an incomplete candidate set for the next layer.
        if getattr(self, '_inVersion', None) != self.prevInput._version:
            self._fresh = True
            self._auxPlanes = None
        ext_out = getattr(self, '_outVersion', None) != self.prevOutput._version
        if self._scratch is None:
            self._scratch = cg.alloc_scratch((B, H, W), dev, want_map=self.saveChangeMap)
        if self.saveChangeMap and "dil_map" not in self._scratch:
            self._scratch["dil_map"] = torch.zeros(B, H, W, dtype=torch.int8, device=dev)
        s = self._scratch

        if self.gatherComputationStats:
            self._gatherStats(input)

        gemm, packed, bias32 = self._weights(dt, dev)
        if isinstance(changeIndexes, TailDone):
            # the upstream 1x1 layer's kernel already did this layer's frame (cb_tail_update)
            assert changeIndexes.owner is self
            self._lastChanges = changeIndexes.candidates
            self._inVersion = self.prevInput._version
            self._outVersion = self.prevOutput._version
            if self.propChangeIndexes:
                return 'changeIndexes', self.prevOutput, changeIndexes.candidates
            return self.prevOutput
        tail = self._tryTail(input, changeIndexes, gemm, packed, bias32, s, outpSize, ext_out)
        if tail is not None:
            return tail
        aux = self._aux(gemm)
        aux_arg = None if aux is None else (aux[:2] if aux[0] == 'tf32' else aux)

        candidates = None
        mask = None
        tiled = False             # this frame's contraction walks the dirty-tile list
        use_tiles = self._useTiles(dt, gemm, (B, H, W))
        # a directly following CBPoolMax2d (+ the detection of the layer after it) rides along in the
        # tile kernel's epilogue when everything is warmed up (cb_conv_update_tiled_pool); nobody
        # then needs this layer's ordered index list, so it is only compacted on demand
        pool_args = None
        fp = getattr(self, '_fusedPool', None)
        if use_tiles and fp and self.propChangeIndexes and not ext_out \
                and os.environ.get("CBINFER_FUSE_POOL", "1") != "0" \
                and _lib.C.cb_conv_tiled_pool_supported(_lib.dtype_code(dt), gemm, self.out_channels):
            pool_args = fp[0]._fusedPoolTarget(outpSize, dt, dev)
        detected = isinstance(changeIndexes, DetectionDone)
        if detected:
            assert changeIndexes.owner is self
            changeIndexes = None
        if changeIndexes is not None and getattr(self, 'candidateDetect', False):
            candidates, changeIndexes = changeIndexes, None
            if not isinstance(candidates, ChangeIndexes):
                candidates = ChangeIndexes.from_tensor(candidates.detach(), (B, H, W))
            if self._fresh or self._lastThr is None or self.threshold < self._lastThr \
                    or tuple(candidates.shape) != (B, H, W) or not getattr(candidates, 'complete', True):
                candidates = None          # exactness conditions not met: full scan

        if changeIndexes is None:
            # reference: detect (feedback updates prevInput at changed pixels, :222-224), nonzero
            # (:232), then prevInput.copy_(input) when not in feedback mode (:234-238).  Here the
            # copy is part of the detection pass; copyInput=False (alias the input as state) is
            # honoured as a copy -- the state always owns its memory.
            mode = _lib.UPDATE_CHANGED if self.feedbackLoop else _lib.UPDATE_ALL
            sparse_next = bool(getattr(self, 'candidateDetect', False)) and not plain_input
            # (measured: one launch fewer, but its serial compaction tail makes it slower than the
            #  two-kernel path once there are >~10k candidates, so it is opt-in: fuse1x1=True)
            fused11 = (candidates is not None and tuple(self.kernel_size) == (1, 1)
                       and getattr(self, 'fuse1x1', False)
                       and not self.saveChangeMap and input.stride(1) == 1)
            # extension: no compaction at all for a 1x1 layer on the candidate path - the
            # contraction walks the candidate list and skips pixels whose raw bit is not set
            # (cb_conv_update_masked); this layer's own index list is then never materialised
            masked = (candidates is not None and not detected and tuple(self.kernel_size) == (1, 1)
                      and getattr(self, 'maskedConv', False) and not fused11
                      and not self.saveChangeMap and gemm != _lib.GEMM_SIMT_F32)
            fusedSmall = False
            if detected:
                pass                     # the upstream pool kernel already did it
            elif fused11:
                # no dilation needed: candidate detection + ordered compaction in one launch
                if "dcs_ws" not in s:
                    s["dcs_ws"] = torch.zeros(_lib.C.cb_detect_compact_ws_bytes(B, H, W),
                                              dtype=torch.uint8, device=dev)
                cg.detect_compact_sparse(input, self.prevInput, self.threshold, mode, candidates,
                                         s["idx"], s["count"], s["dcs_ws"], aux=aux_arg)
            elif candidates is not None and not masked and not use_tiles and not self.saveChangeMap \
                    and input.stride(1) == 1 and self._smallMap(B, H, W):
                # small map: candidate detection + dilation + ordered compaction in ONE launch (the
                # last block of the detection kernel compacts), cb_change_detect_sparse_compact
                if "sdc_sync" not in s:
                    s["sdc_sync"] = torch.zeros(1, dtype=torch.int32, device=dev)
                cg.detect_sparse_compact(input, self.prevInput, s["raw_bits"], self.threshold, mode, candidates,
                                         self.kernel_size, s["idx"], s["count"], s["sdc_sync"], aux=aux_arg,
                                         bits_are_clear=s.get("raw_clear", False), dil_bits=s["dil_bits"],
                                         clear_raw=sparse_next)
                fusedSmall = True
            elif candidates is not None:
                cg.detect_sparse(input, self.prevInput, s["raw_bits"], self.threshold, mode,
                                 candidates, aux=aux_arg,
                                 bits_are_clear=s.get("raw_clear", False))
                if masked:
                    if "mask_sync" not in s:
                        s["mask_sync"] = torch.zeros(2, dtype=torch.int32, device=dev)
                    mask = dict(bits=s["raw_bits"], clear=True, count=s["count"], sync=s["mask_sync"])
            elif u8norm is not None:
                cg.detect_u8(input, self.prevInput, s["raw_bits"], self.threshold, mode,
                             u8norm[0], u8norm[1], aux=aux_arg)
            else:
                cg.detect(input, self.prevInput, s["raw_bits"], self.threshold, mode, aux=aux_arg)
            self._fresh = False
            self._lastThr = self.threshold
            if not detected and fused11:
                s["raw_clear"] = False
                changeIndexes = ChangeIndexes(s["idx"], s["count"], (B, H, W), bits=None)
            elif fusedSmall:
                s["raw_clear"] = sparse_next
                changeIndexes = ChangeIndexes(s["idx"], s["count"], (B, H, W), bits=s["dil_bits"])
            elif mask is not None:
                s["raw_clear"] = True                # the masked contraction clears the bitmap
                changeIndexes = ChangeIndexes(candidates.buffer, candidates.count, (B, H, W), bits=None)
                changeIndexes.superset = True
            else:
                changeIndexes = self._compact(s, B, H, W, sparse_next, tiles=use_tiles,
                                              lazy=pool_args is not None and not self.saveChangeMap
                                              and os.environ.get("CBINFER_LAZY_LIST", "1") != "0")
                tiled = use_tiles
        else:
            if not isinstance(changeIndexes, ChangeIndexes):
                assert(changeIndexes.dim() == 1)
                changeIndexes = ChangeIndexes.from_tensor(changeIndexes.detach(), (B, H, W))
            elif tuple(changeIndexes.shape) != (B, H, W):
                # e.g. a CBPoolMax2d with the reference's propChangeIndexes forwards INPUT-resolution
                # indices (conv2d.py:75-76): using them here would scatter out of bounds
                raise _lib.CBinferError("change indexes refer to a %s grid, this layer's map is %s"
                                        % (tuple(changeIndexes.shape), (B, H, W)))
            if not self.feedbackLoop:
                self.prevInput.copy_(input)                                            # :234-236
                self._auxPlanes = None
                aux = self._aux(gemm)
            # (with feedbackLoop the reference never refreshes prevInput here either, :220,234)

        lo_buf = aux[2] if aux is not None and aux[0] == 'tf32' else None
        planes16 = aux[1:] if aux is not None and aux[0] == 'bf16' else None
        if tiled:
            # spatially clustered change sets: contraction over the dirty 8x16 tiles (TMA-staged
            # halo, implicit im2col through the UMMA descriptors), rows masked by the dilated bitmap
            cg.conv_update_tiled(self._inBuf, s["tile_ws"], s["dil_bits"], packed, bias32, self._outBuf,
                                 self.in_channels, self.out_channels, self.kernel_size, self.withReLU,
                                 gemm, lo_buf=lo_buf, planes16=planes16, pool=pool_args,
                                 self_list=changeIndexes.__dict__.pop('selfList', None))  # :242-251
            if pool_args is not None:
                changeIndexes.pooledBy = fp[0]
        else:
            cg.conv_update(self._inBuf, changeIndexes, packed, bias32, self._outBuf, self.in_channels,
                           self.out_channels, self.kernel_size, self.withReLU, gemm,
                           lo_buf=lo_buf, planes16=planes16,
                           ws=self._workspace(dev), mask=mask)                         # :242-251
        self._inVersion = self.prevInput._version
        self._outVersion = self.prevOutput._version
        self._lastChanges = changeIndexes
        if ext_out and isinstance(changeIndexes, ChangeIndexes):
            changeIndexes.complete = False
        elif isinstance(changeIndexes, ChangeIndexes):
            changeIndexes.complete = True

        if self.propChangeIndexes:
            return 'changeIndexes', self.prevOutput, changeIndexes
        else:
            return self.prevOutput

    # ---- two chained 1x1 layers in one launch (extension) ------------------------------------------
    def _tailTarget(self, shape, dt, dev, feedback):
        """What the 1x1 layer feeding this one needs to run this layer's frame inside its own kernel
        (cb_tail_update), or None while the exactness conditions of candidate detection do not hold
        for this layer (fresh state, lowered threshold, state touched from outside, ...)."""
        if not (getattr(self, 'candidateDetect', False) and getattr(self, 'maskedConv', False)) \
                or self.finegrained or self.saveChangeMap or self.gatherComputationStats \
                or getattr(self, 'fuse1x1', False) \
                or tuple(self.kernel_size) != (1, 1) or self.feedbackLoop != feedback \
                or self._fresh or self._inBuf is None or self._outBuf is None or self._scratch is None \
                or tuple(self.prevInput.shape) != tuple(shape) or self.prevInput.dtype != dt \
                or self.prevInput.device != dev or self._lastThr is None or self.threshold < self._lastThr \
                or getattr(self, '_inVersion', None) != self.prevInput._version \
                or getattr(self, '_outVersion', None) != self.prevOutput._version \
                or self.prevInput.data_ptr() != self._inBuf.data_ptr() \
                or self.prevOutput.data_ptr() != self._outBuf.data_ptr() \
                or not self._scratch.get("raw_clear", False):
            return None
        gemm, packed, bias32 = self._weights(dt, dev)
        return dict(gemm=gemm, packed=packed, bias=bias32, state=self._inBuf, out=self._outBuf,
                    thr=self.threshold, relu=self.withReLU, count=self._scratch["count"])

    def _tryTail(self, input, changeIndexes, gemm, packed, bias32, s, outpSize, ext_out):
        ft = getattr(self, '_fusedTail', None)
        if not ft or os.environ.get("CBINFER_FUSE_TAIL", "1") == "0":
            return None
        B, C0, H, W = input.shape
        if not isinstance(changeIndexes, ChangeIndexes) or not getattr(self, 'candidateDetect', False) \
                or not getattr(self, 'maskedConv', False) or tuple(self.kernel_size) != (1, 1) \
                or getattr(self, 'fuse1x1', False) \
                or self.saveChangeMap or self.gatherComputationStats or ext_out or input.dtype != torch.float32 \
                or self._fresh or self._lastThr is None or self.threshold < self._lastThr \
                or tuple(changeIndexes.shape) != (B, H, W) or not getattr(changeIndexes, 'complete', True) \
                or not s.get("raw_clear", False):
            return None
        nxt = ft[0]
        C1, C2 = self.out_channels, nxt.out_channels
        if not _lib.C.cb_tail_supported(_lib.dtype_code(input.dtype), gemm, C0, C1, C2):
            return None
        # the input must be the upstream layer's pixel-major map itself (pitch == channels)
        if input.stride() != (H * W * C0, 1, W * C0, C0) or self._inBuf.shape[3] != C0 \
                or self._outBuf.shape[3] != C1:
            return None
        tgt = nxt._tailTarget(outpSize, input.dtype, input.device, self.feedbackLoop)
        if tgt is None or tgt['gemm'] != gemm:
            return None
        if "tail_sync" not in s:
            s["tail_sync"] = torch.zeros(4, dtype=torch.int32, device=input.device)
        x_buf = input.permute(0, 2, 3, 1)
        cg.tail_update(x_buf, self._inBuf, packed, bias32, self._outBuf, self.withReLU, self.threshold,
                       tgt['state'], tgt['packed'], tgt['bias'], tgt['out'], tgt['relu'], tgt['thr'],
                       changeIndexes, C0, C1, C2,
                       _lib.UPDATE_CHANGED if self.feedbackLoop else _lib.UPDATE_ALL,
                       s["count"], tgt['count'], s["tail_sync"])
        # the kernel splits its operands on the fly: the layers' operand planes are stale now
        self._auxPlanes = None
        nxt._auxPlanes = None
        self._lastThr = self.threshold
        nxt._lastThr = nxt.threshold
        self._inVersion = self.prevInput._version
        self._outVersion = self.prevOutput._version
        sup = ChangeIndexes(changeIndexes._buffer, changeIndexes.count, (B, H, W), bits=changeIndexes.bits,
                            ws=changeIndexes._ws, listed=changeIndexes.listed)
        sup.superset = True
        self._lastChanges = sup
        return 'changeIndexes', self.prevOutput, TailDone(nxt, sup)

    def lastChangeIndexes(self):
        """This frame's change indexes as an int32 tensor (ascending b*H*W + y*W + x): what the
        reference hands on as ``changeIndexes`` (conv2d.py:256-257).  The list of a layer on the
        tile path is compacted only now, on demand."""
        ci = getattr(self, '_lastChanges', None)
        if ci is None:
            raise _lib.CBinferError("no frame processed yet")
        return ci.tensor() if isinstance(ci, ChangeIndexes) else ci

    def _workspace_holder(self):
        if getattr(self, '_wsHolder', None) is None:
            self._wsHolder = cg.ConvWorkspace()
        return self._wsHolder

    def _workspace(self, dev):
        return self._workspace_holder().get(dev)

    def detectInput(self, input):
        """Run only this layer's (dense) change detection on `input` and return the tuple that makes
        the following ``forward`` skip it: ``model(first.detectInput(frame))``.  Lets a pipeline read
        every frame where it lies (e.g. straight from a decoder's output buffer) while the rest of
        the model replays as one CUDA graph captured on that tuple - no copy into a static input
        tensor.  Needs a warmed-up layer (state allocated by an ordinary forward of this shape)."""
        input = input.detach()
        _lib.require_cuda(input)
        B, _, H, W = input.shape
        dt = torch.float32 if input.dtype == torch.uint8 else input.dtype
        if self._inBuf is None or self._scratch is None or self.prevInput.size() != input.size() \
                or self.prevInput.dtype != dt or self.finegrained or self.gatherComputationStats:
            raise _lib.CBinferError("detectInput: run one ordinary forward of this input shape first")
        if getattr(self, '_inVersion', None) != self.prevInput._version:
            self._fresh = True
            self._auxPlanes = None
        gemm, _, _ = self._weights(dt, input.device)
        aux = self._aux(gemm)
        aux_arg = None if aux is None else (aux[:2] if aux[0] == 'tf32' else aux)
        mode = _lib.UPDATE_CHANGED if self.feedbackLoop else _lib.UPDATE_ALL
        if input.dtype == torch.uint8:
            norm = getattr(self, 'inputNorm', None)
            if norm is None or self.in_channels > 4:
                raise _lib.CBinferError("detectInput: uint8 input needs inputNorm and <= 4 channels")
            cg.detect_u8(input, self.prevInput, self._scratch["raw_bits"], self.threshold, mode,
                         norm[0], norm[1], aux=aux_arg)
        else:
            cg.detect(input, self.prevInput, self._scratch["raw_bits"], self.threshold, mode, aux=aux_arg)
        self._fresh = False
        self._lastThr = self.threshold
        self._inVersion = self.prevInput._version
        return 'changeIndexes', input, DetectionDone(self, dense=True)

    def _smallMap(self, B, H, W):
        """bitmap small enough for the one-launch detection + compaction (cb_change_detect_sparse_compact)?"""
        if os.environ.get("CBINFER_FUSE_SMALL", "1") == "0":
            return False
        return 0 < _lib.C.cb_bitmap_words(B, H, W) <= _lib.C.cb_compact_small_max_words()

    def _useTiles(self, dt, gemm, shape):
        """tile path for this layer / shape?  (tileMode 'auto': where the library recommends it)"""
        mode = getattr(self, 'tileMode', 'auto')
        if mode == 'off' or os.environ.get("CBINFER_TILES", "1") == "0" or gemm == _lib.GEMM_SIMT_F32:
            return False
        key = (dt, gemm, tuple(shape), mode)
        cache = self.__dict__.setdefault('_tileOk', {})
        if key not in cache:
            sup = cg.tiled_supported(dt, gemm, shape, self.in_channels, self.out_channels, self.kernel_size)
            cache[key] = sup >= 1 if mode == 'on' else sup == 1
        return cache[key]

    def _hints(self, B, H, W, dev):
        """L2 prefetch hints for this frame's dilation (cb_dilate_compact_hinted): the previous-input
        state of the CB layers that will threshold the pixels this layer is about to rewrite
        (``_prefetchNext``, wired by models.enableCandidateDetection).  Those rows were last touched
        when the pixels last changed, so the consumer would otherwise fetch them from DRAM inside a
        latency-bound kernel; the dilation knows the pixels one contraction earlier."""
        nxt = getattr(self, '_prefetchNext', None)
        # (opt-in: measured on the bench model the hinted L3 dilation costs 12 us more and saves the tail
        #  kernel 2 us -- profiles/r02_experiments.md)
        if not nxt or os.environ.get("CBINFER_PREFETCH", "0") != "1":
            return None
        tg = []
        for m, sh in nxt:
            buf = getattr(m, '_inBuf', None)
            if buf is None or buf.device != dev or buf.dim() != 4 or buf.shape[0] != B:
                continue
            th, tw = buf.shape[1], buf.shape[2]
            if (sh == 0 and (th, tw) == (H, W)) or \
                    (sh == 1 and th in (H // 2, (H + 1) // 2) and tw in (W // 2, (W + 1) // 2)):
                tg.append((buf, sh))
        key = tuple((t.data_ptr(), tuple(t.shape), sh) for t, sh in tg)
        hc = getattr(self, '_hintCache', None)
        if hc is None or hc[0] != key:
            hc = (key, cg.PrefetchHints(tg))
            self._hintCache = hc
        return hc[1] if hc[1].n else None

    def _selfTiles(self, B, H, W):
        """dilation + tile list inside the tile contraction itself (cb_conv_update_tiled_self)?"""
        # opt-in (CBINFER_SELF_TILES=1; =N > 1: only bitmaps of at most N words): bit-identical, but no faster --
        # the runtime admits ONE block of a TMEM-allocating kernel per SM to a cooperative launch, and even
        # at full occupancy (plain launch, CBINFER_SELF_COOP=0) the step does not gain (profiles/r02_experiments.md)
        lim = int(os.environ.get("CBINFER_SELF_TILES", "0") or 0)
        if lim <= 0 or (lim > 1 and _lib.C.cb_bitmap_words(B, H, W) > lim):
            return False
        return bool(_lib.C.cb_conv_tiled_self_supported(B, H, W, self.kernel_size[0], self.kernel_size[1]))

    def _compact(self, s, B, H, W, sparse_next, tiles=False, lazy=False):
        """dilate the raw bitmap by the filter footprint and compact it to the index list."""
        if tiles and "tile_ws" not in s:
            s["tile_ws"] = cg.alloc_tile_ws((B, H, W), s["idx"].device)
        hints = self._hints(B, H, W, s["idx"].device)
        if tiles and lazy and self._selfTiles(B, H, W):
            # nothing to launch: the tile contraction dilates the raw bitmap and lists its tiles itself
            # (cb_conv_update_tiled_self); bitmap, tile list and count exist once it has run
            s["raw_clear"] = sparse_next
            ci = ChangeIndexes(s["idx"], s["count"], (B, H, W), bits=s["dil_bits"], ws=s["ws"], listed=False)
            ci.selfList = dict(raw_bits=s["raw_bits"], count=s["count"], ws=s["ws"], clear_raw=sparse_next)
            return ci
        if tiles and lazy:
            # tiles + dilated bitmap + count only; the ordered list is compacted if somebody asks
            cg.dilate_tiles(s["raw_bits"], (B, H, W), self.kernel_size, s["count"], s["ws"], s["dil_bits"],
                            s["tile_ws"], clear_raw=sparse_next, hints=hints)
            s["raw_clear"] = sparse_next
            return ChangeIndexes(s["idx"], s["count"], (B, H, W), bits=s["dil_bits"], ws=s["ws"], listed=False)
        dil_map = s.get("dil_map") if self.saveChangeMap else None
        # a layer on the candidate path lets the compaction zero the raw bitmap once it has been
        # consumed, so the next frame's candidate detection needs no memset
        cg.dilate_compact(s["raw_bits"], (B, H, W), self.kernel_size, s["idx"], s["count"],
                          s["ws"], dil_bits=s["dil_bits"], dil_map=dil_map, clear_raw=sparse_next,
                          tile_ws=s["tile_ws"] if tiles else None, hints=hints)
        s["raw_clear"] = sparse_next
        if self.saveChangeMap:
            self.changeMap = dil_map[0] if B == 1 else dil_map
        return ChangeIndexes(s["idx"], s["count"], (B, H, W), bits=s["dil_bits"])

    def _gatherStats(self, input):
        """op-count bookkeeping of conv2d.py:201-218 (dense torch ops, only when enabled)."""
        changeTensor = (input - self.prevInput).abs().gt(self.threshold)
        ones = torch.ones(changeTensor.size(-3), 1, self.weight.size(2), self.weight.size(3),
                          device=input.device)
        proped = F.conv2d(changeTensor.float(), ones, groups=changeTensor.size(-3)).gt(0)
        opsPerValue = self.weight.size(0) * self.weight.size(2) * self.weight.size(3) * 2
        nC = changeTensor.size(-3)
        self.compStats = dict(
            numInputChangesPerFeatureMap=changeTensor.sum() * opsPerValue,
            numInputChanges=changeTensor.sum(-3).gt(0).sum() * nC * opsPerValue,
            numInputPropedChangesPerFeatureMap=proped.sum() * opsPerValue,
            numInputPropedChanges=proped.sum(-3).gt(0).sum() * nC * opsPerValue,
            totalInputValues=changeTensor.size(-1) * changeTensor.size(-2) * nC * opsPerValue,
        )

    def forward(self, inp):
        self._setDefaultValues()
        with torch.no_grad():
            if self.finegrained:
                assert(self.feedbackLoop == False)
                return self.forward_fg(inp)
            else:
                return self.forward_normal(inp)

    def __repr__(self):
        self._setDefaultValues()
        s = ('{name} (th={threshold}, {in_channels}->{out_channels}, k={kernel_size}'
             ', s={stride}, copyInput={copyInput}')
        if self.padding != (0,) * len(self.padding):
            s += ', pad={padding}'
        if self.dilation != (1,) * len(self.dilation):
            s += ', dilation={dilation}'
        if self.output_padding != (0,) * len(self.output_padding):
            s += ', outpad={output_padding}'
        if self.groups != 1:
            s += ', grp={groups}'
        if self.bias is None:
            s += ', bias=False'
        if self.withReLU:
            s += ', withReLU={withReLU}'
        s += ', propChgIdxs={propChangeIndexes}'
        s += ')'
        return s.format(name=self.__class__.__name__, **self.__dict__)

    def _setDefaultValues(self):
        for name, val in (('saveChangeMap', False), ('propChangeIndexes', False),
                          ('gatherComputationStats', False), ('finegrained', False),
                          ('copyInput', True), ('feedbackLoop', False), ('gemmMode', 'auto'),
                          ('candidateDetect', False), ('fuse1x1', False), ('inputNorm', None),
                          ('maskedConv', False), ('tileMode', 'auto')):
            if not(hasattr(self, name)):
                setattr(self, name, val)
