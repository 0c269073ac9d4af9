"""Coarse-grained change-based ops: python wrappers over the sm_100a C ABI.

Mirrors the op surface of the reference's ``pycbinfer/conv2d_cg.py`` -- same function names,
argument meaning and tensor layouts (planar NCHW, batch 1, int8 change maps, ascending int32
indices) -- so code and tests written against the reference read the same here:

    changeDetection      conv2d_cg.py:100-122      changePropagation   conv2d_cg.py:159-177
    changeIndexesExtr    conv2d_cg.py:200-213      genXMatrix          conv2d_cg.py:239-261
    matrixMult           conv2d_cg.py:342-349      updateOutput        conv2d_cg.py:292-313
    maxPool2d            conv2d_cg.py:58-82

Differences, all deliberate: no ``useHalf`` flag (the dtype is taken from the tensor), launch
geometry lives in the library, the change count never has to visit the host
(:class:`ChangeIndexes`), and every op is CUDA-only -- the reference's ``*_python`` CPU twins
exist in this repo only inside ``oracle/`` as the parity checker.

The fused per-frame path used by ``CBConv2d`` (``detect`` -> ``dilate_compact`` ->
``conv_update``) is exposed as well.
"""
import os

import torch

from . import _lib
from ._lib import C, check, dtype_code, require_cuda, stream_ptr


def _strides4(t):
    """(sb, sc, sy, sx) element strides of a [B,C,H,W] tensor."""
    assert t.dim() == 4
    return t.stride(0), t.stride(1), t.stride(2), t.stride(3)


class ChangeIndexes(object):
    """Device-resident change-index list: ``buffer[:count]`` holds ascending int32 pixel indices
    ``b*H*W + y*W + x``.  It stands in for the exact-length tensor the reference passes between
    layers (conv2d.py:256-257) without forcing the device->host sync ``torch.nonzero`` implies
    (conv2d_cg.py:202): consumers in this package read ``count`` on the device; anything that
    needs a real tensor (``len``, ``.tensor()``, indexing) synchronises on demand."""

    def __init__(self, buffer, count, shape, bits=None, ws=None, listed=True):
        self._buffer = buffer         # int32 [capacity]
        # listed=False: only the dilated bitmap, the tile list and the count exist so far (cb_dilate_tiles);
        # the ordered list is compacted from `bits` the first time somebody asks for it
        self.listed = listed
        self._ws = ws                 # compaction workspace for that
        self.count = count            # int32 [1] on the device
        self.shape = shape            # (B, H, W) the indices refer to
        self.bits = bits              # optional dilated bitmap the list was compacted from
        # False when the producer's output was also modified outside its own kernels (e.g. a state
        # snapshot restored with copy_): the list is then not a complete candidate set downstream
        self.complete = True
        # True when the list is only a superset of the changed pixels (a masked 1x1 layer hands its
        # own candidates on instead of compacting): fine as detection candidates, not an exact list
        self.superset = False

    @property
    def buffer(self):
        if not self.listed:
            # compacted on every access, never cached: under CUDA-graph replay this python object
            # outlives the frame it was created for, while `bits` always holds the current frame
            assert self.bits is not None and self._ws is not None
            dilate_compact(self.bits, self.shape, (1, 1), self._buffer, self.count, self._ws)
        return self._buffer

    @classmethod
    def from_tensor(cls, idx, shape):
        idx = idx.to(torch.int32).contiguous()
        count = torch.full((1,), idx.numel(), dtype=torch.int32, device=idx.device)
        if idx.numel() == 0:                  # keep a valid device pointer for the C ABI
            idx = torch.zeros(1, dtype=torch.int32, device=idx.device)
        return cls(idx, count, shape)

    def tensor(self):
        if self.superset:
            raise _lib.CBinferError("this change list is a candidate superset (producer ran with "
                                    "maskedConv=True); set maskedConv=False on it for exact lists")
        buf = self.buffer
        return buf[: int(self.count.item())]

    # tensor-like conveniences (all synchronise)
    def __len__(self):
        return int(self.count.item())

    def numel(self):
        return len(self)

    def dim(self):
        return 1

    def size(self, d=None):
        return torch.Size([len(self)]) if d is None else len(self)

    def contiguous(self):
        return self

    @property
    def data(self):
        return self

    @property
    def is_cuda(self):
        return self._buffer.is_cuda

    def cpu(self):
        return self.tensor().cpu()

    def __getitem__(self, i):
        return self.tensor()[i]


# ---------------------------------------------------------------------------------------------
# fused-path primitives (native pixel-major layout, batch-capable, no host sync)
# ---------------------------------------------------------------------------------------------

def _aux_args(aux, state):
    """aux: None | ('tf32', lo_view) | ('bf16', hi16_buf, lo16_buf) -> (mode, hi_ptr, lo_ptr)."""
    if aux is None:
        return _lib.AUX_NONE, None, None
    if aux[0] == 'tf32':
        assert aux[1].stride() == state.stride() and aux[1].dtype == state.dtype
        return _lib.AUX_TF32_LO, None, aux[1].data_ptr()
    assert aux[0] == 'bf16' and aux[1].dtype == torch.bfloat16 and aux[2].dtype == torch.bfloat16
    return _lib.AUX_BF16_PAIR, aux[1].data_ptr(), aux[2].data_ptr()


def detect(x, state, raw_bits, threshold, update_mode, aux=None):
    """cb_change_detect: raw (un-dilated) change bitmap of x vs state; updates state and, when
    given, the auxiliary operand planes `aux` (see _aux_args)."""
    require_cuda(x, state, raw_bits)
    B, Cc, H, W = x.shape
    assert state.shape == x.shape and state.dtype == x.dtype
    mode, hi, lo = _aux_args(aux, state)
    check(C.cb_change_detect(stream_ptr(x.device), dtype_code(x), x.data_ptr(), *_strides4(x),
                             state.data_ptr(), *_strides4(state), mode, hi, lo,
                             raw_bits.data_ptr(), B, Cc, H, W, float(threshold), int(update_mode)))


def detect_u8(x, state, raw_bits, threshold, update_mode, divisor, bias, aux=None):
    """cb_change_detect_u8: detection on a uint8 frame normalised on the fly as u8/divisor + bias
    (the readers' /255 resp. /256 - 0.5); `state` is the fp32 pixel-major state.  Bit-identical to
    :func:`detect` on ``x.float() / divisor + bias``."""
    require_cuda(x, state, raw_bits)
    B, Cc, H, W = x.shape
    assert x.dtype == torch.uint8 and state.dtype == torch.float32 and state.shape == x.shape
    mode, hi, lo = _aux_args(aux, state)
    check(C.cb_change_detect_u8(stream_ptr(x.device), x.data_ptr(), *_strides4(x),
                                state.data_ptr(), *_strides4(state), mode, hi, lo,
                                raw_bits.data_ptr(), B, Cc, H, W, float(divisor), float(bias),
                                float(threshold), int(update_mode)))


def detect_sparse(x, state, raw_bits, threshold, update_mode, candidates, aux=None,
                  bits_are_clear=False):
    """cb_change_detect_sparse: the detection test at the candidate pixels only (see the header
    for the exactness conditions).  `candidates` is a :class:`ChangeIndexes` at x's resolution."""
    require_cuda(x, state, raw_bits)
    B, Cc, H, W = x.shape
    assert state.shape == x.shape and state.dtype == x.dtype
    assert tuple(candidates.shape) == (B, H, W)
    mode, hi, lo = _aux_args(aux, state)
    check(C.cb_change_detect_sparse(stream_ptr(x.device), dtype_code(x), x.data_ptr(), *_strides4(x),
                                    state.data_ptr(), *_strides4(state), mode, hi, lo,
                                    candidates.buffer.data_ptr(), candidates.count.data_ptr(),
                                    raw_bits.data_ptr(), B, Cc, H, W, float(threshold),
                                    int(update_mode), int(bool(bits_are_clear))))


def detect_sparse_compact(x, state, raw_bits, threshold, update_mode, candidates, filtSize, idx, count, sync,
                          aux=None, bits_are_clear=False, dil_bits=None, clear_raw=False):
    """cb_change_detect_sparse_compact: candidate detection + dilation + ordered compaction in one launch
    (small maps: at most cb_compact_small_max_words() bitmap words).  `sync` = int32[1], zeroed once."""
    require_cuda(x, state, raw_bits)
    B, Cc, H, W = x.shape
    assert state.shape == x.shape and state.dtype == x.dtype
    assert tuple(candidates.shape) == (B, H, W)
    mode, hi, lo = _aux_args(aux, state)
    check(C.cb_change_detect_sparse_compact(
        stream_ptr(x.device), dtype_code(x), x.data_ptr(), *_strides4(x), state.data_ptr(), *_strides4(state),
        mode, hi, lo, candidates.buffer.data_ptr(), candidates.count.data_ptr(), raw_bits.data_ptr(), B, Cc, H, W,
        float(threshold), int(update_mode), int(bool(bits_are_clear)),
        dil_bits.data_ptr() if dil_bits is not None else None, idx.data_ptr(), count.data_ptr(), sync.data_ptr(),
        (filtSize[0] - 1) // 2, (filtSize[1] - 1) // 2, int(bool(clear_raw))))


def detect_compact_sparse(x, state, threshold, update_mode, candidates, idx, count, ws, aux=None,
                          bits=None):
    """cb_detect_compact_sparse: candidate detection + ordered compaction in one launch (layers
    without dilation, pixel-major tensors)."""
    require_cuda(x, state)
    B, Cc, H, W = x.shape
    assert state.shape == x.shape and tuple(candidates.shape) == (B, H, W)
    mode, hi, lo = _aux_args(aux, state)
    check(C.cb_detect_compact_sparse(stream_ptr(x.device), dtype_code(x), x.data_ptr(), *_strides4(x),
                                     state.data_ptr(), *_strides4(state), mode, hi, lo,
                                     candidates.buffer.data_ptr(), candidates.count.data_ptr(),
                                     idx.data_ptr(), count.data_ptr(),
                                     bits.data_ptr() if bits is not None else None, ws.data_ptr(),
                                     B, Cc, H, W, float(threshold), int(update_mode)))


def pool_compact(in_bits, in_shape, out_shape, idx, count, ws, out_bits=None):
    """cb_pool_compact: change candidates at the 2x2-pooled resolution from an input bitmap."""
    B, H, W = in_shape
    B2, oH, oW = out_shape
    assert B == B2
    check(C.cb_pool_compact(stream_ptr(in_bits.device), in_bits.data_ptr(),
                            out_bits.data_ptr() if out_bits is not None else None,
                            idx.data_ptr(), count.data_ptr(), ws.data_ptr(), B, H, W, oH, oW))


class PrefetchHints(object):
    """L2 prefetch hints of cb_dilate_compact_hinted: pixel-major maps ``[B, tH, tW, pitch]`` (contiguous)
    whose rows at the changed pixels a later kernel will read -- ``(buffer, shift)`` pairs, shift 1 for a
    map at the 2x2-pooled resolution.  Holds the ctypes argument arrays (and the buffers alive)."""
    MAX = 3

    def __init__(self, targets):
        import ctypes
        targets = [(t, int(sh)) for t, sh in targets
                   if t is not None and t.dim() == 4 and t.is_contiguous() and t.data_ptr() % 16 == 0
                   and (t.shape[3] * t.element_size()) % 16 == 0 and t.numel() > 0][:self.MAX]
        self.targets = targets
        self.n = len(targets)
        self.key = tuple((t.data_ptr(), tuple(t.shape), sh) for t, sh in targets)
        self.base = (ctypes.c_void_p * max(self.n, 1))(*[t.data_ptr() for t, _ in targets])
        ia = ctypes.c_int * max(self.n, 1)
        self.row_bytes = ia(*[t.shape[3] * t.element_size() for t, _ in targets])
        self.shift = ia(*[sh for _, sh in targets])
        self.h = ia(*[t.shape[1] for t, _ in targets])
        self.w = ia(*[t.shape[2] for t, _ in targets])


def _dilate_hinted(raw_bits, shape, filtSize, idx, count, ws, dil_bits, dil_map, tile_ws, clear_raw, no_list,
                   hints):
    B, H, W = shape
    check(C.cb_dilate_compact_hinted(stream_ptr(raw_bits.device), raw_bits.data_ptr(),
                                     dil_bits.data_ptr() if dil_bits is not None else None,
                                     dil_map.data_ptr() if dil_map is not None else None,
                                     idx.data_ptr() if idx is not None else None, count.data_ptr(),
                                     ws.data_ptr(), tile_ws.data_ptr() if tile_ws is not None else None,
                                     B, H, W, (filtSize[0] - 1) // 2, (filtSize[1] - 1) // 2,
                                     int(bool(clear_raw)), int(bool(no_list)), hints.n, hints.base,
                                     hints.row_bytes, hints.shift, hints.h, hints.w))


def dilate_tiles(raw_bits, shape, filtSize, count, ws, dil_bits, tile_ws, clear_raw=False, hints=None):
    """cb_dilate_tiles: dilation + dirty-tile list + change count, no ordered index list.  `hints`
    (:class:`PrefetchHints`): cb_dilate_compact_hinted, same outputs."""
    B, H, W = shape
    if hints is not None and hints.n:
        return _dilate_hinted(raw_bits, shape, filtSize, None, count, ws, dil_bits, None, tile_ws, clear_raw,
                              True, hints)
    check(C.cb_dilate_tiles(stream_ptr(raw_bits.device), raw_bits.data_ptr(), dil_bits.data_ptr(),
                            count.data_ptr(), ws.data_ptr(), tile_ws.data_ptr(), B, H, W,
                            (filtSize[0] - 1) // 2, (filtSize[1] - 1) // 2, int(bool(clear_raw))))


def dilate_compact(raw_bits, shape, filtSize, idx, count, ws, dil_bits=None, dil_map=None,
                   clear_raw=False, tile_ws=None, hints=None):
    """cb_dilate_compact: dilation by the filter footprint + ordered compaction.  With `tile_ws`
    (cb_tile_ws_bytes of zeroed int32 device memory) the dirty 8x16 output tiles are listed as well
    (cb_dilate_compact_tiles), for :func:`conv_update_tiled`.  `hints` (:class:`PrefetchHints`):
    cb_dilate_compact_hinted, same outputs."""
    B, H, W = shape
    if hints is not None and hints.n:
        return _dilate_hinted(raw_bits, shape, filtSize, idx, count, ws, dil_bits, dil_map, tile_ws, clear_raw,
                              False, hints)
    if tile_ws is not None:
        check(C.cb_dilate_compact_tiles(stream_ptr(raw_bits.device), raw_bits.data_ptr(),
                                        dil_bits.data_ptr() if dil_bits is not None else None,
                                        dil_map.data_ptr() if dil_map is not None else None,
                                        idx.data_ptr(), count.data_ptr(), ws.data_ptr(),
                                        tile_ws.data_ptr(), B, H, W, (filtSize[0] - 1) // 2,
                                        (filtSize[1] - 1) // 2, int(bool(clear_raw))))
        return
    check(C.cb_dilate_compact(stream_ptr(raw_bits.device), raw_bits.data_ptr(),
                              dil_bits.data_ptr() if dil_bits is not None else None,
                              dil_map.data_ptr() if dil_map is not None else None,
                              idx.data_ptr(), count.data_ptr(), ws.data_ptr(), B, H, W,
                              (filtSize[0] - 1) // 2, (filtSize[1] - 1) // 2, int(bool(clear_raw))))


def alloc_tile_ws(shape, device):
    """tile workspace of cb_dilate_compact_tiles / cb_conv_update_tiled for a [B,H,W] pixel grid."""
    B, H, W = shape
    return torch.zeros(C.cb_tile_ws_bytes(B, H, W) // 4, dtype=torch.int32, device=device)


def tiled_supported(dtype, gemm, shape, Cin, Cout, filtSize):
    """cb_conv_tiled_supported: 1 = run the tile path, 2 = possible but not recommended, 0 = no."""
    B, H, W = shape
    return C.cb_conv_tiled_supported(dtype_code(dtype), gemm, B, H, W, Cin, Cout, filtSize[0], filtSize[1])


def alloc_scratch(shape, device, want_map=False):
    """bitmaps / index buffer / count / compaction workspace for a [B,H,W] pixel grid."""
    B, H, W = shape
    nwords = C.cb_bitmap_words(B, H, W)
    s = dict(
        raw_bits=torch.zeros(max(nwords, 1), dtype=torch.int32, device=device),
        dil_bits=torch.zeros(max(nwords, 1), dtype=torch.int32, device=device),
        idx=torch.zeros(max(B * H * W, 1), dtype=torch.int32, device=device),
        count=torch.zeros(1, dtype=torch.int32, device=device),
        ws=torch.zeros(C.cb_compact_ws_bytes(B, H, W), dtype=torch.uint8, device=device),
    )
    if want_map:
        s["dil_map"] = torch.zeros(B, H, W, dtype=torch.int8, device=device)
    return s


def pack_weights(weight, gemm):
    """cb_pack_weights: [Cout,Cin,kH,kW] -> the layout cb_conv_update expects for `gemm`."""
    require_cuda(weight)
    Cout, Cin, kH, kW = weight.shape
    w = weight.detach().contiguous()
    nbytes = C.cb_packed_weight_bytes(dtype_code(w), gemm, Cout, Cin, kH, kW)
    packed = torch.empty(nbytes + 256, dtype=torch.uint8, device=w.device)
    off = (-packed.data_ptr()) % 256                       # TMA wants >=128-byte alignment
    packed = packed[off: off + nbytes]
    check(C.cb_pack_weights(stream_ptr(w.device), dtype_code(w), gemm, w.data_ptr(),
                            packed.data_ptr(), Cout, Cin, kH, kW))
    return packed


def tf32_lo(x):
    """v - trunc_tf32(v): the remainder plane the 3xTF32 contraction consumes (exact in fp32)."""
    hi = (x.contiguous().view(torch.int32) & -8192).view(torch.float32)
    return x - hi


def bf16_pair(x):
    """(bf16(v), bf16(v - bf16(v))): the operand planes of the 3xBF16 contraction."""
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


def bf16_planes(state_buf, C_):
    """pixel-major bf16 hi/lo planes [B,H,W,pitch16] of a pixel-major fp32 buffer."""
    B, H, W, _ = state_buf.shape
    p16 = C.cb_plane_pitch16(C_)
    hi = torch.zeros(B, H, W, p16, dtype=torch.bfloat16, device=state_buf.device)
    lo = torch.zeros_like(hi)
    h, l = bf16_pair(state_buf[..., :C_])
    hi[..., :C_] = h
    lo[..., :C_] = l
    return hi, lo


class ConvWorkspace(object):
    """Stream-K workspace of cb_conv_update (cb_conv_ws_bytes() of zeroed device memory; the kernel
    leaves it clean).  Launches that may overlap need their own: one holder per model instance,
    shared by its layers (a model's layers run in order).  Not pickled."""

    def __init__(self):
        self.buf = None

    def get(self, device):
        if os.environ.get("CBINFER_STREAMK", "1") == "0":
            return None
        if self.buf is None or self.buf.device != device:
            self.buf = torch.zeros(C.cb_conv_ws_bytes(), dtype=torch.uint8, device=device)
        return self.buf

    def clear(self):
        self.buf = None

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.buf = None


def conv_update(state_buf, changes, packed_w, bias_f32, out_buf, Cin, Cout, filtSize, relu, gemm,
                lo_buf=None, planes16=None, ws=None, mask=None):
    """cb_conv_update on pixel-major buffers [B,H,W,pitch].  GEMM_TC_3X (fp32) consumes the tf32
    remainder plane `lo_buf`, GEMM_TC_BF16X3 the bf16 hi/lo planes `planes16`; both are derived
    on the fly when the caller does not maintain them (cb_change_detect can)."""
    B, H, W, Cp = state_buf.shape
    assert out_buf.shape[:3] == state_buf.shape[:3]
    src, src_lo, pitch = state_buf, lo_buf, Cp
    if gemm == _lib.GEMM_TC_BF16X3:
        assert state_buf.dtype == torch.float32
        if planes16 is None:
            planes16 = bf16_planes(state_buf, Cin)
        src, src_lo = planes16
        pitch = src.shape[3]
    elif gemm == _lib.GEMM_TC_3X and state_buf.dtype == torch.float32 and lo_buf is None:
        src_lo = tf32_lo(state_buf)
    if mask is not None:
        # `changes` is a superset list; mask = dict(bits=raw bitmap, clear=bool, count=int32[1] out,
        # sync=int32[2] zeroed once): only pixels whose bit is set are processed (cb_conv_update_masked)
        check(C.cb_conv_update_masked(stream_ptr(state_buf.device), dtype_code(state_buf), gemm,
                                      src.data_ptr(), src_lo.data_ptr() if src_lo is not None else None,
                                      pitch, changes.buffer.data_ptr(), changes.count.data_ptr(),
                                      packed_w.data_ptr(), bias_f32.data_ptr(), out_buf.data_ptr(),
                                      out_buf.shape[3], B, H, W, Cin, Cout, filtSize[0], filtSize[1],
                                      int(bool(relu)), ws.data_ptr() if ws is not None else None,
                                      ws.numel() if ws is not None else 0,
                                      mask['bits'].data_ptr(), int(bool(mask.get('clear', False))),
                                      mask['count'].data_ptr(), mask['sync'].data_ptr()))
        return
    check(C.cb_conv_update(stream_ptr(state_buf.device), dtype_code(state_buf), gemm,
                           src.data_ptr(), src_lo.data_ptr() if src_lo is not None else None,
                           pitch, changes.buffer.data_ptr(),
                           changes.count.data_ptr(), packed_w.data_ptr(), bias_f32.data_ptr(),
                           out_buf.data_ptr(), out_buf.shape[3], B, H, W, Cin, Cout,
                           filtSize[0], filtSize[1], int(bool(relu)),
                           ws.data_ptr() if ws is not None else None,
                           ws.numel() if ws is not None else 0))


def fg_detect(x, prev_view, prev_buf, planes16, raw_bits, threshold, count=None):
    """cb_fg_detect: per-VALUE thresholded delta of x against the pixel-major fp32 state (`prev_view` =
    the [B,C,H,W] view of `prev_buf` [B,H,W,pitch]) -> bf16 hi/lo delta planes `planes16`, raw pixel
    bitmap, changed-value count; the state is overwritten with x."""
    require_cuda(x, prev_buf, raw_bits)
    B, Cc, H, W = x.shape
    assert x.dtype == torch.float32 and prev_buf.dtype == torch.float32
    assert tuple(prev_view.shape) == tuple(x.shape) and prev_view.stride(1) == 1
    hi, lo = planes16
    assert hi.dtype == torch.bfloat16 and hi.shape == (B, H, W, C.cb_plane_pitch16(Cc)) and lo.shape == hi.shape
    check(C.cb_fg_detect(stream_ptr(x.device), x.data_ptr(), *_strides4(x), prev_view.data_ptr(),
                         prev_view.stride(0), prev_view.stride(2), prev_view.stride(3),
                         hi.data_ptr(), lo.data_ptr(), raw_bits.data_ptr(),
                         count.data_ptr() if count is not None else None, B, Cc, H, W, float(threshold)))


def conv_accumulate(planes16, changes, packed_w, out_buf, Cin, Cout, filtSize, gemm, ws=None):
    """cb_conv_accumulate: out_buf[pix, :] += contraction of the operand planes around pix with the
    packed weights, for the listed pixels (no bias, no ReLU) - the fine-grained update."""
    src, src_lo = planes16
    B, H, W, pitch = src.shape
    assert gemm == _lib.GEMM_TC_BF16X3 and out_buf.dtype == torch.float32
    assert out_buf.shape[:3] == src.shape[:3]
    check(C.cb_conv_accumulate(stream_ptr(src.device), _lib.F32, gemm, src.data_ptr(), src_lo.data_ptr(),
                               pitch, changes.buffer.data_ptr(), changes.count.data_ptr(),
                               packed_w.data_ptr(), out_buf.data_ptr(), out_buf.shape[3], B, H, W, Cin,
                               Cout, filtSize[0], filtSize[1],
                               ws.data_ptr() if ws is not None else None,
                               ws.numel() if ws is not None else 0))


def conv_update_tiled(state_buf, tile_ws, dil_bits, packed_w, bias_f32, out_buf, Cin, Cout, filtSize,
                      relu, gemm, lo_buf=None, planes16=None, pool=None, self_list=None):
    """cb_conv_update_tiled on pixel-major buffers: the contraction over the dirty 8x16 tiles listed
    in `tile_ws` (TMA-staged halo tiles, implicit im2col), writing the pixels set in `dil_bits`.
    Operand conventions as :func:`conv_update`.  `self_list` = dict(raw_bits, count, ws, clear_raw):
    cb_conv_update_tiled_self -- the kernel dilates the raw bitmap and lists the tiles itself (no
    cb_dilate_tiles launch before it); dil_bits, tile_ws and count are outputs then."""
    B, H, W, Cp = state_buf.shape
    assert out_buf.shape[:3] == state_buf.shape[:3]
    src, src_lo, pitch = state_buf, lo_buf, Cp
    if gemm == _lib.GEMM_TC_BF16X3:
        assert state_buf.dtype == torch.float32
        if planes16 is None:
            planes16 = bf16_planes(state_buf, Cin)
        src, src_lo = planes16
        pitch = src.shape[3]
    elif gemm == _lib.GEMM_TC_3X and state_buf.dtype == torch.float32 and lo_buf is None:
        src_lo = tf32_lo(state_buf)
    if self_list is not None:
        pa = [None, 0, 0, 0, 0, 0, None, 0, 0, 0, 0, None, None, None, 0.0, 0]
        if pool is not None:
            po, ns = pool["out"], pool["next_state"]
            assert po.stride(1) == 1 and ns.stride(1) == 1 and po.shape == ns.shape, "pixel-major maps only"
            mode, hi, lo = _aux_args(pool.get("aux"), ns)
            pa = [po.data_ptr(), po.stride(0), po.stride(2), po.stride(3), po.size(2), po.size(3),
                  ns.data_ptr(), ns.stride(0), ns.stride(2), ns.stride(3), mode, hi, lo,
                  pool["next_raw_bits"].data_ptr(), float(pool["threshold"]), int(pool["mode"])]
        check(C.cb_conv_update_tiled_self(stream_ptr(state_buf.device), dtype_code(state_buf), gemm,
                                          src.data_ptr(), src_lo.data_ptr() if src_lo is not None else None,
                                          pitch, tile_ws.data_ptr(), dil_bits.data_ptr(), packed_w.data_ptr(),
                                          bias_f32.data_ptr(), out_buf.data_ptr(), out_buf.shape[3], B, H, W,
                                          Cin, Cout, filtSize[0], filtSize[1], int(bool(relu)), *pa,
                                          self_list["raw_bits"].data_ptr(), self_list["count"].data_ptr(),
                                          self_list["ws"].data_ptr(), int(bool(self_list.get("clear_raw")))))
        return
    if pool is not None:
        # pool = dict(out=pooled map view, next_state=view, next_raw_bits, threshold, mode, aux): the
        # epilogue also re-pools the touched 2x2 windows and runs the next layer's detection on them
        po, ns = pool["out"], pool["next_state"]
        assert po.stride(1) == 1 and ns.stride(1) == 1 and po.shape == ns.shape, "pixel-major maps only"
        mode, hi, lo = _aux_args(pool.get("aux"), ns)
        check(C.cb_conv_update_tiled_pool(stream_ptr(state_buf.device), dtype_code(state_buf), gemm,
                                          src.data_ptr(), src_lo.data_ptr() if src_lo is not None else None,
                                          pitch, tile_ws.data_ptr(), dil_bits.data_ptr(), packed_w.data_ptr(),
                                          bias_f32.data_ptr(), out_buf.data_ptr(), out_buf.shape[3], B, H, W,
                                          Cin, Cout, filtSize[0], filtSize[1], int(bool(relu)),
                                          po.data_ptr(), po.stride(0), po.stride(2), po.stride(3),
                                          po.size(2), po.size(3), ns.data_ptr(), ns.stride(0), ns.stride(2),
                                          ns.stride(3), mode, hi, lo, pool["next_raw_bits"].data_ptr(),
                                          float(pool["threshold"]), int(pool["mode"])))
        return
    check(C.cb_conv_update_tiled(stream_ptr(state_buf.device), dtype_code(state_buf), gemm,
                                 src.data_ptr(), src_lo.data_ptr() if src_lo is not None else None,
                                 pitch, tile_ws.data_ptr(), dil_bits.data_ptr(), packed_w.data_ptr(),
                                 bias_f32.data_ptr(), out_buf.data_ptr(), out_buf.shape[3], B, H, W,
                                 Cin, Cout, filtSize[0], filtSize[1], int(bool(relu))))


def tail_update(x_buf, state1_buf, packed1, bias1, out1_buf, relu1, thr1, state2_buf, packed2, bias2,
                out2_buf, relu2, thr2, candidates, C0, C1, C2, update_mode, count1, count2, sync):
    """cb_tail_update: two chained 1x1 change-based layers (detect, contraction, detect,
    contraction) for the candidate pixels in one launch; pixel-major fp32 buffers [B,H,W,pitch]."""
    assert x_buf.dtype == torch.float32 and x_buf.shape[3] == C0 and state1_buf.shape == x_buf.shape
    assert out1_buf.shape[3] == C1 and state2_buf.shape == out1_buf.shape
    check(C.cb_tail_update(stream_ptr(x_buf.device), x_buf.data_ptr(), state1_buf.data_ptr(),
                           packed1.data_ptr(), bias1.data_ptr(), out1_buf.data_ptr(), int(bool(relu1)),
                           float(thr1), state2_buf.data_ptr(), packed2.data_ptr(), bias2.data_ptr(),
                           out2_buf.data_ptr(), out2_buf.shape[3], int(bool(relu2)), float(thr2),
                           candidates.buffer.data_ptr(), candidates.count.data_ptr(), C0, C1, C2,
                           int(update_mode), count1.data_ptr(), count2.data_ptr(), sync.data_ptr()))


def pixel_major(shape, dtype, device, fill):
    """Allocate a [B,C,H,W]-shaped *view* over a pixel-major [B,H,W,pitch] buffer whose pitch is
    C rounded up to 16 bytes.  Real channels are set to `fill`, pad channels to 0 (they meet zero
    weights in the contraction).  Returns (view, buffer)."""
    B, Cc, H, W = shape
    pitch = _lib.channel_pitch(dtype, Cc)
    buf = torch.zeros(B, H, W, pitch, dtype=dtype, device=device)
    view = buf[..., :Cc].permute(0, 3, 1, 2)
    if fill != 0:
        view.fill_(fill)
    return view, buf


# ---------------------------------------------------------------------------------------------
# reference-named ops (planar layout, batch 1 unless noted)
# ---------------------------------------------------------------------------------------------

def changeDetection(input, prevInput, filtSize, threshold, updateInputState=False):
    """conv2d_cg.py:100-122: int8 change map [H,W] (batch 1) / [B,H,W], dilated by filtSize;
    with updateInputState the changed pixels of prevInput are overwritten in place."""
    assert input.size() == prevInput.size() and input.dim() == 4
    require_cuda(input, prevInput)
    B, _, H, W = input.shape
    s = alloc_scratch((B, H, W), input.device, want_map=True)
    detect(input, prevInput, s["raw_bits"], threshold,
           _lib.UPDATE_CHANGED if updateInputState else _lib.UPDATE_NONE)
    dilate_compact(s["raw_bits"], (B, H, W), filtSize, s["idx"], s["count"], s["ws"],
                   dil_map=s["dil_map"])
    return s["dil_map"][0] if B == 1 else s["dil_map"]


def _map_to_bits(changeMap):
    H, W = changeMap.shape[-2:]
    B = changeMap.numel() // (H * W)
    m = changeMap.reshape(B, H, W).to(torch.int8).contiguous()
    require_cuda(m)
    bits = torch.zeros(max(C.cb_bitmap_words(B, H, W), 1), dtype=torch.int32, device=m.device)
    check(C.cb_map_to_bits(stream_ptr(m.device), m.data_ptr(), bits.data_ptr(), B, H, W))
    return bits, (B, H, W)


def changePropagation(changeMap, filtSize):
    """conv2d_cg.py:159-177: gather-dilate an int8/bool change map by filtSize."""
    assert len(filtSize) == 2
    if filtSize[0] == 1 and filtSize[1] == 1:
        return changeMap
    bits, shape = _map_to_bits(changeMap)
    s = alloc_scratch(shape, changeMap.device, want_map=True)
    dilate_compact(bits, shape, filtSize, s["idx"], s["count"], s["ws"], dil_map=s["dil_map"])
    return s["dil_map"].reshape(changeMap.shape).to(changeMap.dtype)


def changeIndexesExtr(changeMap, lazy=False):
    """conv2d_cg.py:200-213: ascending int32 indices of the non-zero map cells.  With
    ``lazy=True`` returns a :class:`ChangeIndexes` (no host sync)."""
    bits, shape = _map_to_bits(changeMap)
    s = alloc_scratch(shape, changeMap.device)
    dilate_compact(bits, shape, (1, 1), s["idx"], s["count"], s["ws"])
    ci = ChangeIndexes(s["idx"], s["count"], shape)
    return ci if lazy else ci.tensor()


changeIndexesExtr_python = changeIndexesExtr     # the reference's forward calls the *_python name


def genXMatrix(input, changeIndexes, filtSize):
    """conv2d_cg.py:239-261: im2col of the changed pixels, X[n, Cin*kH*kW] (planar input)."""
    require_cuda(input)
    inC, inH, inW = input.size(-3), input.size(-2), input.size(-1)
    kH, kW = filtSize
    idx = changeIndexes.tensor() if isinstance(changeIndexes, ChangeIndexes) else changeIndexes
    idx = idx.to(torch.int32).contiguous()
    n = idx.numel()
    X = input.new_empty(n, inC * kH * kW)
    if n > 0:
        inp = input.contiguous()
        check(C.cb_gen_xmatrix(stream_ptr(input.device), dtype_code(input), X.data_ptr(),
                               inp.data_ptr(), idx.data_ptr(), kW, kH, inC, inW, inH, n))
    return X


def matrixMult(Xmatrix, weights, bias, activFun=None):
    """conv2d_cg.py:342-349: Y = X . W.view(Cout,-1)^T + bias (fp32 accumulation)."""
    if Xmatrix.numel() == 0:
        return Xmatrix.clone()
    require_cuda(Xmatrix, weights, bias)
    n, K = Xmatrix.shape
    Cout = weights.size(0)
    Y = Xmatrix.new_empty(n, Cout)
    w = weights.detach().contiguous().view(Cout, -1)
    assert w.size(1) == K
    check(C.cb_matrix_mult(stream_ptr(Xmatrix.device), dtype_code(Xmatrix),
                           Xmatrix.contiguous().data_ptr(), w.data_ptr(),
                           bias.detach().contiguous().data_ptr(), Y.data_ptr(), n, K, Cout))
    if activFun is not None:
        Y = activFun(Y)
    return Y


matrixMult_python = matrixMult


def updateOutput(YMatrix, changeIndexes, prevOutput, withReLU=False):
    """conv2d_cg.py:292-313: scatter Y^T[Cout,n] into the planar prevOutput (in place)."""
    require_cuda(YMatrix, prevOutput)
    outC, outH, outW = prevOutput.size(-3), prevOutput.size(-2), prevOutput.size(-1)
    idx = changeIndexes.tensor() if isinstance(changeIndexes, ChangeIndexes) else changeIndexes
    idx = idx.to(torch.int32).contiguous()
    n = idx.numel()
    if n > 0:
        assert prevOutput.is_contiguous()
        Yt = YMatrix.contiguous()
        check(C.cb_update_output(stream_ptr(prevOutput.device), dtype_code(prevOutput),
                                 Yt.data_ptr(), prevOutput.data_ptr(), idx.data_ptr(),
                                 outW * outH, n, outC, int(bool(withReLU))))
    return prevOutput


def maxPool2d(input, outputState, changeIndexes, kernelSize=(2, 2), stride=(2, 2)):
    """conv2d_cg.py:58-82: recompute the 2x2/stride-2 windows touched by the changed input
    pixels into outputState (in place).  Any strides, batch >= 1."""
    assert tuple(kernelSize) == (2, 2) and tuple(stride) == (2, 2)
    require_cuda(input, outputState)
    assert input.dim() == 4 and outputState.dim() == 4
    B, Cc, H, W = input.shape
    oH, oW = outputState.size(-2), outputState.size(-1)
    if not isinstance(changeIndexes, ChangeIndexes):
        changeIndexes = ChangeIndexes.from_tensor(changeIndexes, (B, H, W))
    bits = changeIndexes.bits
    check(C.cb_maxpool2x2(stream_ptr(input.device), dtype_code(input), input.data_ptr(),
                          *_strides4(input), changeIndexes.buffer.data_ptr(),
                          changeIndexes.count.data_ptr(),
                          bits.data_ptr() if bits is not None else None,
                          outputState.data_ptr(), *_strides4(outputState), B, Cc, H, W, oH, oW))
    return outputState


def maxPool2d_detect(input, outputState, changeIndexes, next_state, next_raw_bits, threshold,
                     update_mode, aux=None):
    """cb_maxpool2x2_detect: change-based pooling fused with the next layer's detection (pixel-major
    tensors only; `changeIndexes.bits` required; `next_raw_bits` must be clear)."""
    require_cuda(input, outputState, next_state)
    B, Cc, H, W = input.shape
    oH, oW = outputState.size(-2), outputState.size(-1)
    assert changeIndexes.bits is not None and next_state.shape == outputState.shape
    for t in (input, outputState, next_state):
        assert t.stride(1) == 1, "pixel-major tensors only"
    mode, hi, lo = _aux_args(aux, next_state)
    check(C.cb_maxpool2x2_detect(stream_ptr(input.device), dtype_code(input), input.data_ptr(),
                                 input.stride(0), input.stride(2), input.stride(3),
                                 changeIndexes.buffer.data_ptr(), changeIndexes.count.data_ptr(),
                                 changeIndexes.bits.data_ptr(), outputState.data_ptr(),
                                 outputState.stride(0), outputState.stride(2), outputState.stride(3),
                                 B, Cc, H, W, oH, oW, next_state.data_ptr(), next_state.stride(0),
                                 next_state.stride(2), next_state.stride(3), mode, hi, lo,
                                 next_raw_bits.data_ptr(), float(threshold), int(update_mode)))
    return outputState
