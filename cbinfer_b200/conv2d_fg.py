"""Fine-grained change-based convolution: python wrapper over cb_fg_update.

Mirrors the reference's ``pycbinfer/conv2d_fg.py``: ``cbconvFG(input, prevInput, output, weight,
threshold)`` (conv2d_fg.py:75-96) pushes W*delta for every input *value* whose change exceeds the
threshold into ``output`` (planar NCHW fp32).  The reference's three steps -- changeDetectionFG
(:34-46), torch.nonzero (:82, host sync) and updateOutputFG (:48-72) -- are one kernel here, and
the reference's CPU branch (conv2d_fg_cpu) is not built: CUDA only.

Note the library also performs the ``prevInput = input`` of conv2d.py:175 (prevInput is updated in
place); pass ``updatePrev=False`` semantics by handing in a clone if that is not wanted.
"""
import torch

from ._lib import C, check, require_cuda, stream_ptr


def cbconvFG(input, prevInput, output, weight, threshold, count=None):
    require_cuda(input, prevInput, output, weight)
    assert input.dtype == torch.float32, "the fine-grained path is fp32 only (as in the reference)"
    assert input.is_contiguous() and prevInput.is_contiguous() and output.is_contiguous()
    assert weight.dim() == 4
    B, Cin, H, W = input.shape
    Cout, Cin2, kH, kW = weight.shape
    assert Cin2 == Cin and output.shape == (B, Cout, H, W) and prevInput.shape == input.shape
    if count is None:
        count = torch.zeros(1, dtype=torch.int32, device=input.device)
    check(C.cb_fg_update(stream_ptr(input.device), input.data_ptr(), prevInput.data_ptr(),
                         weight.detach().contiguous().data_ptr(), output.data_ptr(),
                         count.data_ptr(), B, Cin, Cout, H, W, kH, kW, float(threshold)))
    return output
