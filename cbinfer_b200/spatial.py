"""Spatial split of very large frames (e.g. 3840x2160) over the GPUs of one box: row bands with a
one-off input halo exchange and a gather of the result bands.

Each rank owns a band of image rows (aligned to the network's total stride) and runs the UNCHANGED
change-based model on its band plus `halo` rows taken from each neighbour; the halo covers the
receptive-field radius of the whole network, so every output row of the band's interior is computed
from exactly the inputs the full-frame model would use, and the rows near the slab edges (whose
receptive fields stick out of the slab) are dropped.  For the scene-labeling net: radius
3 + 2*3 + 4*3 = 21 input rows, rounded to 24 to keep the two 2x2 pools aligned.
Communication per frame: 2 * halo input rows to/from the neighbours (NCCL send/recv over NVLink)
and one all_gather of the output bands; no per-layer collective.  The reference has nothing
comparable (single GPU, and its int32 im2col offsets overflow at 4K,
pycbinfer/cbconv2d_cg_backend.cu:154).

All change-based state (prevInput, prevOutput, pooled maps) is per rank and covers the slab only.
"""
import torch
import torch.distributed as dist


def band_rows(H, world, rank, align=4):
    """[lo, hi) input rows owned by `rank`: contiguous bands, boundaries multiples of `align`."""
    units = H // align
    lo = (units * rank) // world * align
    hi = (units * (rank + 1)) // world * align if rank + 1 < world else H
    return lo, hi


def slab_rows(H, world, rank, halo, align=4):
    """[lo, hi) input rows of the slab (band + halo, clipped to the image)."""
    lo, hi = band_rows(H, world, rank, align)
    return max(0, lo - halo), min(H, hi + halo)


class SpatialSplit(object):
    """Run `model` (any module mapping [B,C,H,W] -> [B,C',H/stride,W/stride], e.g. a converted
    CBinfer model) on this rank's slab and assemble the full-resolution result.

    `stride` = total down-sampling of the model (4 for the scene net), `halo` = input rows taken
    from each neighbour (>= receptive-field radius, multiple of `stride`)."""

    def __init__(self, model, H, world=None, rank=None, halo=24, stride=4, group=None):
        self.model = model
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        assert halo % stride == 0 and H % stride == 0
        self.H, self.halo, self.stride = H, halo, stride
        self.band = band_rows(H, self.world, self.rank, stride)
        self.slab = slab_rows(H, self.world, self.rank, halo, stride)

    # ---- input side ------------------------------------------------------------------------
    def exchange_halo(self, band_frame):
        """band_frame: this rank's own rows [B,C,band,W] (e.g. its part of the camera feed).
        Returns the slab [B,C,slab,W] after receiving `halo` rows from each existing neighbour."""
        lo, hi = self.band
        slo, shi = self.slab
        top, bot = lo - slo, shi - hi
        B, C, _, W = band_frame.shape
        slab = band_frame.new_empty(B, C, shi - slo, W)
        slab[:, :, top:top + (hi - lo)] = band_frame
        ops, bufs = [], []
        if self.world > 1:
            if self.rank > 0:                     # rows for / from the upper neighbour
                send_up = band_frame[:, :, :self.halo].contiguous()
                recv_up = band_frame.new_empty(B, C, top, W)
                ops += [dist.P2POp(dist.isend, send_up, self.rank - 1, self.group),
                        dist.P2POp(dist.irecv, recv_up, self.rank - 1, self.group)]
                bufs.append(("top", recv_up))
            if self.rank + 1 < self.world:
                send_dn = band_frame[:, :, -self.halo:].contiguous()
                recv_dn = band_frame.new_empty(B, C, bot, W)
                ops += [dist.P2POp(dist.isend, send_dn, self.rank + 1, self.group),
                        dist.P2POp(dist.irecv, recv_dn, self.rank + 1, self.group)]
                bufs.append(("bot", recv_dn))
            if ops:
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
        for side, buf in bufs:
            if side == "top":
                slab[:, :, :top] = buf
            else:
                slab[:, :, top + (hi - lo):] = buf
        return slab

    # ---- compute + output side ---------------------------------------------------------------
    def forward_slab(self, slab):
        """Run the model on the slab and crop to the band's output rows [B,C',band/stride,W']."""
        out = self.model(slab)
        if isinstance(out, tuple) and out and out[0] == 'changeIndexes':
            out = out[1]
        lo, hi = self.band
        slo, _ = self.slab
        o0 = (lo - slo) // self.stride
        return out[:, :, o0:o0 + (hi - lo) // self.stride]

    def gather(self, band_out):
        """all_gather the output bands -> full [B,C',H/stride,W'] on every rank."""
        if self.world == 1:
            return band_out
        sizes = [(band_rows(self.H, self.world, r, self.stride)) for r in range(self.world)]
        parts = [band_out.new_empty(band_out.shape[0], band_out.shape[1], (h - l) // self.stride,
                                    band_out.shape[3]) for (l, h) in sizes]
        dist.all_gather(parts, band_out.contiguous(), group=self.group)
        return torch.cat(parts, dim=2)

    def __call__(self, band_frame):
        return self.gather(self.forward_slab(self.exchange_halo(band_frame)))
