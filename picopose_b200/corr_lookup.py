"""Stage-3 lookup: host side of the reference's ``utils/corr_lookup.py`` on libpicopose_b200.

Same public names and signatures as the reference module:
``coords_grid`` (:9-26), ``bilinear_sample`` (:29-65), ``CorrLookup`` (:69-134).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib


def coords_grid(batch: int, xx: Tensor, yy: Tensor) -> Tensor:
    """(batch, 2, H, W) float grid: channel 0 repeats `xx` along rows, channel 1 repeats `yy` along columns."""
    H, W = yy.shape[0], xx.shape[0]
    gx = xx.float().view(1, W).expand(H, W)
    gy = yy.float().view(H, 1).expand(H, W)
    return torch.stack([gx, gy], dim=0)[None].repeat(batch, 1, 1, 1)


def bilinear_sample(feat: Tensor, grid: Tensor, mode: str = 'bilinear', padding_mode: str = 'zeros',
                    align_corners: bool = False, scale: bool = True) -> Tensor:
    """Samples `feat` (N,C,Hf,Wf) at `grid` ((N,Ho,Wo,2) or (N,2,Ho,Wo); pixel coordinates when `scale`).

    Unlike the reference this does not scale the caller's `grid` in place (its callers pass temporaries).
    """
    if mode != 'bilinear' or padding_mode != 'zeros':
        raise NotImplementedError(
            f"picopose_b200.bilinear_sample implements mode='bilinear', padding_mode='zeros' "
            f"(the only combination PicoPose uses); got mode='{mode}', padding_mode='{padding_mode}'")
    _lib.require_cuda(feat, grid)
    _lib.require_inference("bilinear_sample", feat, grid)
    lib = _lib.load()
    feat = feat.float().contiguous()
    grid = grid.float().contiguous()
    N, Cc, Hf, Wf = feat.shape
    chw = grid.shape[-1] != 2
    Ho, Wo = (grid.shape[2], grid.shape[3]) if chw else (grid.shape[1], grid.shape[2])
    out = torch.empty(N, Cc, Ho, Wo, dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _lib.check(lib.pp_bilinear_sample(_lib.ptr(feat), _lib.ptr(grid), N, Cc, Hf, Wf, Ho, Wo, int(chw),
                                          int(bool(align_corners)), int(bool(scale)), _lib.ptr(out),
                                          _lib.stream_of(feat)), "pp_bilinear_sample")
    return out


def corr_lookup(corr_pyramid: Sequence[Tensor], flow: Tensor, radius: int) -> Tensor:
    """Functional form of CorrLookup.forward -> (B, L*(2r+1)^2, H, W) fp32.  A `TiledPyramid` is read in its own
    layout (radius 1..8), anything else as the reference's list of row-major volumes."""
    from .correlation import TiledPyramid
    tiled = isinstance(corr_pyramid, TiledPyramid)
    if tiled and not 1 <= int(radius) <= 8:
        tiled, corr_pyramid = False, corr_pyramid.rowmajor()
    if tiled:
        corr_pyramid = corr_pyramid.tiled_levels
    _lib.require_cuda(flow, *corr_pyramid)
    _lib.require_inference("CorrLookup", flow, *corr_pyramid)
    lib = _lib.load()
    flow = flow.float().contiguous()
    B, two, H, W = flow.shape
    if two != 2:
        raise ValueError("flow must be (B, 2, H, W)")
    L = len(corr_pyramid)
    vols = []
    for lvl, c in enumerate(corr_pyramid):
        c = c.float().contiguous()
        if c.dim() != 4 or c.shape[0] != B * H * W or c.shape[1] != 1:
            raise ValueError(
                f"pyramid level {lvl} must be (B*H*W, 1, h, w); got {tuple(c.shape)} for flow {tuple(flow.shape)}")
        vols.append(c)
    D = 2 * int(radius) + 1
    out = torch.empty(B, L * D * D, H, W, dtype=torch.float32, device=flow.device)
    ptrs = (C.c_void_p * L)(*[v.data_ptr() for v in vols])
    hs = (C.c_int * L)(*[v.shape[2] for v in vols])
    ws = (C.c_int * L)(*[v.shape[3] for v in vols])
    fn = lib.pp_corr_lookup_tiled if tiled else lib.pp_corr_lookup
    with torch.cuda.device(flow.device):
        _lib.check(fn(ptrs, hs, ws, L, _lib.ptr(flow), B, H, W, int(radius), _lib.ptr(out), _lib.stream_of(flow)),
                   "pp_corr_lookup")
    return out


class CorrLookup(nn.Module):
    """Correlation lookup operator (RAFT), same constructor as the reference (utils/corr_lookup.py:88-98).

    Parameter-free, so checkpoints are unaffected; kept an nn.Module because FlowDecoder stores it in an
    nn.ModuleList (model/stage3/flow_decoder.py:44).
    """

    def __init__(self, radius: int = 4, mode: str = 'bilinear', padding_mode: str = 'zeros',
                 align_corners: bool = True) -> None:
        super().__init__()
        self.r = radius
        self.mode = mode
        self.padding_mode = padding_mode
        self.align_corners = align_corners

    def forward(self, corr_pyramid: Sequence[Tensor], flow: Tensor) -> Tensor:
        if self.mode != 'bilinear' or self.padding_mode != 'zeros' or not self.align_corners:
            raise NotImplementedError(
                "picopose_b200.CorrLookup implements the configuration PicoPose uses "
                "(bilinear, zeros padding, align_corners=True)")
        from .correlation import LazyCorrelationPyramid, LazyLookup, encoder_fusion_enabled, windowed_correlation
        if isinstance(corr_pyramid, LazyCorrelationPyramid):
            if corr_pyramid.fusable(self.r) and encoder_fusion_enabled():
                # our MotionEncoder consumes the result (overlay): leave the lookup to it, fused with its first 1x1 conv
                return LazyLookup(corr_pyramid, flow, self.r)
            if corr_pyramid.fusable(self.r):
                # fused CorrelationPyramid + lookup: the all-pairs volume is never built
                return windowed_correlation(corr_pyramid.feat1, corr_pyramid.feat2, flow, corr_pyramid.num_levels, self.r)
            corr_pyramid = corr_pyramid.for_lookup(self.r)
        return corr_lookup(corr_pyramid, flow, self.r)
