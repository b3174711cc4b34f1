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


_SAMPLE_MODES = {'bilinear': 0, 'nearest': 1, 'bicubic': 2}      # pp_sample_mode
_PAD_MODES = {'zeros': 0, 'border': 1, 'reflection': 2}          # pp_pad_mode


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
    if mode not in _SAMPLE_MODES or padding_mode not in _PAD_MODES:
        raise ValueError(f"bilinear_sample: mode must be one of {sorted(_SAMPLE_MODES)}, padding_mode one of "
                         f"{sorted(_PAD_MODES)}; got '{mode}', '{padding_mode}'")      # F.grid_sample raises ValueError too
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
        if mode == 'bilinear' and padding_mode == 'zeros':            # what PicoPose uses: the tuned kernel
            _lib.check(lib.pp_bilinear_sample(_lib.ptr(feat), _lib.ptr(grid), N, Cc, Hf, Wf, Ho, Wo, int(chw),
                                              int(bool(align_corners)), int(bool(scale)), _lib.ptr(out),
                                              _lib.stream_of(feat)), "pp_bilinear_sample")
        else:
            _lib.check(lib.pp_grid_sample(_lib.ptr(feat), _lib.ptr(grid), N, Cc, Hf, Wf, Ho, Wo, int(chw),
                                          int(bool(align_corners)), int(bool(scale)), _SAMPLE_MODES[mode],
                                          _PAD_MODES[padding_mode], _lib.ptr(out), _lib.stream_of(feat)), "pp_grid_sample")
    return out


def _corr_lookup_general(corr_pyramid: Sequence[Tensor], flow: Tensor, radius: int, mode: str, padding_mode: str,
                         align_corners: bool) -> Tensor:
    """CorrLookup.forward for the argument combinations PicoPose does not use (utils/corr_lookup.py:100-134 with another
    interpolation / padding mode, or align_corners=False): the window coordinates are laid out as the reference does
    and sampled level by level with pp_grid_sample.  A compatibility path, not a tuned one."""
    B, _, H, W = flow.shape
    r = int(radius)
    D = 2 * r + 1
    dev = flow.device
    xx = torch.arange(0, W, device=dev)
    yy = torch.arange(0, H, device=dev)
    centre = (coords_grid(B, xx, yy) + flow.float()).permute(0, 2, 3, 1).reshape(B * H * W, 1, 1, 2)
    d = torch.linspace(-r, r, D, device=dev)
    # delta[a, b] = (d[a], d[b]) is added to an (x, y) centre: the FIRST window axis moves x (:116-121,126)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, D, D, 2)
    outs = []
    for i, corr in enumerate(corr_pyramid):
        coords = centre / 2 ** i + delta
        outs.append(bilinear_sample(corr, coords, mode, padding_mode, align_corners).view(B, H, W, -1))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


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
        from .correlation import (LazyCorrelationPyramid, LazyLookup, TiledPyramid, encoder_fusion_enabled,
                                  windowed_correlation)
        if self.mode != 'bilinear' or self.padding_mode != 'zeros' or not self.align_corners:
            # not a configuration PicoPose uses (model/stage3/flow_decoder.py:44 takes the defaults): generic sampler
            if isinstance(corr_pyramid, LazyCorrelationPyramid):
                corr_pyramid = corr_pyramid.for_lookup(0)
            if isinstance(corr_pyramid, TiledPyramid):
                corr_pyramid = corr_pyramid.rowmajor()
            _lib.require_cuda(flow, *corr_pyramid)
            _lib.require_inference("CorrLookup", flow, *corr_pyramid)
            return _corr_lookup_general(corr_pyramid, flow, self.r, self.mode, self.padding_mode, self.align_corners)
        if isinstance(corr_pyramid, LazyCorrelationPyramid):
            if corr_pyramid.fusable(self.r) and encoder_fusion_enabled():
                # our MotionEncoder consumes the result (overlay): leave the lookup to it, fused with its first 1x1 conv
                return LazyLookup(corr_pyramid, flow, self.r)
            if corr_pyramid.fusable(self.r):
                # fused CorrelationPyramid + lookup: the all-pairs volume is never built
                return windowed_correlation(corr_pyramid.feat1, corr_pyramid.feat2, flow, corr_pyramid.num_levels, self.r)
            corr_pyramid = corr_pyramid.for_lookup(self.r)
        return corr_lookup(corr_pyramid, flow, self.r)
