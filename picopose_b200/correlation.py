"""All-pairs correlation pyramid on the tensor cores: host side of the reference's ``CorrelationPyramid``
(model/stage3/raft_decoder.py:14-53), the producer of the volumes ``CorrLookup`` reads."""
from __future__ import annotations

import ctypes as C
import math
import os
from collections.abc import Sequence
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .matching import default_cluster, prepare_features


def default_corr_mode() -> str:
    # the reference computes this product in fp32 (torch.matmul), so the fp32-accurate split mode is the default
    return os.environ.get("PICOPOSE_B200_CORR_MODE", "fp32")


def tileable(H: int, W: int, num_levels: int) -> bool:
    """Every level of an (H x W) pyramid can be stored as 4 x 8 tiles (h % 4 == 0, w % 8 == 0, exact halvings)."""
    s = 1 << (num_levels - 1)
    return H % s == 0 and W % s == 0 and (H // s) % 4 == 0 and (W // s) % 8 == 0


def retile_volume(vol: torch.Tensor, to_tiled: bool = True) -> torch.Tensor:
    """(Q, 1, h, w) fp32 slices: reference row-major layout <-> the tiled layout of `TiledPyramid` (a permutation
    inside every slice; the tensor keeps its logical shape)."""
    _lib.require_cuda(vol)
    lib = _lib.load()
    vol = vol.float().contiguous()
    Q, _, h, w = vol.shape
    out = torch.empty_like(vol)
    with torch.cuda.device(vol.device):
        _lib.check(lib.pp_volume_retile(_lib.ptr(vol), _lib.ptr(out), Q, h, w, int(bool(to_tiled)), _lib.stream_of(vol)),
                   "pp_volume_retile")
    return out


class TiledPyramid(Sequence):
    """A correlation pyramid whose slices are stored as 4-row x 8-column tiles, one 128-byte line per tile.

    DRAM moves whole 128-byte lines on this GPU, so what a lookup window costs is the number of lines it touches: a
    (2r+2)^2 window crosses ~6.9 tiles at r = 4 against ~11.7 row segments in the reference's row-major slices.  Our own
    CorrelationPyramid can write this layout at no extra cost (the contraction's epilogue and the pooling kernel address
    it directly); `CorrLookup` recognises the type and reads it with `pp_corr_lookup_tiled`.  Any other consumer that
    indexes or iterates it gets the reference's row-major volumes (converted once, on demand).
    """

    def __init__(self, tiled_levels):
        self.tiled_levels = list(tiled_levels)        # logical shape (Q, 1, h, w), tiled storage
        self._rowmajor = None

    @classmethod
    def from_volumes(cls, volumes) -> "TiledPyramid":
        return cls([retile_volume(v, True) for v in volumes])

    def rowmajor(self):
        if self._rowmajor is None:
            self._rowmajor = [retile_volume(v, False) for v in self.tiled_levels]
        return self._rowmajor

    def __len__(self):
        return len(self.tiled_levels)

    def __getitem__(self, i):
        return self.rowmajor()[i]

    def __iter__(self):
        return iter(self.rowmajor())


def correlation_pyramid(feat1: torch.Tensor, feat2: torch.Tensor, num_levels: int,
                        mode: Optional[str] = None, layout: str = "rowmajor"):
    """-> [ (N*H*W, 1, H>>l, W>>l) fp32 for l in range(num_levels) ], corr = <f1, f2> / sqrt(C), 2x2 average pools.
    layout="tiled" returns a `TiledPyramid` (same values, lookup-friendly storage)."""
    if layout not in ("rowmajor", "tiled"):
        raise ValueError("layout must be 'rowmajor' or 'tiled'")
    _lib.require_cuda(feat1, feat2)
    _lib.require_inference("CorrelationPyramid", feat1, feat2)
    lib = _lib.load()
    N, Cc, H, W = feat1.shape
    if tuple(feat2.shape) != (N, Cc, H, W):
        raise ValueError("feat1 and feat2 must have the same shape")
    mode = default_corr_mode() if mode is None else mode
    a, _ = prepare_features(feat1, mode, is_query=True)
    b, _ = prepare_features(feat2, mode, is_query=False)
    levels = [torch.empty(N * H * W, 1, H >> l, W >> l, dtype=torch.float32, device=feat1.device)
              for l in range(num_levels)]
    ptrs = (C.c_void_p * num_levels)(*[t.data_ptr() for t in levels])
    if layout == "tiled" and not tileable(H, W, num_levels):
        raise ValueError(f"a {H}x{W} pyramid of {num_levels} levels cannot be tiled (every level needs h % 4 == 0, w % 8 == 0)")
    fn = lib.pp_correlation_pyramid_tiled if layout == "tiled" else lib.pp_correlation_pyramid
    with torch.cuda.device(feat1.device):
        _lib.check(fn(_lib.ptr(a), _lib.ptr(b), N, H, W, a.shape[-1], 1.0 / math.sqrt(Cc),
                      num_levels, ptrs, default_cluster(), _lib.stream_of(feat1)), "pp_correlation_pyramid")
    return TiledPyramid(levels) if layout == "tiled" else levels


def tiled_layout_enabled() -> bool:
    return os.environ.get("PICOPOSE_B200_TILED_VOLUME", "1") != "0"


def fused_corr_enabled() -> bool:
    return os.environ.get("PICOPOSE_B200_FUSED_CORR", "1") != "0"


def windowed_correlation(feat1: torch.Tensor, feat2: torch.Tensor, flow: torch.Tensor, num_levels: int,
                         radius: int) -> torch.Tensor:
    """CorrLookup(radius)(CorrelationPyramid(num_levels)(feat1, feat2), flow) without the all-pairs volume.

    -> (N, L*(2r+1)^2, H, W) fp32, same channel order and values (to ~1e-6) as the two-step path.
    """
    _lib.require_cuda(feat1, feat2, flow)
    _lib.require_inference("CorrLookup (fused correlation)", feat1, feat2, flow)
    lib = _lib.load()
    f1 = feat1.float().contiguous()
    f2 = feat2.float().contiguous()
    fl = flow.float().contiguous()
    N, Cc, H, W = f1.shape
    dev = f1.device
    D = 2 * int(radius) + 1
    with torch.cuda.device(dev):
        st = _lib.stream_of(f1)
        f1t, levels, ptrs = _position_major(f1, f2, num_levels)
        out = torch.empty(N, num_levels * D * D, H, W, dtype=torch.float32, device=dev)
        _lib.check(lib.pp_windowed_correlation(_lib.ptr(f1t), ptrs, num_levels, _lib.ptr(fl), N, Cc, H, W, int(radius),
                                               _lib.ptr(out), st), "pp_windowed_correlation")
    return out


def _position_major(feat1, feat2, num_levels):
    """One launch: position-major copies of feat1 and of feat2 average-pooled to every level (inputs of the fused kernels)."""
    lib = _lib.load()
    f1 = feat1.float().contiguous()
    f2 = feat2.float().contiguous()
    N, Cc, H, W = f1.shape
    dev = f1.device
    f1t = torch.empty(N, H * W, Cc, dtype=torch.float32, device=dev)
    levels = [torch.empty(N, (H >> l) * (W >> l), Cc, dtype=torch.float32, device=dev) for l in range(num_levels)]
    ptrs = (C.c_void_p * num_levels)(*[t.data_ptr() for t in levels])
    _lib.check(lib.pp_windowed_correlation_prepare_all(_lib.ptr(f1), _lib.ptr(f2), N, Cc, H, W, num_levels,
                                                       _lib.ptr(f1t), ptrs, _lib.stream_of(f1)), "pp_windowed_correlation_prepare_all")
    return f1t, levels, ptrs


def conv_fusable(feat1: torch.Tensor, num_levels: int, radius: int, cout: int) -> bool:
    """Shapes pp_windowed_correlation_conv1x1 covers (the TMA-tiled kernel with the weight matrix in its stage buffers)."""
    Cc = feat1.shape[1]
    D = 2 * int(radius) + 1
    return (1 <= radius <= 2 and Cc % 32 == 0 and 1 <= num_levels <= 4 and cout % 8 == 0 and cout * num_levels * D * D <= 40960
            and num_levels * D * D * 65 * 4 + 2 * (24 * 24 + 64) * 32 * 4 + 64 * (D + 2) ** 2 * 4 + num_levels * 64 * 4 * D * 4 < 215 * 1024)


def windowed_correlation_conv(feat1: torch.Tensor, feat2: torch.Tensor, flow: torch.Tensor, num_levels: int, radius: int,
                              weight: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = True) -> torch.Tensor:
    """act(conv1x1(CorrLookup(radius)(CorrelationPyramid(num_levels)(feat1, feat2), flow))) in one kernel after the layout
    pass: the lookup tile is multiplied by the (cout x L*D*D) weight matrix while it sits in shared memory (SURVEY 8(f)-3;
    reference: model/stage3/flow_decoder.py:59-62 + model/stage3/raft_decoder.py:113-116,157).  -> (N, cout, H, W) fp32."""
    _lib.require_cuda(feat1, feat2, flow, weight, bias)
    _lib.require_inference("fused lookup + 1x1 conv", feat1, feat2, flow, weight)
    lib = _lib.load()
    N, Cc, H, W = feat1.shape
    D = 2 * int(radius) + 1
    w = weight.detach().float().reshape(weight.shape[0], -1).contiguous()
    cout = w.shape[0]
    if w.shape[1] != num_levels * D * D:
        raise ValueError(f"weight has {w.shape[1]} input channels, the lookup produces {num_levels * D * D}")
    b = None if bias is None else bias.detach().float().contiguous()
    fl = flow.float().contiguous()
    with torch.cuda.device(feat1.device):
        f1t, levels, ptrs = _position_major(feat1, feat2, num_levels)
        out = torch.empty(N, cout, H, W, dtype=torch.float32, device=feat1.device)
        _lib.check(lib.pp_windowed_correlation_conv1x1(_lib.ptr(f1t), ptrs, num_levels, _lib.ptr(fl), N, Cc, H, W, int(radius),
                                                       _lib.ptr(w), _lib.ptr(b), cout, int(bool(relu)), _lib.ptr(out),
                                                       _lib.stream_of(feat1)), "pp_windowed_correlation_conv1x1")
    return out


class LazyLookup:
    """What CorrLookup returns for a LazyCorrelationPyramid when the consumer is our MotionEncoder (the overlay's
    model/stage3/raft_decoder.py): the lookup not yet computed, so that the encoder's first 1x1 convolution can be fused
    into it.  `materialise()` gives the plain (N, L*D*D, H, W) lookup tensor for anybody else."""

    def __init__(self, pyramid: "LazyCorrelationPyramid", flow: torch.Tensor, radius: int):
        self.pyramid, self.flow, self.radius = pyramid, flow, int(radius)

    def materialise(self) -> torch.Tensor:
        p = self.pyramid
        return windowed_correlation(p.feat1, p.feat2, self.flow, p.num_levels, self.radius)

    def conv1x1(self, weight, bias=None, relu=True):
        """-> act(conv1x1(lookup)) fused, or None when the shapes are outside the fused kernel's range."""
        p = self.pyramid
        cout = weight.shape[0]
        if not conv_fusable(p.feat1, p.num_levels, self.radius, cout) or p.feat1.shape[-1] * p.feat1.shape[-2] < 64:
            return None
        return windowed_correlation_conv(p.feat1, p.feat2, self.flow, p.num_levels, self.radius, weight, bias, relu)


ENCODER_FUSION = False     # set by the overlay's raft_decoder once OUR MotionEncoder is the consumer of CorrLookup's result


def encoder_fusion_enabled() -> bool:
    return ENCODER_FUSION and os.environ.get("PICOPOSE_B200_FUSE_CONV", "1") != "0"


def motion_encoder_forward(encoder, corr, flow):
    """MotionEncoder.forward (model/stage3/raft_decoder.py:146-161) with the first corr_net layer fused into the lookup
    when `corr` is a LazyLookup; `encoder` is the reference's module (its weights, its remaining layers)."""
    corr_feat = None
    if isinstance(corr, LazyLookup):
        first = encoder.corr_net[0]
        conv = getattr(first, "conv", None)
        act = getattr(first, "activate", getattr(first, "act", None))
        plain = (isinstance(conv, nn.Conv2d) and conv.kernel_size == (1, 1) and conv.stride == (1, 1)
                 and conv.padding == (0, 0) and conv.groups == 1 and not getattr(first, "with_norm", False)
                 and isinstance(act, nn.ReLU))
        x = corr.conv1x1(conv.weight, conv.bias, relu=True) if plain else None
        if x is None:
            corr = corr.materialise()
        else:
            rest = encoder.corr_net[1:]
            corr_feat = rest(x) if len(rest) else x
    if corr_feat is None:
        corr_feat = encoder.corr_net(corr)
    flow_feat = encoder.flow_net(flow)
    out = encoder.out_net(torch.cat([corr_feat, flow_feat], dim=1))
    return torch.cat([out, flow], dim=1)


class LazyCorrelationPyramid(Sequence):
    """What CorrelationPyramid.forward returns: the two feature maps, not yet multiplied.

    CorrLookup recognises it and runs the fused windowed correlation (the 64 MiB/sample volume of
    model/stage3/raft_decoder.py:43-47 is never built).  Any other consumer can still index / iterate it like
    the reference's list of volumes; they are materialised on first use by the tensor-core path.
    """

    def __init__(self, feat1: torch.Tensor, feat2: torch.Tensor, num_levels: int):
        self.feat1, self.feat2, self.num_levels = feat1, feat2, num_levels
        self._volumes = None

    def materialise(self):
        if self._volumes is None:
            self._volumes = correlation_pyramid(self.feat1, self.feat2, self.num_levels)
        return self._volumes

    def for_lookup(self, radius: int):
        """The volumes for a CorrLookup that cannot take the fused path: tiled when the shapes allow (nobody else sees
        them), else the reference layout."""
        if self._volumes is not None:
            return self._volumes
        H, W = self.feat1.shape[-2:]
        # tiled volumes: a third less DRAM traffic and 7-25 % faster than the reference's row-major layout at every radius
        # the banded kernel is compiled for (profiles/r2z_lookup_sweep.json)
        if tiled_layout_enabled() and 1 <= radius <= 8 and tileable(H, W, self.num_levels):
            return correlation_pyramid(self.feat1, self.feat2, self.num_levels, layout="tiled")
        return self.materialise()

    def fusable(self, radius: int) -> bool:
        """Whether CorrLookup(radius) should take the no-volume path.  Radius 1-2 has the TMA-tiled kernel and always does;
        from radius 3 on only the per-query kernel exists, which loses to pyramid + lookup on mid-sized maps (16 x 256 x
        32^2, L=2, r=4: 0.29 ms against 0.13 ms; at 64^2 the two are level, 1.2 ms, and the volume would be 84 MB per sample;
        at 16^2 both are launch-bound -- profiles/r2f_stage3_r4.json), so those materialise their (small) volumes."""
        Cc = self.feat1.shape[1]
        H, W = self.feat1.shape[-2:]
        if radius >= 3 and 512 <= H * W <= 2048:
            return False
        return (self._volumes is None and 1 <= radius <= 8 and Cc % 4 == 0 and Cc <= 2048
                and self.num_levels * (2 * radius + 1) ** 2 * 33 * 4 + 8 * Cc * 4 < 190 * 1024)

    def __len__(self):
        return self.num_levels

    def __getitem__(self, i):
        return self.materialise()[i]

    def __iter__(self):
        return iter(self.materialise())


class CorrelationPyramid(nn.Module):
    """Drop-in for model/stage3/raft_decoder.py:14-53 (same constructor, parameter-free)."""

    def __init__(self, num_levels: int = 4) -> None:
        super().__init__()
        self.num_levels = num_levels

    def forward(self, feat1: torch.Tensor, feat2: torch.Tensor) -> Sequence[torch.Tensor]:
        if fused_corr_enabled():
            return LazyCorrelationPyramid(feat1, feat2, self.num_levels)
        return correlation_pyramid(feat1, feat2, self.num_levels)
