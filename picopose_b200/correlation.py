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


def correlation_pyramid(feat1: torch.Tensor, feat2: torch.Tensor, num_levels: int,
                        mode: Optional[str] = None) -> Sequence[torch.Tensor]:
    """-> [ (N*H*W, 1, H>>l, W>>l) fp32 for l in range(num_levels) ], corr = <f1, f2> / sqrt(C), 2x2 average pools."""
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
    with torch.cuda.device(feat1.device):
        _lib.check(lib.pp_correlation_pyramid(_lib.ptr(a), _lib.ptr(b), N, H, W, a.shape[-1], 1.0 / math.sqrt(Cc),
                                              num_levels, ptrs, default_cluster(), _lib.stream_of(feat1)),
                   "pp_correlation_pyramid")
    return levels


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
        f1t = torch.empty(N, H * W, Cc, dtype=torch.float32, device=dev)
        levels = [torch.empty(N, (H >> l) * (W >> l), Cc, dtype=torch.float32, device=dev) for l in range(num_levels)]
        ptrs = (C.c_void_p * num_levels)(*[t.data_ptr() for t in levels])
        # one launch: position-major copies of feat1 and of feat2 average-pooled to every level
        _lib.check(lib.pp_windowed_correlation_prepare_all(_lib.ptr(f1), _lib.ptr(f2), N, Cc, H, W, num_levels,
                                                           _lib.ptr(f1t), ptrs, st), "pp_windowed_correlation_prepare_all")
        out = torch.empty(N, num_levels * D * D, H, W, dtype=torch.float32, device=dev)
        _lib.check(lib.pp_windowed_correlation(_lib.ptr(f1t), ptrs, num_levels, _lib.ptr(fl), N, Cc, H, W, int(radius),
                                               _lib.ptr(out), st), "pp_windowed_correlation")
    return out


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

    def fusable(self, radius: int) -> bool:
        Cc = self.feat1.shape[1]
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
