"""All-pairs correlation pyramid on the tensor cores: host side of the reference's ``CorrelationPyramid``
(model/stage3/raft_decoder.py:14-53), the producer of the volumes ``CorrLookup`` reads."""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Sequence

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


class CorrelationPyramid(nn.Module):
    """Drop-in for model/stage3/raft_decoder.py:14-53 (same constructor, parameter-free)."""

    def __init__(self, num_levels: int = 4) -> None:
        super().__init__()
        self.num_levels = num_levels

    def forward(self, feat1: torch.Tensor, feat2: torch.Tensor) -> Sequence[torch.Tensor]:
        return correlation_pyramid(feat1, feat2, self.num_levels)
