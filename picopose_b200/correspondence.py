"""Correspondence glue: host side of the reference's ``utils/correspondence.py`` on libpicopose_b200."""
from __future__ import annotations

import torch

from . import _lib
from .corr_lookup import coords_grid  # re-exported like the reference module does  # noqa: F401


def compute_init_correspondences(pred_Ms, tem_mask, size=(16, 16)):
    """Drop-in for utils/correspondence.py:10-26 -> (init_flow (B,2,h,w), init_certainty (B,1,h,w))."""
    _lib.require_cuda(pred_Ms, tem_mask)
    _lib.require_inference("compute_init_correspondences", pred_Ms)
    lib = _lib.load()
    B, H, W = tem_mask.shape
    assert H == W
    Ms = pred_Ms.float().contiguous()
    mask = tem_mask.float().contiguous()
    h, w = int(size[0]), int(size[1])
    flow = torch.empty(B, 2, h, w, dtype=torch.float32, device=mask.device)
    cert = torch.empty(B, 1, h, w, dtype=torch.float32, device=mask.device)
    with torch.cuda.device(mask.device):
        _lib.check(lib.pp_init_correspondences(_lib.ptr(Ms), _lib.ptr(mask), B, H, W, h, w, _lib.ptr(flow),
                                               _lib.ptr(cert), _lib.stream_of(mask)), "pp_init_correspondences")
    return flow, cert


def compute_stage3_correspondences(pred_flow, pred_certainty, threshold=0.5):
    """Drop-in for utils/correspondence.py:28-59 -> (tar_pts, src_pts), each (B, H*W, 2) int64.

    No host synchronisation (the reference goes through torch.nonzero).
    """
    _lib.require_cuda(pred_flow, pred_certainty)
    lib = _lib.load()
    flow = pred_flow.float().contiguous()
    cert = pred_certainty.float().contiguous()
    B, _, H, W = flow.shape
    tar = torch.empty(B, H * W, 2, dtype=torch.int64, device=flow.device)
    src = torch.empty(B, H * W, 2, dtype=torch.int64, device=flow.device)
    with torch.cuda.device(flow.device):
        _lib.check(lib.pp_stage3_correspondences(_lib.ptr(flow), _lib.ptr(cert), B, H, W, float(threshold),
                                                 _lib.ptr(tar), _lib.ptr(src), _lib.stream_of(flow)),
                   "pp_stage3_correspondences")
    return tar, src
