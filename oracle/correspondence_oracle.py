"""Oracle for the correspondence glue (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates the reference's ``utils/correspondence.py``:

* ``init_correspondences``    <- compute_init_correspondences, utils/correspondence.py:10-26
  (with utils/torch_utils.py:114-135 apply_affine and :297-305 init_points2d_torch)
* ``stage3_correspondences``  <- compute_stage3_correspondences, utils/correspondence.py:28-59

Parity pinning: tests/golden/corresp_*.npz (reference outputs, made by
oracle/make_golden.py) are replayed by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch


def init_correspondences(pred_Ms, tem_mask, size=(16, 16)):
    """-> init_flow (B,2,h,w), init_certainty (B,1,h,w).

    Patch centres ((i+0.5)*patch) are pushed through the 3x3 affine, divided by
    the patch size, masked, and the integer grid is subtracted
    (utils/correspondence.py:13-25).  Channel 0 is x (varies with w).
    """
    B, Hm, Wm = tem_mask.shape
    assert Hm == Wm
    h, w = size
    patch = Hm // h
    ys = torch.div(torch.arange(h) * Hm, h, rounding_mode="floor")
    xs = torch.div(torch.arange(w) * Wm, w, rounding_mode="floor")
    m = tem_mask.float()[:, ys][:, :, xs]                                  # (B,h,w)  :14
    c = torch.arange(0, Hm, patch, dtype=torch.float32) + patch / 2        # torch_utils.py:298-301
    cx = c.view(1, 1, -1).expand(B, h, w)                                  # x centre depends on w
    cy = c.view(1, -1, 1).expand(B, h, w)
    M = pred_Ms.float()

    def row(k):
        return M[:, k, 0].view(B, 1, 1) * cx + M[:, k, 1].view(B, 1, 1) * cy + M[:, k, 2].view(B, 1, 1)

    den = row(2)
    px = row(0) / den / patch                                              # :17
    py = row(1) / den / patch
    gx = torch.arange(w, dtype=torch.float32).view(1, 1, w).expand(B, h, w)
    gy = torch.arange(h, dtype=torch.float32).view(1, h, 1).expand(B, h, w)
    flow = torch.stack([px * m - gx, py * m - gy], dim=1)                  # :24
    return flow, m.unsqueeze(1)                                            # :25


def stage3_correspondences(pred_flow, pred_certainty, threshold=0.5):
    """-> (tar_pts, src_pts), each (B, H*W, 2) int64, flat index k = w*H + h.

    A cell is kept when 0 < x < H-1, 0 < y < W-1 (strict) and
    sigmoid(certainty) > threshold; kept cells give src=(w,h), tar=trunc(x,y),
    all others -1 (utils/correspondence.py:34-57).
    """
    B, _, H, W = pred_flow.shape
    gx = torch.arange(W, dtype=torch.float32).view(1, 1, W)
    gy = torch.arange(H, dtype=torch.float32).view(1, H, 1)
    x = pred_flow[:, 0].float() + gx
    y = pred_flow[:, 1].float() + gy
    keep = (x > 0) & (y > 0) & (x < H - 1) & (y < W - 1)
    keep &= torch.sigmoid(pred_certainty[:, 0].float()) > threshold
    neg = torch.full((B, H, W), -1, dtype=torch.long)
    wi = torch.arange(W).view(1, 1, W).expand(B, H, W)
    hi = torch.arange(H).view(1, H, 1).expand(B, H, W)
    src = torch.stack([torch.where(keep, wi, neg), torch.where(keep, hi, neg)], dim=-1)
    tar = torch.stack([torch.where(keep, x.long(), neg), torch.where(keep, y.long(), neg)], dim=-1)
    # "b h w c -> b (w h) c"
    src = src.permute(0, 2, 1, 3).reshape(B, W * H, 2)
    tar = tar.permute(0, 2, 1, 3).reshape(B, W * H, 2)
    return tar, src
