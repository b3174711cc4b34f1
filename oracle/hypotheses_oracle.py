"""Oracle for the hypothesis hand-over (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates ``Net.select_template_data`` (reference model/picopose.py:52-70) with plain indexing: the reference's
``torch.gather(x, 1, idx[:, None, ...].repeat(1, 1, *x.shape[2:])).squeeze(1)`` picks, for every detection b, the
slice ``x[b, idx[b]]``.  Pinned by tests/golden/hyp_select.npz (outputs of the reference method itself, minted by
oracle/make_golden.py) in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import torch

TEMPLATE_KEYS = ("tem_pose", "tem_K", "tem_M", "tem_mask", "tem_rgb", "tem_pts3d")   # :55-62
REAL_KEYS = ("real_pts2d", "real_K", "real_M", "real_mask", "real_pose")              # :65-69


def select_template_data(end_points, pred_id_src, k):
    b = torch.arange(pred_id_src.shape[0])
    out = {key: end_points[key][b, pred_id_src[:, k]] for key in TEMPLATE_KEYS}
    for key in REAL_KEYS:
        out[key] = end_points[key]
    return out


def hypothesis_loop(select, forward_hyp, end_points, pred_id_src, features_real):
    """The reference's loop (model/picopose.py:107-110): K sequential stage-2/3 passes."""
    return [forward_hyp(select(end_points, pred_id_src, k), features_real) for k in range(pred_id_src.shape[1])]
