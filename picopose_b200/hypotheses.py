"""Top-k hypothesis hand-over from stage 1 to stages 2/3.

The reference ranks the template views, then for each of the `hyp` best views gathers that view's data with six
``torch.gather(x, 1, idx[:, None, ...].repeat(...))`` calls (``Net.select_template_data``, model/picopose.py:52-70)
and runs stages 2/3 on it (``forward_test_hyp``), one hypothesis after the other (:107-110).  Here:

* ``select_template_data(end_points, pred_id_src, k)`` -- same arguments and result as the reference method, ONE
  kernel launch for all six tensors (no index tensors);
* ``select_all_hypotheses(end_points, pred_id_src)`` -- all hypotheses in one launch, hypothesis-major
  ``(K*B, ...)`` batch, the detection-side entries repeated to match;
* ``forward_test_batched(net, end_points, hyp)`` -- ``Net.forward_test`` with ONE stage-2/3 pass over the
  ``K*B`` batch instead of K passes over B; returns the same list of K output dicts;
* ``patch_net(net)`` -- installs the two on a reference ``Net`` instance (or class).

Every stage-2/3 function of the reference is batch-agnostic, so the batched pass computes per sample what the
loop computes (convolutions may pick different algorithms for another batch size: equal within float tolerance).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch

from . import _lib

TEMPLATE_KEYS = ("tem_pose", "tem_K", "tem_M", "tem_mask", "tem_rgb", "tem_pts3d")       # model/picopose.py:55-62
REAL_KEYS = ("real_pts2d", "real_K", "real_M", "real_mask", "real_pose")                  # model/picopose.py:65-69


def _gather_views(tensors: List[torch.Tensor], pred_id_src: torch.Tensor, hyp_sel: int) -> List[torch.Tensor]:
    """tensors[i] (B, N, ...) -> (rows, ...) with rows = B (hypothesis hyp_sel) or K*B (hyp_sel = -1, hypothesis-major)."""
    _lib.require_cuda(pred_id_src, *tensors)
    lib = _lib.load()
    B, K = pred_id_src.shape
    idx = pred_id_src.to(torch.int64).contiguous()
    srcs = [t.contiguous() for t in tensors]
    N = srcs[0].shape[1]
    for t in srcs:
        if t.shape[0] != B or t.shape[1] != N:
            raise ValueError(f"per-view tensors must be (B={B}, N={N}, ...); got {tuple(t.shape)}")
    rows = B if hyp_sel >= 0 else B * K
    outs = [torch.empty((rows,) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device) for t in srcs]
    n = len(srcs)
    view_bytes = [t[0, 0].numel() * t.element_size() for t in srcs]
    if B == 0 or any(v == 0 for v in view_bytes):
        return outs
    with torch.cuda.device(idx.device):
        _lib.check(lib.pp_select_templates((C.c_void_p * n)(*[t.data_ptr() for t in srcs]),
                                           (C.c_void_p * n)(*[o.data_ptr() for o in outs]),
                                           (C.c_int64 * n)(*view_bytes), n, B, N, idx.data_ptr(), K, int(hyp_sel),
                                           _lib.stream_of(idx)), "pp_select_templates")
    return outs


def select_template_data(end_points: Dict[str, torch.Tensor], pred_id_src: torch.Tensor, k: int) -> Dict[str, torch.Tensor]:
    """Drop-in for Net.select_template_data(end_points, pred_id_src, k), model/picopose.py:52-70."""
    if not 0 <= k < pred_id_src.shape[1]:
        raise IndexError(f"hypothesis {k} of {pred_id_src.shape[1]}")
    picked = _gather_views([end_points[key] for key in TEMPLATE_KEYS], pred_id_src, k)
    out = dict(zip(TEMPLATE_KEYS, picked))
    for key in REAL_KEYS:
        out[key] = end_points[key]
    return out


def select_all_hypotheses(end_points: Dict[str, torch.Tensor], pred_id_src: torch.Tensor) -> Dict[str, torch.Tensor]:
    """All K hypotheses of every detection as one (K*B, ...) batch, hypothesis-major: rows [k*B, (k+1)*B) are what
    select_template_data(end_points, pred_id_src, k) returns."""
    K = pred_id_src.shape[1]
    picked = _gather_views([end_points[key] for key in TEMPLATE_KEYS], pred_id_src, -1)
    out = dict(zip(TEMPLATE_KEYS, picked))
    for key in REAL_KEYS:
        t = end_points[key]
        out[key] = t.repeat((K,) + (1,) * (t.dim() - 1))
    return out


@torch.no_grad()
def forward_test_batched(net, end_points: Dict[str, torch.Tensor], hyp: int = 5):
    """Net.forward_test (model/picopose.py:97-112) with one stage-2/3 pass over the K*B hypothesis batch.

    `net` is the reference's Net (feature_extractor, forward_test_hyp are its own); the matcher and the gather are
    ours.  Returns the reference's list of `hyp` output dicts."""
    import torch.nn.functional as F
    from .matching import matching_templates
    features_real = net.feature_extractor(end_points["real_rgb"])
    feature_tem = F.normalize(end_points["template_feature"], dim=2)
    _, pred_id_src = matching_templates(feature_tem, features_real[-1], end_points["tem_mask"], end_points["real_mask"], topk=hyp)
    batch = select_all_hypotheses(end_points, pred_id_src)
    feats = [f.repeat((hyp,) + (1,) * (f.dim() - 1)) for f in features_real]
    out = net.forward_test_hyp(batch, feats)
    B = pred_id_src.shape[0]
    return [{key: val[k * B:(k + 1) * B] for key, val in out.items()} for k in range(hyp)]


def patch_net(net, batched: bool = True):
    """Installs the single-launch gather (and, with `batched`, the one-pass hypothesis loop) on a reference Net
    instance or on the Net class; everything else of the object is untouched (no parameters are added)."""
    import types
    if isinstance(net, type):
        net.select_template_data = lambda self, end_points, pred_id_src, k: select_template_data(end_points, pred_id_src, k)
        if batched:
            net.forward_test = lambda self, end_points, hyp=5: forward_test_batched(self, end_points, hyp)
        return net
    net.select_template_data = types.MethodType(lambda self, e, p, k: select_template_data(e, p, k), net)
    if batched:
        net.forward_test = types.MethodType(lambda self, e, hyp=5: forward_test_batched(self, e, hyp), net)
    return net
