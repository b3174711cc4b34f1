"""Stage-1 matching: host side of the reference's ``utils/matching.py`` on libpicopose_b200.

Same names, argument meaning and results as the reference:

* ``matching_templates(src_feats, tar_feat, src_masks, tar_mask, topk=5)``  (utils/matching.py:29-69)
* ``matching_features_similarity(src_feat, tar_feat, src_mask, tar_mask)``  (utils/matching.py:6-26)

plus the pieces a serving loop wants: ``TemplateBank`` (a template bank
normalised / cast / laid out once per object, the way run_test.py:121-134 caches
template features once per object) and ``template_scores`` (dense per-view
scores and the coarse 2D-2D correspondences = bidirectional argmax indices).

PyTorch only allocates tensors and provides the stream; every arithmetic step
runs in the sm_100a kernels behind the C ABI (include/picopose_b200.h).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib

_WORKSPACE_LIMIT = int(os.environ.get("PICOPOSE_B200_WORKSPACE_MB", "1024")) << 20
_MAX_DETS_PER_LAUNCH = 1024   # tile-prefix table of the contraction kernel lives in shared memory


def default_mode() -> str:
    return os.environ.get("PICOPOSE_B200_MODE", "bf16")


def default_cluster() -> int:
    return int(os.environ.get("PICOPOSE_B200_CLUSTER", "0"))


def _mode_id(mode: Optional[str]) -> int:
    mode = default_mode() if mode is None else mode
    if mode not in _lib.MODES:
        raise ValueError(f"unknown mode '{mode}' (expected one of {sorted(_lib.MODES)})")
    return _lib.MODES[mode]


def _as_f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def prepare_features(feats: torch.Tensor, mode: Optional[str] = None, is_query: bool = False):
    """(..., C, H, W) fp32 -> (prepared (..., H*W, Kp) bf16 K-major split per `mode`, rnorm (..., H*W) fp32).

    rnorm = 1 / max(||x||_2, 1e-12) over C (F.normalize's rule); the contraction applies it in its epilogue.
    """
    _lib.require_cuda(feats)
    lib = _lib.load()
    mid = _mode_id(mode)
    feats = _as_f32(feats)
    *lead, Cc, H, W = feats.shape
    G = 1
    for d in lead:
        G *= d
    P = H * W
    kp = lib.pp_match_kp(Cc, mid)
    if kp <= 0:
        raise RuntimeError(f"picopose_b200: bad feature dim {Cc}")
    out = torch.empty((*lead, P, kp), dtype=torch.bfloat16, device=feats.device)
    rnorm = torch.empty((*lead, P), dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        _lib.check(lib.pp_match_prepare(_lib.ptr(feats), G, Cc, P, mid, int(is_query), _lib.ptr(out), _lib.ptr(rnorm),
                                        _lib.stream_of(feats)), "pp_match_prepare")
    return out, rnorm


class TemplateBank:
    """Template banks prepared once: (n_banks, N, C, H, W) fp32 -> resident (n_banks, N, H*W, Kp) bf16 + inverse norms."""

    def __init__(self, prepared: torch.Tensor, rnorm: torch.Tensor, C: int, H: int, W: int, mode: str):
        self.prepared = prepared
        self.rnorm = rnorm
        self.C, self.H, self.W, self.mode = C, H, W, mode

    @classmethod
    def from_features(cls, src_feats: torch.Tensor, mode: Optional[str] = None) -> "TemplateBank":
        if src_feats.dim() == 4:
            src_feats = src_feats.unsqueeze(0)
        if src_feats.dim() != 5:
            raise ValueError("expected (n_banks, N, C, H, W) template features")
        mode = default_mode() if mode is None else mode
        _, _, Cc, H, W = src_feats.shape
        prep, rnorm = prepare_features(src_feats, mode, is_query=False)
        return cls(prep, rnorm, Cc, H, W, mode)

    @property
    def n_banks(self) -> int:
        return self.prepared.shape[0]

    @property
    def n_views(self) -> int:
        return self.prepared.shape[1]

    @property
    def device(self):
        return self.prepared.device

    def view_slice(self, start: int, stop: int) -> "TemplateBank":
        """Bank restricted to views [start, stop) (template-axis sharding); copies to keep rows dense."""
        return TemplateBank(self.prepared[:, start:stop].contiguous(), self.rnorm[:, start:stop].contiguous(),
                            self.C, self.H, self.W, self.mode)


def _resolve_bank(src_feats, mode, B):
    """-> (TemplateBank, bank_of_det or None).  Accepts a TemplateBank, a dense (B,N,C,H,W) tensor, or a
    batch-expanded view (stride 0 over B) of one shared bank."""
    if isinstance(src_feats, TemplateBank):
        return src_feats, None
    from .serving import BankHandle
    if isinstance(src_feats, BankHandle):            # resident banks travelling through the reference's own loop
        return src_feats.bank, src_feats.bank_index
    if src_feats.dim() != 5:
        raise ValueError("src_feats must be (B, N, C, H, W)")
    if src_feats.shape[0] > 1 and src_feats.stride(0) == 0:
        bank = TemplateBank.from_features(src_feats[:1], mode)
        return bank, torch.zeros(B, dtype=torch.int32, device=src_feats.device)
    return TemplateBank.from_features(src_feats, mode), None


class _on_device:
    """`with torch.cuda.device(dev)` only when dev is not already current (the context manager costs microseconds)."""

    def __init__(self, dev):
        self.ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _checked_index(bank_index, src_feats):
    """Range check of `bank_index` where it costs nothing: an index that lives on the host is validated here; one that
    lives on the device is clamped and reported by the contraction kernel (fault record, `_lib.check_device_faults`)."""
    if bank_index is None or bank_index.is_cuda or bank_index.numel() == 0:
        return bank_index
    from .serving import BankHandle
    if isinstance(src_feats, BankHandle):
        n_banks = src_feats.bank.n_banks
    elif isinstance(src_feats, TemplateBank):
        n_banks = src_feats.n_banks
    else:
        n_banks = src_feats.shape[0]
    lo, hi = int(bank_index.min()), int(bank_index.max())
    if lo < 0 or hi >= n_banks:
        raise IndexError(f"bank_index out of range: values span [{lo}, {hi}] for {n_banks} banks")
    return bank_index


def _check_bank(bank, bank_index, tar_feat):
    B, Cc, H, W = tar_feat.shape
    if H != W:
        raise AssertionError("matching_templates expects a square patch grid (H == W)")
    if bank.device != tar_feat.device:
        raise RuntimeError("template bank and query features live on different devices")
    if (bank.C, bank.H, bank.W) != (Cc, H, W):
        raise ValueError(f"bank features {(bank.C, bank.H, bank.W)} do not match the query {(Cc, H, W)}")
    if bank_index is None and bank.n_banks != B:
        raise ValueError(f"{bank.n_banks} banks for {B} detections: pass bank_index")
    if bank_index is not None:
        if bank_index.numel() != B:
            raise ValueError(f"bank_index names {bank_index.numel()} banks for {B} detections")
        bank_index = bank_index.to(device=tar_feat.device, dtype=torch.int32).contiguous()
    return bank_index


def _match_one_call(bank, tar_feat, tar_mask, bank_index, k, cluster, want_sim):
    """Single library call: query prologue + contraction + finalisation (+ top-k).  None if the batch must be chunked."""
    lib = _lib.load()
    B, Cc, H, W = tar_feat.shape
    N, T = bank.n_views, H * W
    mid = _mode_id(bank.mode)
    need = lib.pp_match_templates_workspace(B, N, Cc, H, W, mid)
    if B > _MAX_DETS_PER_LAUNCH or lib.pp_match_scores_workspace(B, N, T) > _WORKSPACE_LIMIT:
        return None
    dev = tar_feat.device
    feat, mask = _as_f32(tar_feat), _as_f32(tar_mask)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    score = torch.empty(B, k, dtype=torch.float32, device=dev) if k else None
    idx = torch.empty(B, k, dtype=torch.int64, device=dev) if k else None
    sim = torch.empty(B, N, dtype=torch.float32, device=dev) if want_sim else None
    cl = default_cluster() if cluster is None else cluster
    with _on_device(dev):
        _lib.check(lib.pp_match_templates(
            feat.data_ptr(), mask.data_ptr(), bank.prepared.data_ptr(), bank.rnorm.data_ptr(), bank.n_banks,
            _lib.ptr(bank_index), B, N, Cc, H, W, mask.shape[-2], mask.shape[-1], mid, k, _lib.ptr(score), _lib.ptr(idx),
            _lib.ptr(sim), ws.data_ptr(), need, cl, _lib.stream_of(tar_feat)), "pp_match_templates")
    return score, idx, sim


def _match_dense_call(src_feats, tar_feat, tar_mask, mode, k, cluster, want_sim, bank_index=None):
    """matching_templates on dense fp32 template features in ONE library call (bank prologue and query prologue run
    concurrently inside it).  None if the shapes need the chunked path."""
    lib = _lib.load()
    B, Cc, H, W = tar_feat.shape
    G, N = src_feats.shape[0], src_feats.shape[1]
    T = H * W
    mode = default_mode() if mode is None else mode
    mid = _mode_id(mode)
    need = lib.pp_match_templates_workspace(B, N, Cc, H, W, mid)
    if B > _MAX_DETS_PER_LAUNCH or lib.pp_match_scores_workspace(B, N, T) > _WORKSPACE_LIMIT or G * N > 65535:
        return None
    dev = tar_feat.device
    src, feat, mask = _as_f32(src_feats), _as_f32(tar_feat), _as_f32(tar_mask)
    kp = lib.pp_match_kp(Cc, mid)
    prep = torch.empty(G, N, T, kp, dtype=torch.bfloat16, device=dev)
    rnorm = torch.empty(G, N, T, dtype=torch.float32, device=dev)
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    score = torch.empty(B, k, dtype=torch.float32, device=dev) if k else None
    idx = torch.empty(B, k, dtype=torch.int64, device=dev) if k else None
    sim = torch.empty(B, N, dtype=torch.float32, device=dev) if want_sim else None
    if bank_index is not None:
        if bank_index.numel() != B:
            raise ValueError(f"bank_index names {bank_index.numel()} banks for {B} detections")
        bank_of_det = bank_index.to(device=dev, dtype=torch.int32).contiguous()
    else:
        bank_of_det = torch.zeros(B, dtype=torch.int32, device=dev) if (G == 1 and B > 1) else None
    cl = default_cluster() if cluster is None else cluster
    with _on_device(dev):
        _lib.check(lib.pp_match_templates_dense(
            src.data_ptr(), G, feat.data_ptr(), mask.data_ptr(), _lib.ptr(bank_of_det), B, N, Cc, H, W, mask.shape[-2],
            mask.shape[-1], mid, k, prep.data_ptr(), rnorm.data_ptr(), _lib.ptr(score), _lib.ptr(idx), _lib.ptr(sim),
            ws.data_ptr(), need, cl, _lib.stream_of(tar_feat)), "pp_match_templates_dense")
    return score, idx, sim


def _dense_features(src_feats, tar_feat, bank_index=None):
    """The reference's own call shape: a plain (B | 1-expanded, N, C, H, W) tensor on the query's device (any number of
    banks when `bank_index` maps detections to them) -> the tensor to prepare ((1, ...) for a stride-0 batch view),
    else None (banks, handles and odd shapes take the general path)."""
    from .serving import BankHandle
    if not isinstance(src_feats, torch.Tensor) or isinstance(src_feats, BankHandle) or src_feats.dim() != 5:
        return None
    B, Cc, H, W = tar_feat.shape
    if src_feats.device != tar_feat.device or tuple(src_feats.shape[2:]) != (Cc, H, W) or H != W:
        return None
    if bank_index is not None:
        return src_feats if src_feats.stride(0) != 0 else None
    if src_feats.shape[0] > 1 and src_feats.stride(0) == 0:
        return src_feats[:1]
    return src_feats if src_feats.shape[0] == B else None


def template_scores(src_feats, tar_feat: torch.Tensor, tar_mask: torch.Tensor, *, mode: Optional[str] = None,
                    bank_index: Optional[torch.Tensor] = None, want_indices: bool = False,
                    want_mutual: bool = False, cluster: Optional[int] = None):
    """Dense scores sim_avg (B, N) of utils/matching.py:38-67.

    With want_indices also returns (score_t2s (B,N,T) f32, idx_t2s (B,N,T) i32, idx_s2t (B,N,T) i32): the
    bidirectional nearest-neighbour patch correspondences the reference computes at :50-51.
    want_mutual appends mutual_nn (B,N,T) uint8: 1 where query patch t and template patch idx_t2s[t] pick each
    other (an extra the reference does not compute).  `bank_index` (B,) maps detections to banks of a shared
    TemplateBank.
    """
    _lib.require_cuda(tar_feat, tar_mask)
    _lib.require_inference("template_scores", src_feats, tar_feat)
    lib = _lib.load()
    B, Cc, H, W = tar_feat.shape
    bank_index = _checked_index(bank_index, src_feats)
    if not want_indices and not want_mutual:
        dense = _dense_features(src_feats, tar_feat, bank_index)
        if dense is not None:
            out = _match_dense_call(dense, tar_feat, tar_mask, mode, 0, cluster, True, bank_index)
            if out is not None:
                return out[2]
    bank, auto_index = _resolve_bank(src_feats, mode, B)
    bank_index = _check_bank(bank, auto_index if bank_index is None else bank_index, tar_feat)
    if not want_indices and not want_mutual:
        out = _match_one_call(bank, tar_feat, tar_mask, bank_index, 0, cluster, True)
        if out is not None:
            return out[2]
    N, T = bank.n_views, H * W
    dev = tar_feat.device
    mid = _mode_id(bank.mode)
    kp = lib.pp_match_kp(Cc, mid)
    feat = _as_f32(tar_feat)
    mask = _as_f32(tar_mask)
    Hm, Wm = mask.shape[-2:]
    sim_avg = torch.empty(B, N, dtype=torch.float32, device=dev)
    sc = it = is_ = mu = None
    if want_mutual:
        want_indices = True
        mu = torch.empty(B, N, T, dtype=torch.uint8, device=dev)
    if want_indices:
        sc = torch.empty(B, N, T, dtype=torch.float32, device=dev)
        it = torch.empty(B, N, T, dtype=torch.int32, device=dev)
        is_ = torch.empty(B, N, T, dtype=torch.int32, device=dev)
    cl = default_cluster() if cluster is None else cluster
    if mid == _lib.MODE_BF16:
        cl |= _lib.MATCH_FAST_KEYS          # cheaper reduction keys, resolution 7.6e-6 (include/picopose_b200.h)
    # bound the scratch (two 64-bit keys per (b, n, t)) by slicing the detection batch
    per_det = max(1, lib.pp_match_scores_workspace(1, N, T))
    chunk = max(1, min(B, _WORKSPACE_LIMIT // per_det, _MAX_DETS_PER_LAUNCH))
    ws = torch.empty(lib.pp_match_scores_workspace(chunk, N, T), dtype=torch.uint8, device=dev)
    q = torch.empty(chunk, T, kp, dtype=torch.bfloat16, device=dev)
    q_rn = torch.empty(chunk, T, dtype=torch.float32, device=dev)
    q_meta = torch.empty(lib.pp_match_query_meta_bytes(chunk, T), dtype=torch.uint8, device=dev)
    with _on_device(dev):
        st = _lib.stream_of(tar_feat)
        for b0 in range(0, B, chunk):
            b1 = min(B, b0 + chunk)
            nb = b1 - b0
            # query side: mask resize + compaction of the unmasked patches + cast/transposition + inverse norms
            _lib.check(lib.pp_match_prepare_query(_lib.ptr(feat[b0:b1]), _lib.ptr(mask[b0:b1]), nb, Cc, H, W, Hm, Wm, mid,
                                                  _lib.ptr(q), _lib.ptr(q_rn), _lib.ptr(q_meta), st),
                       "pp_match_prepare_query")
            if bank_index is not None:
                bidx, n_banks = _lib.ptr(bank_index[b0:b1]), bank.n_banks
                bank_ptr, bank_rn = _lib.ptr(bank.prepared), _lib.ptr(bank.rnorm)
            else:
                bidx, n_banks = 0, nb
                bank_ptr, bank_rn = _lib.ptr(bank.prepared[b0:b1]), _lib.ptr(bank.rnorm[b0:b1])
            _lib.check(lib.pp_match_scores(
                _lib.ptr(q), _lib.ptr(q_rn), _lib.ptr(q_meta), bank_ptr, bank_rn, n_banks, bidx, nb, N, H, W, kp,
                _lib.ptr(sim_avg[b0:b1]), _lib.ptr(sc[b0:b1]) if want_indices else 0,
                _lib.ptr(it[b0:b1]) if want_indices else 0, _lib.ptr(is_[b0:b1]) if want_indices else 0,
                _lib.ptr(mu[b0:b1]) if want_mutual else 0, _lib.ptr(ws), ws.numel(), cl, st), "pp_match_scores")
    if want_mutual:
        return sim_avg, sc, it, is_, mu
    if want_indices:
        return sim_avg, sc, it, is_
    return sim_avg


def topk_scores(sim_avg: torch.Tensor, k: int, idx_offset: int = 0):
    """torch.topk(sim_avg, k, dim=1) replacement (sorted descending, int64 indices)."""
    _lib.require_cuda(sim_avg)
    lib = _lib.load()
    sim_avg = _as_f32(sim_avg)
    B, N = sim_avg.shape
    if k > N:
        raise RuntimeError(f"selected index k out of range (k={k}, N={N})")
    score = torch.empty(B, k, dtype=torch.float32, device=sim_avg.device)
    idx = torch.empty(B, k, dtype=torch.int64, device=sim_avg.device)
    with torch.cuda.device(sim_avg.device):
        _lib.check(lib.pp_topk(_lib.ptr(sim_avg), B, N, k, idx_offset, _lib.ptr(score), _lib.ptr(idx),
                               _lib.stream_of(sim_avg)), "pp_topk")
    return score, idx


def matching_templates(src_feats, tar_feat, src_masks, tar_mask, topk=5, *, mode: Optional[str] = None,
                       bank_index: Optional[torch.Tensor] = None):
    """Drop-in for utils/matching.py:29-69.  Returns (pred_score_src (B,k) f32, pred_id_src (B,k) i64).

    `src_masks` is accepted and ignored, as in the reference.  `src_feats` may also
    be a TemplateBank (pre-normalised once per object) with `bank_index` mapping detections to banks.
    """
    _lib.require_cuda(tar_feat, tar_mask)
    _lib.require_inference("matching_templates", src_feats, tar_feat)
    bank_index = _checked_index(bank_index, src_feats)
    dense = _dense_features(src_feats, tar_feat, bank_index)
    if dense is not None:
        if topk > dense.shape[1]:
            raise RuntimeError(f"selected index k out of range (k={topk}, N={dense.shape[1]})")
        out = _match_dense_call(dense, tar_feat, tar_mask, mode, topk, None, False, bank_index)
        if out is not None:
            return out[0], out[1]
    bank, auto_index = _resolve_bank(src_feats, mode, tar_feat.shape[0])
    bank_index = _check_bank(bank, auto_index if bank_index is None else bank_index, tar_feat)
    if topk > bank.n_views:
        raise RuntimeError(f"selected index k out of range (k={topk}, N={bank.n_views})")
    out = _match_one_call(bank, tar_feat, tar_mask, bank_index, topk, None, False)
    if out is not None:
        return out[0], out[1]
    sim_avg = template_scores(bank, tar_feat, tar_mask, bank_index=bank_index)
    return topk_scores(sim_avg, topk)


def matching_features_similarity(src_feat, tar_feat, src_mask, tar_mask, *, mode: Optional[str] = None):
    """Drop-in for utils/matching.py:6-26.  Returns the (B, H*W, H, W) stage-2 similarity volume."""
    _lib.require_cuda(src_feat, tar_feat, src_mask)
    _lib.require_inference("matching_features_similarity", src_feat, tar_feat)
    lib = _lib.load()
    B, Cc, H, W = src_feat.shape
    if H != W:
        raise AssertionError("matching_features_similarity expects a square patch grid (H == W)")
    mid = _mode_id(mode)
    if tuple(tar_feat.shape) != (B, Cc, H, W):
        raise ValueError("src_feat and tar_feat must have the same shape")
    src, tar, mask = _as_f32(src_feat), _as_f32(tar_feat), _as_f32(src_mask)
    Hm, Wm = mask.shape[-2:]
    T = H * W
    out = torch.empty(B, T, H, W, dtype=torch.float32, device=src_feat.device)
    need = lib.pp_match_similarity_dense_workspace(B, Cc, H, W, mid)
    if need == 0 and B > 0:
        raise RuntimeError(f"picopose_b200: bad feature dim {Cc}")
    ws = torch.empty(max(need, 1), dtype=torch.uint8, device=src_feat.device)
    with _on_device(src_feat.device):
        # one prologue launch for both operands + one contraction whose epilogue writes the finished volume
        _lib.check(lib.pp_match_similarity_dense(_lib.ptr(src), _lib.ptr(tar), _lib.ptr(mask), B, Cc, H, W, Hm, Wm, mid,
                                                 _lib.ptr(out), _lib.ptr(ws), need, default_cluster(),
                                                 _lib.stream_of(src_feat)), "pp_match_similarity_dense")
    return out
