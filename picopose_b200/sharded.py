"""Multi-GPU stage-1 matching: template-axis sharding + one all-gather top-k merge.

One process per GPU (torchrun).  Rank g keeps views [lo_g, hi_g) of every object's template bank
(prepared bf16, resident), all ranks see the whole detection batch (a rank's own detections come from its
host, the others' over NVLink: `ShardedMatcher.gather_queries`), each computes its local
sim_avg[:, lo_g:hi_g] and a local top-k with GLOBAL view indices packed as one (B, k, 2) tensor, and a single
NCCL all-gather of those pairs over NVLink is merged identically on every rank by one small kernel.  Every (detection, view)
score is independent (utils/matching.py:47-67); only topk (:68) couples views, hence one exchange.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_views: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, as-even-as-possible split: the first n_views % world ranks get one extra view."""
    base, extra = divmod(n_views, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def topk_pairs(sim: torch.Tensor, k: int, idx_offset: int = 0) -> torch.Tensor:
    """Local top-k of (B, N_local) scores as one (B, k, 2) float64 tensor of (score, global index) pairs,
    padded with (-inf, -1) when the shard has fewer than k views."""
    from . import _lib
    _lib.require_cuda(sim)
    lib = _lib.load()
    sim = sim.float().contiguous()
    B, N = sim.shape
    out = torch.empty(B, k, 2, dtype=torch.float64, device=sim.device)
    if N == 0:
        out[..., 0] = float("-inf")
        out[..., 1] = -1
        return out
    with torch.cuda.device(sim.device):
        _lib.check(lib.pp_topk_pairs(_lib.ptr(sim), B, N, k, idx_offset, _lib.ptr(out), _lib.stream_of(sim)),
                   "pp_topk_pairs")
    return out


def _merge_cuda(gathered: torch.Tensor, k: int):
    """gathered (R, B, k_in, 2) float64 on a CUDA device -> (score (B,k) f32, idx (B,k) i64)."""
    from . import _lib
    lib = _lib.load()
    R, B, k_in, _ = gathered.shape
    score = torch.empty(B, k, dtype=torch.float32, device=gathered.device)
    idx = torch.empty(B, k, dtype=torch.int64, device=gathered.device)
    with torch.cuda.device(gathered.device):
        _lib.check(lib.pp_topk_merge(_lib.ptr(gathered), R, B, k_in, k, _lib.ptr(score), _lib.ptr(idx),
                                     _lib.stream_of(gathered)), "pp_topk_merge")
    return score, idx


def _merge_torch(gathered: torch.Tensor, k: int):
    """Same selection rule in plain torch ops; used by the CPU/gloo tests of the exchange plumbing only."""
    R, B, k_in, _ = gathered.shape
    flat = gathered.permute(1, 0, 2, 3).reshape(B, R * k_in, 2)
    order = torch.argsort(flat[..., 0], dim=1, descending=True, stable=True)[:, :k]   # stable: lowest rank/slot first
    picked = torch.gather(flat, 1, order.unsqueeze(-1).expand(B, k, 2))
    return picked[..., 0].float(), picked[..., 1].long()


def merge_topk(local_pairs: torch.Tensor, k: int, group=None, merge: Optional[Callable] = None):
    """One all-gather of every rank's (B, k, 2) candidate pairs, then the same merge on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    B, k_in, _ = local_pairs.shape
    local_pairs = local_pairs.contiguous()
    if world == 1:
        gathered = local_pairs.unsqueeze(0)
    else:
        # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept)
        flat = torch.empty(world * B, k_in, 2, dtype=local_pairs.dtype, device=local_pairs.device)
        dist.all_gather_into_tensor(flat, local_pairs, group=group)
        gathered = flat.view(world, B, k_in, 2)
    merge = merge or (_merge_cuda if local_pairs.is_cuda else _merge_torch)
    return merge(gathered, k)


class ShardedMatcher:
    """Template-sharded `matching_templates` over a process group."""

    def __init__(self, n_views: int, group=None, merge: Optional[Callable] = None, query_group=None):
        self.group = group
        # gather_queries on its own communicator: collectives of one NCCL communicator run in issue order on one
        # internal stream, so a query gather for step i+1 would otherwise queue behind step i's top-k exchange
        self.query_group = query_group if query_group is not None else group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_views = n_views
        self.lo, self.hi = shard_range(n_views, self.rank, self.world)
        self.merge = merge
        self.bank = None

    def load_bank(self, src_feats_shard: torch.Tensor, mode: Optional[str] = None):
        """src_feats_shard: this rank's views (n_banks, hi-lo, C, H, W) fp32 -> resident prepared bank."""
        from .matching import TemplateBank
        assert src_feats_shard.shape[1] == self.hi - self.lo
        self.bank = TemplateBank.from_features(src_feats_shard, mode)
        return self.bank

    def gather_queries(self, tar_local: torch.Tensor, mask_local: torch.Tensor, out=None):
        """Each rank received its own detections' query features (b, C, H, W) and masks (b, Hm, Wm) from its host;
        every rank needs the whole batch (template-axis sharding), so the queries travel GPU-to-GPU: two
        all-gathers over NVLink instead of every rank uploading world x the data over PCIe.
        -> (tar (world*b, C, H, W), mask (world*b, Hm, Wm)), rank-major; `out` = optional preallocated pair."""
        if self.world == 1:
            return tar_local, mask_local
        tar_local, mask_local = tar_local.contiguous(), mask_local.contiguous()
        if out is None:
            out = (tar_local.new_empty((self.world * tar_local.shape[0],) + tuple(tar_local.shape[1:])),
                   mask_local.new_empty((self.world * mask_local.shape[0],) + tuple(mask_local.shape[1:])))
        dist.all_gather_into_tensor(out[0], tar_local, group=self.query_group)
        dist.all_gather_into_tensor(out[1], mask_local, group=self.query_group)
        return out

    def match(self, src, tar_feat, tar_mask, topk=5, bank_index=None, mode=None):
        """src: this rank's shard, a TemplateBank or raw (B|n_banks, hi-lo, C, H, W) features."""
        from .matching import matching_templates, template_scores
        src = self.bank if src is None else src
        if self.world == 1:
            # single rank: nothing to exchange, one library call ranks the whole bank
            return matching_templates(src, tar_feat, None, tar_mask, topk, mode=mode, bank_index=bank_index)
        sim = template_scores(src, tar_feat, tar_mask, mode=mode, bank_index=bank_index)     # (B, hi-lo)
        return merge_topk(topk_pairs(sim, topk, idx_offset=self.lo), topk, self.group, self.merge)
