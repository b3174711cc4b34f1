"""Multi-GPU stage-1 matching: template-axis sharding + one all-gather top-k merge.

One process per GPU (torchrun).  Rank g keeps views [lo_g, hi_g) of every object's template bank
(prepared bf16, resident), all ranks see the whole detection batch, each computes its local
sim_avg[:, lo_g:hi_g] and a local top-k with GLOBAL view indices, and a single NCCL all-gather of the
(B, k) score / index pairs over NVLink is merged identically on every rank.  Every (detection, view)
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


def _select_cuda(scores: torch.Tensor, k: int):
    from .matching import topk_scores
    return topk_scores(scores, k)


def merge_topk(local_score: torch.Tensor, local_idx: torch.Tensor, k: int, group=None,
               select: Optional[Callable] = None):
    """All-gathers per-rank candidates (B, k_local) and keeps the global top-k (sorted, ties -> lowest rank/slot).

    Ranks whose shard holds fewer than k views pad with (-inf, -1).  `select(scores, k) -> (values, positions)`
    defaults to the library's top-k kernel; tests on CPU/gloo inject a torch.topk-based one.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    B, kl = local_score.shape
    if kl < k:
        pad_s = torch.full((B, k - kl), float("-inf"), dtype=local_score.dtype, device=local_score.device)
        pad_i = torch.full((B, k - kl), -1, dtype=local_idx.dtype, device=local_idx.device)
        local_score = torch.cat([local_score, pad_s], dim=1)
        local_idx = torch.cat([local_idx, pad_i], dim=1)
    local_score = local_score.contiguous()
    local_idx = local_idx.contiguous()
    if world == 1:
        all_s, all_i = local_score, local_idx
    else:
        # rank-major concatenation along dim 0 (the layout both NCCL and gloo accept)
        gs = torch.empty(world * B, k, dtype=local_score.dtype, device=local_score.device)
        gi = torch.empty(world * B, k, dtype=local_idx.dtype, device=local_idx.device)
        dist.all_gather_into_tensor(gs, local_score, group=group)
        dist.all_gather_into_tensor(gi, local_idx, group=group)
        all_s = gs.view(world, B, k).permute(1, 0, 2).reshape(B, world * k)
        all_i = gi.view(world, B, k).permute(1, 0, 2).reshape(B, world * k)
    select = select or _select_cuda
    val, pos = select(all_s.contiguous(), k)
    return val, torch.gather(all_i, 1, pos)


class ShardedMatcher:
    """Template-sharded `matching_templates` over a process group."""

    def __init__(self, n_views: int, group=None, select: Optional[Callable] = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_views = n_views
        self.lo, self.hi = shard_range(n_views, self.rank, self.world)
        self.select = select
        self.bank = None

    def load_bank(self, src_feats_shard: torch.Tensor, mode: Optional[str] = None):
        """src_feats_shard: this rank's views (n_banks, hi-lo, C, H, W) fp32 -> resident prepared bank."""
        from .matching import TemplateBank
        assert src_feats_shard.shape[1] == self.hi - self.lo
        self.bank = TemplateBank.from_features(src_feats_shard, mode)
        return self.bank

    def match(self, src, tar_feat, tar_mask, topk=5, bank_index=None, mode=None):
        """src: this rank's shard, a TemplateBank or raw (B|n_banks, hi-lo, C, H, W) features."""
        from .matching import template_scores, topk_scores
        src = self.bank if src is None else src
        sim = template_scores(src, tar_feat, tar_mask, mode=mode, bank_index=bank_index)     # (B, hi-lo)
        kl = min(topk, sim.shape[1])
        if kl > 0:
            s, i = topk_scores(sim, kl, idx_offset=self.lo)
        else:
            s = sim.new_empty(sim.shape[0], 0)
            i = torch.empty(sim.shape[0], 0, dtype=torch.int64, device=sim.device)
        if self.world == 1 and kl == topk:
            return s, i                      # single rank: the local top-k is the answer, nothing to exchange
        return merge_topk(s, i, topk, self.group, self.select)
