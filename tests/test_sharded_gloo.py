"""world_size-2 gloo test of the multi-GPU host logic: ragged view split, local top-k with global indices,
all-gather, merge.  The selection kernel is replaced by torch.topk (the CUDA kernel is covered by -m gpu)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _pairs(score, idx, k):
    """(B, kl) local top-k -> (B, k, 2) float64 pairs padded with (-inf, -1), like pp_topk_pairs."""
    B, kl = score.shape
    out = torch.empty(B, k, 2, dtype=torch.float64)
    out[..., 0] = float("-inf")
    out[..., 1] = -1
    out[:, :kl, 0] = score.double()
    out[:, :kl, 1] = idx.double()
    return out


def _worker(rank, world, port, n_views, k, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from picopose_b200.sharded import merge_topk, shard_range
        g = torch.Generator().manual_seed(0)
        full = torch.randn(3, n_views, generator=g)                  # identical "dense scores" on every rank
        lo, hi = shard_range(n_views, rank, world)
        local = full[:, lo:hi]
        kl = min(k, hi - lo)
        s, i = torch.topk(local, kl, dim=1)
        val, idx = merge_topk(_pairs(s, i + lo, k), k)          # CPU tensors -> torch merge, gloo all-gather
        ref_v, ref_i = torch.topk(full, k, dim=1)
        ok = torch.equal(val, ref_v) and torch.equal(idx, ref_i)
        # query exchange: each rank holds its own detection, gather_queries assembles the rank-major batch
        from picopose_b200.sharded import ShardedMatcher
        m = ShardedMatcher(n_views)
        assert (m.lo, m.hi) == (lo, hi)
        tar_all = torch.arange(world * 2 * 3 * 2 * 2, dtype=torch.float32).view(world * 2, 3, 2, 2)
        mask_all = (torch.arange(world * 2 * 4 * 4).view(world * 2, 4, 4) % 3 == 0).float()
        t, mk = m.gather_queries(tar_all[2 * rank:2 * rank + 2], mask_all[2 * rank:2 * rank + 2])
        ok = ok and torch.equal(t, tar_all) and torch.equal(mk, mask_all)
        pre = (torch.empty_like(tar_all), torch.empty_like(mask_all))
        t2, _ = m.gather_queries(tar_all[2 * rank:2 * rank + 2], mask_all[2 * rank:2 * rank + 2], out=pre)
        ok = ok and t2 is pre[0] and torch.equal(pre[0], tar_all) and torch.equal(pre[1], mask_all)
        gathered = [None] * world
        dist.all_gather_object(gathered, (ok, idx.tolist()))
        if rank == 0:
            ret["ok"] = all(g[0] for g in gathered) and all(g[1] == gathered[0][1] for g in gathered)
    finally:
        dist.destroy_process_group()


def _run(n_views, k):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n_views, k, ret), nprocs=2, join=True)
    assert ret.get("ok") is True


def test_merge_topk_world2_even_and_ragged():
    _run(162, 5)
    _run(7, 5)      # ragged: 4 + 3 views, one rank has fewer than k candidates and pads with -inf


def test_merge_topk_single_process_no_group():
    from picopose_b200.sharded import merge_topk
    s = torch.tensor([[0.75, 0.5, 0.25]])
    i = torch.tensor([[4, 2, 7]])
    v, idx = merge_topk(_pairs(s, i, 3), 2)
    assert v.tolist() == [[0.75, 0.5]] and idx.tolist() == [[4, 2]]
