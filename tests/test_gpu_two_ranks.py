"""Two real ranks (two processes, two GPUs) over CUDA IPC peer memory: the top-k exchange kernel and the copy-engine
query gather between processes, as `bench.py --gpus 2` uses them.  Skipped on a one-GPU box (the protocols are also
covered there by the simulated-rank tests in test_gpu_parity.py); run with `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_two_ranks.py -m gpu`."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # object collectives only; data moves over NVLink
    try:
        from picopose_b200 import _lib
        from picopose_b200.sharded import PeerExchange, PeerGather, ShardedMatcher, shard_range
        ok = True
        B, N, k = 6, 21, 5
        lo, hi = shard_range(N, rank, world)
        xchg = PeerExchange(max_b=8, k_max=8, device=dev)
        for epoch in range(3):                                       # both parities, and a reuse of the first
            g = torch.Generator().manual_seed(50 + epoch)
            full = torch.randn(B, N, generator=g)
            full[:, 10] = full[:, 3]                                 # tie between rank 0's view 3 and rank 1's view 10
            score, idx = xchg.exchange(full[:, lo:hi].contiguous().to(dev), k, idx_offset=lo)
            torch.cuda.synchronize(dev)
            ref_s, ref_i = torch.sort(full, dim=1, descending=True, stable=True)
            ok = ok and torch.equal(idx.cpu(), ref_i[:, :k]) and torch.equal(score.cpu(), ref_s[:, :k])
        xchg.close()
        # query gather with the copy engines + flag kernels, then the whole sharded matcher against one full bank
        from picopose_b200 import matching as M
        from picopose_b200 import synth
        src, tar, planted = synth.planted_match_inputs(world, 12, 64, 8, seed=9)      # detection d <-> object d
        mask = synth.disc_mask(world)
        matcher = ShardedMatcher(12)
        lo, hi = matcher.lo, matcher.hi
        tar0, mask0 = tar, mask
        for step in range(4):
            # new data every step (a stale slot of the other parity or of two steps ago would show): the features are
            # rescaled (cosine scores do not move), the mask of the odd steps is the full patch grid
            tar = tar0 * float(step + 1)
            mask = mask0 if step % 2 == 0 else torch.ones_like(mask0)
            t_all, m_all = matcher.gather_queries(tar[rank:rank + 1].to(dev), mask[rank:rank + 1].to(dev))
            ok = ok and matcher.uses_peer_memory
            s, i = matcher.match(src[:, lo:hi].contiguous().to(dev), t_all, m_all, topk=4,
                                 bank_index=torch.arange(world, dtype=torch.int32, device=dev))
            torch.cuda.synchronize(dev)
            ok = ok and torch.equal(t_all.cpu(), tar) and torch.equal(m_all.cpu(), mask)
            s1, i1 = M.matching_templates(src.to(dev), tar.to(dev), None, mask.to(dev), topk=4)     # unsharded, same GPU
            ok = ok and torch.equal(i.cpu(), i1.cpu()) and torch.equal(s.cpu(), s1.cpu())
            ok = ok and (step % 2 == 1 or i.cpu().tolist() == planted[:, :4].tolist())
        # the one-step-ahead contract is enforced on the host
        g = matcher._gather
        g.gather(tar[rank:rank + 1].to(dev), mask[rank:rank + 1].to(dev))
        g.gather(tar[rank:rank + 1].to(dev), mask[rank:rank + 1].to(dev))
        try:
            g.gather(tar[rank:rank + 1].to(dev), mask[rank:rank + 1].to(dev))
            ok = False
        except RuntimeError:
            pass
        torch.cuda.synchronize(dev)
        _lib.check_device_faults()
        matcher.close()
        res = [None] * world
        dist.all_gather_object(res, bool(ok))
        if rank == 0:
            ret["ok"] = all(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node (CUDA IPC peer mappings)")
def test_peer_exchange_and_gather_across_two_processes():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret.get("ok") is True
