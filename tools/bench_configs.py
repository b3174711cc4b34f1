"""BASELINE.json configs[2] / configs[4] on the GPUs of one node: B detections x 642 views, 8 shared object banks
(prepared bf16, resident), template axis sharded over the ranks, one all-gather top-k merge per batch.

    python tools/bench_configs.py --detections 64 [--views 642] [--iters 5]          # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_configs.py --detections 512

Banks are drawn directly on the device (torch.randn with a seeded device generator); every detection's query is a
noisy copy of one view of its object, so the expected top-1 is known and checked."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--detections", type=int, default=64)
    ap.add_argument("--views", type=int, default=642)
    ap.add_argument("--objects", type=int, default=8)
    ap.add_argument("--channels", type=int, default=1024)
    ap.add_argument("--grid", type=int, default=32)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    from picopose_b200 import _lib
    from picopose_b200 import matching as M
    from picopose_b200.sharded import ShardedMatcher, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, N, C, H, n_obj = a.detections, a.views, a.channels, a.grid, a.objects
    lo, hi = shard_range(N, rank, world)
    matcher = ShardedMatcher(N)

    # banks: identical on every rank (same seed), each rank keeps and prepares only its view shard, object by object
    preps, rns = [], []
    queries = torch.empty(B, C, H, H, device=dev)
    gsel = torch.Generator().manual_seed(1)
    top1 = torch.randint(0, N, (B,), generator=gsel)
    obj = torch.arange(B) % n_obj
    for o in range(n_obj):
        g = torch.Generator(device=dev).manual_seed(100 + o)
        bank = torch.randn(N, C, H, H, device=dev, generator=g)            # 2.7 GB fp32 at 642 x 1024 x 32^2
        for b in range(B):
            if int(obj[b]) == o:
                gq = torch.Generator(device=dev).manual_seed(1000 + b)
                queries[b] = bank[int(top1[b])] + 0.5 * torch.randn(C, H, H, device=dev, generator=gq)
        p, rn = M.prepare_features(bank[lo:hi].unsqueeze(0))
        preps.append(p)
        rns.append(rn)
        del bank
    bank = M.TemplateBank(torch.cat(preps), torch.cat(rns), C, H, H, M.default_mode())
    del preps, rns
    from picopose_b200 import synth
    mask = synth.disc_mask(B).to(dev)
    bidx = obj.to(device=dev, dtype=torch.int32)

    def run():
        return matcher.match(bank, queries, mask, topk=5, bank_index=bidx)

    score, idx = run()
    _lib.check_device_faults()
    ok = bool((idx[:, 0].cpu() == top1).all())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(a.iters):
        run()
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / a.iters], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    if rank == 0:
        T = H * H
        flops = 2.0 * B * N * T * T * C
        line = {"workload": "%d detections x %d views x %d ch x %dx%d patches, %d shared banks, %d GPU(s)" % (B, N, C, H, H, n_obj, world),
                "ms_per_batch": ms, "detections_per_s": B * 1e3 / ms, "matches_per_s": B * N * T * 1e3 / ms,
                "algorithmic_TFLOPs_all_gpus": flops / ms / 1e9, "top1_recovered": ok, "views_per_rank": hi - lo,
                "bank_bytes_per_rank": bank.prepared.numel() * 2}
        print(json.dumps(line), flush=True)
        if a.json:
            with open(a.json, "w") as f:
                json.dump(line, f, indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
