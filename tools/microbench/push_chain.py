"""How long the push side of an 8-rank query gather takes on the sending GPU (2 GPUs of one box are enough).

One process, two devices: the sender's own buffer on cuda:0 and seven "peer" buffers on cuda:1, so the chain is what rank
r of `bench.py --gpus 8` issues per step -- 8 x 4 MB feature slices + 8 x 4 KB masks + the flag kernel -- timed with CUDA
events on the issuing stream.  Forms: `separate` = pp_xchg_push x 2 + pp_xchg_signal, `fused` = pp_xchg_push_signal
(the same copies and flag kernel behind one library call; the bulk copies on PICOPOSE_B200_PUSH_STREAMS internal streams
if that is > 1).  profiles/r2i_push_chain.md was taken with an earlier `fused` whose flag kernel stored the masks itself.
The stream count is read once per process: run it once per value.
    for n in 1 2 4 8; do PICOPOSE_B200_PUSH_STREAMS=$n python tools/microbench/push_chain.py; done
With --load a 1 GB device copy runs on cuda:0 beside every chain (HBM and copy engines busy; the real step's bank
prologue is an SM kernel, so this is the pessimistic case).  Results: profiles/r2i_push_chain.md."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from picopose_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--load", action="store_true")
    args = ap.parse_args()
    assert torch.cuda.device_count() >= 2, "needs two GPUs"
    lib = _lib.load()
    world, tar_bytes, mask_bytes, flag_bytes = args.world, 4 << 20, 4 << 10, 256
    mask_off = flag_bytes + world * tar_bytes
    total = mask_off + world * mask_bytes
    d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
    own = torch.zeros(total, dtype=torch.uint8, device=d0)
    remote = [torch.zeros(total, dtype=torch.uint8, device=d1) for _ in range(world - 1)]
    remote[0][:4096].copy_(own[:4096])                 # torch enables peer access between the two devices here
    own[:4096].copy_(remote[0][:4096])
    torch.cuda.synchronize(d0), torch.cuda.synchronize(d1)
    torch.cuda.set_device(d0)
    ptrs = [own.data_ptr()] + [r.data_ptr() for r in remote]
    peers_host = (C.c_void_p * world)(*ptrs)
    peers_dev = torch.tensor(ptrs, dtype=torch.int64, device=d0)
    tar = torch.randn(tar_bytes // 4, device=d0)
    mask = torch.ones(mask_bytes // 4, device=d0)
    st = torch.cuda.Stream(device=d0)
    load_a = load_b = None
    if args.load:
        load_a = torch.empty(1 << 28, dtype=torch.float32, device=d0)
        load_b = torch.empty_like(load_a)

    def separate(epoch):
        s = st.cuda_stream
        _lib.check(lib.pp_xchg_push(tar.data_ptr(), tar_bytes, peers_host, world, flag_bytes, s), "push")
        _lib.check(lib.pp_xchg_push(mask.data_ptr(), mask_bytes, peers_host, world, mask_off, s), "push")
        _lib.check(lib.pp_xchg_signal(peers_dev.data_ptr(), 0, 0, world, epoch, s), "signal")

    def fused(epoch):
        _lib.check(lib.pp_xchg_push_signal(tar.data_ptr(), tar_bytes, flag_bytes, mask.data_ptr(), mask_bytes, mask_off,
                                           peers_host, peers_dev.data_ptr(), 0, 0, world, epoch, st.cuda_stream), "push_signal")

    out = {"world": world, "push_streams": os.environ.get("PICOPOSE_B200_PUSH_STREAMS", "default (1)"), "load": args.load}
    for name, fn in (("separate", separate), ("fused", fused)):
        for e in range(1, 6):
            fn(e)
        torch.cuda.synchronize(d0)
        pairs, host = [], 0.0
        for it in range(args.iters):
            if args.load:
                load_b.copy_(load_a, non_blocking=True)            # ~0.35 ms of HBM traffic on the default stream
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            h0 = time.perf_counter()
            fn(10 + it)
            host += time.perf_counter() - h0
            e1.record(st)
            pairs.append((e0, e1))
            st.synchronize()
        torch.cuda.synchronize(d0), torch.cuda.synchronize(d1)
        ms = sorted(a.elapsed_time(b) for a, b in pairs)
        out[name] = {"chain_us_median": 1e3 * ms[len(ms) // 2], "chain_us_p90": 1e3 * ms[int(0.9 * len(ms))],
                     "host_us_per_call": 1e6 * host / args.iters}
        # what arrived: the last payload and slice in the last remote buffer
        got = remote[-1][flag_bytes:flag_bytes + 16].view(torch.float32).cpu()
        assert torch.equal(got, tar[:4].cpu())
        assert remote[-1][mask_off:mask_off + 16].view(torch.float32).cpu().eq(1).all()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
