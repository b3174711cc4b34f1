"""The configs[1] step of bench.py (dense matching_templates on fp32 template features + the lookup ladder) launched
eagerly (7 library calls from Python per step) against the same step captured once into a CUDA graph and replayed.

Answers two questions on one GPU: (1) is the library capture-safe (no allocation, synchronisation or host round trip
inside a call after its first use), (2) what does replay buy -- device time per step and host time per step.
    python tools/bench_graph.py [--steps 200] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from picopose_b200 import _lib  # noqa: E402
from picopose_b200 import matching as M  # noqa: E402
from picopose_b200.corr_lookup import CorrLookup  # noqa: E402


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h0 = time.perf_counter()
    for _ in range(steps):
        fn()
    host = (time.perf_counter() - h0) * 1e3 / steps
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, host


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    _lib.load()
    src, tar, mask, lookups, planted, _, _ = bench.make_inputs(1, 0)
    src_d, tar_d, mask_d = src.to(dev), tar.to(dev), mask.to(dev)
    look_d = [([p.to(dev) for p in pyr], flow.to(dev)) for pyr, flow in lookups]
    lookup_mod = CorrLookup(radius=bench.CFG["radius"])
    k = bench.CFG["topk"]

    def step():
        score, idx = M.matching_templates(src_d, tar_d, None, mask_d, topk=k)
        return score, idx, [lookup_mod(pyr, flow) for pyr, flow in look_d]

    for _ in range(5):
        ref = step()
    _lib.check_device_faults()
    eager_ms, eager_host = timed(step, args.steps)

    out = {"workload": "configs[1] step as in bench.py (1 x 162 x 1024 x 32^2 + lookup ladder), 1 GPU", "steps": args.steps,
           "eager": {"ms_per_step": eager_ms, "host_ms_per_step": eager_host}}
    side = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            got = step()
        for _ in range(5):
            graph.replay()
        torch.cuda.synchronize()
        _lib.check_device_faults()
        same = (torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
                and all(torch.equal(a, b) for a, b in zip(got[2], ref[2])))
        ms, host = timed(graph.replay, args.steps)
        out["graph"] = {"ms_per_step": ms, "host_ms_per_step": host, "outputs_equal_eager": bool(same),
                        "top5": got[1][0].tolist(), "planted": planted[0, :k].tolist()}
    except Exception as exc:                                   # noqa: BLE001 -- the finding is the message
        out["graph"] = {"capture_failed": "%s: %s" % (type(exc).__name__, str(exc)[:400])}
    print(json.dumps(out))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
