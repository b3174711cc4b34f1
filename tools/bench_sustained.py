"""Same-box, same-state comparison of the stage-1 contraction with cuBLAS under sustained load (the regime of
BASELINE configs[2]/[4]): alternating ~2 s phases of (a) 64 detections x V views against 8 shared prepared banks and
(b) torch.matmul bf16 8192^3, with SM clock and board power sampled through NVML during each phase.

    python tools/bench_sustained.py [--views 642] [--seconds 2] [--rounds 2] [--json out.json]
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


class Sampler:
    def __init__(self):
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        self.clk, self.pw, self._stop, self._thr = [], [], threading.Event(), None

    def _poll(self):
        while not self._stop.is_set():
            self.clk.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.pw.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            time.sleep(0.01)

    def __enter__(self):
        self.clk, self.pw = [], []
        self._stop.clear()
        self._thr = threading.Thread(target=self._poll, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join()

    def stats(self):
        c, p = sorted(self.clk), sorted(self.pw)
        return {"sm_mhz_median": c[len(c) // 2] if c else None, "power_w_median": p[len(p) // 2] if p else None,
                "power_w_max": p[-1] if p else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=642)
    ap.add_argument("--detections", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    sys.argv = [sys.argv[0]]
    import bench
    from picopose_b200 import matching as M
    from picopose_b200 import synth
    dev = torch.device("cuda", 0)
    N, C, H, n_obj, B = a.views, 1024, 32, 8, a.detections
    bank, queries, obj, top1, _ = bench._config_banks(dev, 0, N, n_obj, N, C, H, B)
    mask = synth.disc_mask(B).to(dev)
    bidx = obj.to(device=dev, dtype=torch.int32)
    rows_exec, _ = bench.executed_rows(mask[:1].cpu(), H)
    flops_ours = 2.0 * B * N * rows_exec * H * H * C
    x = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    y = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    flops_cublas = 2.0 * 8192 ** 3

    def phase(fn, flops):
        fn()
        torch.cuda.synchronize()
        n, t0 = 0, time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with Sampler() as s:
            e0.record()
            while time.perf_counter() - t0 < a.seconds:
                for _ in range(4):
                    fn()
                    n += 1
                torch.cuda.synchronize()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        d = s.stats()
        d.update({"ms_per_call": ms, "tflops": flops / ms / 1e9})
        return d

    rows = []
    for r in range(a.rounds):
        ours = phase(lambda: M.matching_templates(bank, queries, None, mask, topk=5, bank_index=bidx), flops_ours)
        cub = phase(lambda: torch.matmul(x, y), flops_cublas)
        row = {"round": r, "ours_issued": ours, "cublas_8192": cub, "ratio": ours["tflops"] / cub["tflops"]}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump({"workload": "%d detections x %d views, 8 shared banks, disc mask (issued FLOPs) vs torch.matmul bf16 8192^3, "
                                   "alternating %.0f s phases on one B200" % (B, N, a.seconds), "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
