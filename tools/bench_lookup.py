"""Stage-3 lookup sweep (BASELINE.json configs[3]): 64x64 maps, batch B, L=1, radius 4..8, fp32 volume.
Prints achieved algorithmic GB/s (SURVEY 8(d) bytes per query) against the measured HBM peak.

    python tools/bench_lookup.py [--batch 256] [--radii 4 5 6 7 8] [--iters 20] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--levels", type=int, default=1)
    ap.add_argument("--radii", type=int, nargs="+", default=[4, 5, 6, 7, 8])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--l2-fetch", type=int, default=0, help="cudaLimitMaxL2FetchGranularity to try (32/64/128), 0 = leave")
    ap.add_argument("--layouts", nargs="+", default=["rowmajor", "tiled"], help="volume layouts to time (rowmajor = reference, tiled = TiledPyramid)")
    ap.add_argument("--once", action="store_true", help="one warm-up and one timed launch per case (for ncu captures)")
    args = ap.parse_args()
    if args.l2_fetch:
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        torch.cuda.init()
        torch.zeros(1, device="cuda:0")
        val = ctypes.c_size_t(0)
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(args.l2_fetch))       # cudaLimitMaxL2FetchGranularity = 0x05
        rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
        print("cudaLimitMaxL2FetchGranularity set rc=%d now=%d" % (rc, val.value), flush=True)
    from picopose_b200.corr_lookup import corr_lookup
    from picopose_b200.correlation import TiledPyramid
    dev = "cuda:0"
    B, H, L = args.batch, args.size, args.levels
    Q = B * H * H
    g = torch.Generator(device=dev).manual_seed(0)
    pyr = [torch.randn(Q, 1, H >> i, H >> i, device=dev, generator=g) for i in range(L)]
    flow = 4.0 * torch.randn(B, 2, H, H, device=dev, generator=g)
    peak = 6534.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p)).get("hbm_gbs", peak))
    rows = []
    vols = {"rowmajor": pyr}
    if "tiled" in args.layouts:
        vols["tiled"] = TiledPyramid.from_volumes(pyr)          # the same values, 4 x 8 tiles per 128-byte line
    iters, warm = (1, 1) if args.once else (args.iters, 3)
    for layout in args.layouts:
        for r in args.radii:
            D = 2 * r + 1
            per_q = 8 + L * D * D * 4 + sum(min((2 * r + 2) ** 2, (H >> i) ** 2) * 4 for i in range(L))
            for _ in range(warm):
                out = corr_lookup(vols[layout], flow, r)
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
            for a, b in ev:
                a.record()
                out = corr_lookup(vols[layout], flow, r)
                b.record()
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in ev)
            med = ts[len(ts) // 2]
            gbs = Q * per_q / (med * 1e-3) / 1e9
            inb = float((out != 0).float().mean())
            if layout == "tiled" and "rowmajor" in args.layouts:
                assert torch.equal(out, corr_lookup(pyr, flow, r)), "tiled and row-major lookups differ"
            rows.append({"layout": layout, "radius": r, "queries": Q, "bytes_per_query": per_q, "ms_median": med, "ms_min": ts[0],
                         "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / peak, "nonzero_fraction_of_outputs": inb,
                         "Mqueries_per_s": Q / med / 1e3})
            print("%-8s r=%d  %.3f ms (min %.3f)  %.0f GB/s algorithmic = %.1f%% of %.0f GB/s   %.0f Mq/s   nonzero outputs %.1f%%"
                  % (layout, r, med, ts[0], gbs, 100 * gbs / peak, peak, Q / med / 1e3, 100 * inb), flush=True)
            del out
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"workload": "corr_lookup B=%d %dx%d L=%d fp32 volume %.1f GB, flow~N(0,16)" % (B, H, H, L, sum(x.numel() for x in pyr) * 4 / 1e9),
                       "hbm_peak_GBps": peak, "rows": rows}, f, indent=1)


if __name__ == "__main__":
    main()
