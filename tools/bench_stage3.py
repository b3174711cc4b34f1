"""Stage-3 correlation step of one FlowDecoder level (model/stage3/flow_decoder.py:59-61):
pyramid + lookup, three ways on the same inputs:
  torch      : the reference's own ops on the GPU (torch.matmul fp32 + AvgPool2d) feeding pp_corr_lookup
  two-step   : pp_correlation_pyramid (tcgen05, fp32-accurate split mode) + pp_corr_lookup
  fused      : pp_windowed_correlation (no all-pairs volume)

    python tools/bench_stage3.py [--batch 4] [--iters 20] [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--radius", type=int, default=2)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    from picopose_b200.corr_lookup import corr_lookup
    from picopose_b200.correlation import correlation_pyramid, windowed_correlation
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []
    for (H, L) in ((16, 1), (32, 2), (64, 3)):
        N, C, r = a.batch, 256, a.radius
        f1 = torch.randn(N, C, H, H, device=dev, generator=g)
        f2 = torch.randn(N, C, H, H, device=dev, generator=g)
        flow = 2.0 * torch.randn(N, 2, H, H, device=dev, generator=g)
        pool = torch.nn.AvgPool2d(2, 2)

        def torch_pyr():
            corr = torch.matmul(f1.view(N, C, -1).permute(0, 2, 1), f2.view(N, C, -1)).view(N * H * H, 1, H, H) / 16.0
            pyr = [corr]
            for _ in range(L - 1):
                pyr.append(pool(pyr[-1]))
            return pyr

        ref = corr_lookup(torch_pyr(), flow, r)
        two = corr_lookup(correlation_pyramid(f1, f2, L), flow, r)
        fused = windowed_correlation(f1, f2, flow, L, r)
        e2, ef = float((two - ref).abs().max()), float((fused - ref).abs().max())
        t_torch = timed(lambda: corr_lookup(torch_pyr(), flow, r), a.iters)
        t_two = timed(lambda: corr_lookup(correlation_pyramid(f1, f2, L), flow, r), a.iters)
        t_two_bf16 = timed(lambda: corr_lookup(correlation_pyramid(f1, f2, L, mode="bf16"), flow, r), a.iters)
        t_fused = timed(lambda: windowed_correlation(f1, f2, flow, L, r), a.iters)
        t_kernel = {}
        for kern in ("direct", "tiled"):
            os.environ["PICOPOSE_WCORR_KERNEL"] = kern
            try:
                if float((windowed_correlation(f1, f2, flow, L, r) - ref).abs().max()) > 1e-4:
                    raise AssertionError("kernel %s disagrees with the two-step path" % kern)
                t_kernel[kern] = timed(lambda: windowed_correlation(f1, f2, flow, L, r), a.iters)
            except RuntimeError:
                t_kernel[kern] = None
        del os.environ["PICOPOSE_WCORR_KERNEL"]
        smooth = flow.mean(dim=(2, 3), keepdim=True) + 0.25 * flow          # a smooth field: what stage 2 hands over
        t_smooth = timed(lambda: windowed_correlation(f1, f2, smooth, L, r), a.iters)
        row = {"level": "%dx%d, L=%d, r=%d, N=%d, C=%d" % (H, H, L, r, N, C), "torch_matmul_plus_lookup_ms": t_torch,
               "pyramid_fp32mode_plus_lookup_ms": t_two, "pyramid_bf16_plus_lookup_ms": t_two_bf16, "fused_ms": t_fused,
               "fused_direct_kernel_ms": t_kernel["direct"], "fused_tiled_kernel_ms": t_kernel["tiled"], "fused_smooth_flow_ms": t_smooth,
               "max_abs_err_two_step_vs_torch": e2, "max_abs_err_fused_vs_torch": ef,
               "volume_bytes_avoided": sum(N * H * H * (H >> i) * (H >> i) * 4 for i in range(L))}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
