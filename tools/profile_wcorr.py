"""Runs the fused stage-3 correlation with each kernel (per-query, TMA-tiled) on a scattered and on a smooth flow field at N=16, 64x64, L=3, r=2: the workload
behind profiles/r1p_prof_wcorr_*.  Meant to be run under ncu:

    ncu --set full --clock-control none --import-source on -k regex:"windowed_corr|wcorr_prepare" -o out python tools/profile_wcorr.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from picopose_b200.correlation import windowed_correlation  # noqa: E402
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
N, C, H, L, r = 16, 256, 64, 3, 2
f1 = torch.randn(N, C, H, H, device=dev, generator=g)
f2 = torch.randn(N, C, H, H, device=dev, generator=g)
flow = 2.0 * torch.randn(N, 2, H, H, device=dev, generator=g)
smooth = flow.mean(dim=(2, 3), keepdim=True) + 0.25 * flow      # what stage 2 hands over: a smooth field
for kern in ("direct", "tiled"):
    os.environ["PICOPOSE_WCORR_KERNEL"] = kern
    for fl in (flow, smooth):
        windowed_correlation(f1, f2, fl, L, r)
torch.cuda.synchronize()
