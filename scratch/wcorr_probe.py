import os, sys
sys.path.insert(0, "/root/repo")
import torch
from picopose_b200.correlation import windowed_correlation
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
N, C, H, L, r = 16, 256, 64, 3, 2
f1 = torch.randn(N, C, H, H, device=dev, generator=g)
f2 = torch.randn(N, C, H, H, device=dev, generator=g)
flow = 2.0 * torch.randn(N, 2, H, H, device=dev, generator=g)
for kern in ("direct", "tiled"):
    os.environ["PICOPOSE_WCORR_KERNEL"] = kern
    for _ in range(2):
        windowed_correlation(f1, f2, flow, L, r)
os.environ["PICOPOSE_WCORR_SMEM_F1"] = "1"
os.environ["PICOPOSE_WCORR_KERNEL"] = "direct"
windowed_correlation(f1, f2, flow, L, r)
torch.cuda.synchronize()
