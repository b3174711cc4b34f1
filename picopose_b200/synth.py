"""Seeded synthetic inputs of the shapes BASELINE.json names.

All generators draw on the CPU from ``torch.Generator().manual_seed(seed)`` so
the same tensors are produced here, in the tests and on the GPU box.
"""
from __future__ import annotations

import math

import torch

RHO = (0.85, 0.70, 0.55, 0.40, 0.25)


def disc_mask(B: int, size: int = 224, frac: float = 0.45) -> torch.Tensor:
    """Centred disc of radius frac*size: ~64 % of patches valid, corner patch 0 masked."""
    c = (size - 1) / 2.0
    ys = torch.arange(size, dtype=torch.float32).view(size, 1)
    xs = torch.arange(size, dtype=torch.float32).view(1, size)
    m = ((ys - c) ** 2 + (xs - c) ** 2 <= (frac * size) ** 2).float()
    return m[None].repeat(B, 1, 1).contiguous()


def bernoulli_mask(B: int, size: int, p: float, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(B, size, size, generator=g) < p).float()


def planted_match_inputs(B, N, C, H, seed=0, noise=0.5, dtype=torch.float32):
    """Template bank with a planted top-6 ranking per detection.

    src ~ N(0,1); per detection views p0..p5 are chosen, the query is
    src[b,p0] + noise*N(0,1), and view pj (j=1..5) is mixed towards the query
    with correlation RHO[j-1], so the reference top-5 is (p0..p4) with adjacent
    score gaps ~0.08 >> 1e-3.  Returns (src_feats (B,N,C,H,H), tar_feat
    (B,C,H,H), planted (B,6) int64).
    """
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(B, N, C, H, H, generator=g)
    tar = torch.empty(B, C, H, H)
    n_plant = min(6, N)
    planted = torch.empty(B, n_plant, dtype=torch.long)
    for b in range(B):
        perm = torch.randperm(N, generator=g)[:n_plant]
        planted[b] = perm
        q = src[b, perm[0]].clone()
        tar[b] = q + noise * torch.randn(C, H, H, generator=g)
        for j in range(1, n_plant):
            rho = RHO[j - 1]
            src[b, perm[j]] = rho * q + math.sqrt(1.0 - rho * rho) * src[b, perm[j]]
    return src.to(dtype), tar.to(dtype), planted


def shared_bank_inputs(n_obj, N, C, H, B, seed=0, noise=0.5):
    """Config-3 style: n_obj template banks shared by B detections (obj = b mod n_obj).

    Returns (banks (n_obj,N,C,H,H), tar (B,C,H,H), obj_idx (B,), planted (B,6)).
    Each detection's query is a noisy copy of one view of its object's bank; the
    other planted views are *not* remixed (banks are shared), so only the top-1
    is guaranteed.
    """
    g = torch.Generator().manual_seed(seed)
    banks = torch.randn(n_obj, N, C, H, H, generator=g)
    obj_idx = torch.arange(B) % n_obj
    tar = torch.empty(B, C, H, H)
    top1 = torch.empty(B, dtype=torch.long)
    for b in range(B):
        p = int(torch.randint(0, N, (1,), generator=g))
        top1[b] = p
        tar[b] = banks[obj_idx[b], p] + noise * torch.randn(C, H, H, generator=g)
    return banks, tar, obj_idx, top1


def lookup_inputs(B, H, L, seed=0, flow_sigma=4.0, W=None, vol_dtype=torch.float32):
    """Correlation pyramid (L levels, level i is (B*H*W,1,H>>i,W>>i) ~N(0,1)) and flow ~N(0,sigma^2)."""
    W = H if W is None else W
    g = torch.Generator().manual_seed(seed)
    pyr = [torch.randn(B * H * W, 1, H >> i, W >> i, generator=g).to(vol_dtype) for i in range(L)]
    flow = flow_sigma * torch.randn(B, 2, H, W, generator=g)
    return pyr, flow


def random_affines(B, seed=0, size=224):
    """Plausible stage-2 outputs: rotation about the crop centre, scale 0.7-1.4, shift +-20 px."""
    g = torch.Generator().manual_seed(seed)
    ang = (torch.rand(B, generator=g) - 0.5) * math.pi
    sc = 0.7 + 0.7 * torch.rand(B, generator=g)
    tr = (torch.rand(B, 2, generator=g) - 0.5) * 40.0
    c = size / 2.0
    M = torch.zeros(B, 3, 3)
    M[:, 0, 0] = sc * torch.cos(ang)
    M[:, 0, 1] = -sc * torch.sin(ang)
    M[:, 1, 0] = sc * torch.sin(ang)
    M[:, 1, 1] = sc * torch.cos(ang)
    M[:, 0, 2] = c - (M[:, 0, 0] * c + M[:, 0, 1] * c) + tr[:, 0]
    M[:, 1, 2] = c - (M[:, 1, 0] * c + M[:, 1, 1] * c) + tr[:, 1]
    M[:, 2, 2] = 1.0
    return M
