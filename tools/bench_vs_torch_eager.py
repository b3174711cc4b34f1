"""Every hot-path function at its native / north-star size: this library vs the same computation written with
stock PyTorch ops on the same B200 (what the reference repository itself executes when it runs on a GPU:
F.normalize + einsum + max/topk for stage 1, matmul + AvgPool2d + F.grid_sample for stage 3).

    python tools/bench_vs_torch_eager.py [--json out.json]

The torch versions below follow utils/matching.py:29-69, :6-26 and utils/corr_lookup.py:29-65, :100-134 of the
reference line by line (they are the baseline being timed, and each result is compared with ours)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


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


def torch_matching_templates(src, tar, tar_mask, topk):
    B, N, C, H, W = src.shape
    m = F.interpolate(tar_mask[:, None], size=(H, W), mode="nearest").flatten(1)
    t = F.normalize(tar, dim=1).flatten(2).transpose(1, 2)                     # b t c
    s = F.normalize(src, dim=2).flatten(3).transpose(2, 3)                     # b n s c
    sim = torch.einsum("btc,bnsc->bnts", t, s) * m[:, None, :, None]
    sc, i_t2s = sim.max(dim=3)
    _, i_s2t = sim.max(dim=2)
    valid = m[:, None, :] * (i_s2t != 0) * (i_t2s != 0)
    avg = (sc * valid).sum(dim=2) / (H * H)
    return torch.topk(avg, topk, dim=1)


def torch_similarity(src, tar, src_mask):
    B, C, H, W = src.shape
    m = F.interpolate(src_mask[:, None], size=(H, W), mode="nearest").flatten(1)
    s = F.normalize(src, dim=1).flatten(2)
    t = F.normalize(tar, dim=1).flatten(2)
    sim = torch.einsum("bct,bcs->bts", t, s) * m[:, None, :]
    return sim.clamp(min=0).permute(0, 2, 1).reshape(B, H * W, W, H).permute(0, 1, 3, 2)   # "b (w h) c -> b c h w"


def torch_lookup(pyr, flow, r):
    B, _, H, W = flow.shape
    d = torch.arange(-r, r + 1, device=flow.device, dtype=torch.float32)
    dx, dy = torch.meshgrid(d, d, indexing="ij")
    delta = torch.stack([dx, dy], dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
    xs = torch.arange(W, device=flow.device, dtype=torch.float32)
    ys = torch.arange(H, device=flow.device, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(xs, ys, indexing="xy"), dim=0)[None] + flow
    cen = grid.permute(0, 2, 3, 1).reshape(B * H * W, 1, 1, 2)
    out = []
    for i, c in enumerate(pyr):
        g = cen / 2 ** i + delta
        h, w = c.shape[-2:]
        g = torch.stack([2 * g[..., 0] / max(w - 1, 1) - 1, 2 * g[..., 1] / max(h - 1, 1) - 1], dim=-1)
        out.append(F.grid_sample(c, g, mode="bilinear", padding_mode="zeros", align_corners=True).view(B, H, W, -1))
    return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    from picopose_b200 import matching as M
    from picopose_b200 import synth
    from picopose_b200.corr_lookup import CorrLookup, bilinear_sample
    from picopose_b200.correlation import CorrelationPyramid
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    rows = []

    def add(name, t_ours, t_torch, err):
        row = {"op": name, "ours_ms": t_ours, "torch_eager_ms": t_torch, "speedup": t_torch / t_ours, "max_abs_diff": err}
        rows.append(row)
        print(json.dumps(row), flush=True)

    # stage 1, native run_test.py size and config 2
    for label, B, N, C, H in (("matching_templates native (4 x 162 x 1024 x 16^2)", 4, 162, 1024, 16),
                              ("matching_templates config 2 (1 x 162 x 1024 x 32^2)", 1, 162, 1024, 32)):
        src = torch.randn(B, N, C, H, H, device=dev, generator=g)
        tar = src[:, 3].clone() + 0.5 * torch.randn(B, C, H, H, device=dev, generator=g)
        mask = synth.disc_mask(B).to(dev)
        ref_s, ref_i = torch_matching_templates(src, tar, mask, 5)
        s, i = M.matching_templates(src, tar, None, mask, topk=5)
        assert i[:, 0].tolist() == ref_i[:, 0].tolist()
        err = float((s - ref_s).abs().max())
        add(label + ", fp32 features in (cold)", timed(lambda: M.matching_templates(src, tar, None, mask, topk=5), a.iters),
            timed(lambda: torch_matching_templates(src, tar, mask, 5), a.iters), err)
        bank = M.TemplateBank.from_features(src)
        add(label + ", resident TemplateBank", timed(lambda: M.matching_templates(bank, tar, None, mask, topk=5), a.iters),
            rows[-1]["torch_eager_ms"], err)
        del src, bank
    # stage 2 input: similarity volume for B*k hypotheses
    B, C, H = 20, 1024, 16
    src = torch.randn(B, C, H, H, device=dev, generator=g)
    tar = torch.randn(B, C, H, H, device=dev, generator=g)
    sm = synth.bernoulli_mask(B, 224, 0.7, 8).to(dev)
    ref = torch_similarity(src, tar, sm)
    out = M.matching_features_similarity(src, tar, sm, None)
    add("matching_features_similarity (20 x 1024 x 16^2)", timed(lambda: M.matching_features_similarity(src, tar, sm, None), a.iters),
        timed(lambda: torch_similarity(src, tar, sm), a.iters), float((out - ref).abs().max()))
    # stage 3: one FlowDecoder level, pyramid + lookup, and the feature warp
    for H, L in ((16, 1), (32, 2), (64, 3)):
        N, C, r = 20, 256, 2
        f1 = torch.randn(N, C, H, H, device=dev, generator=g)
        f2 = torch.randn(N, C, H, H, device=dev, generator=g)
        flow = 2.0 * torch.randn(N, 2, H, H, device=dev, generator=g)
        pool = torch.nn.AvgPool2d(2, 2)

        def torch_level():
            corr = torch.matmul(f1.view(N, C, -1).permute(0, 2, 1), f2.view(N, C, -1)).view(N * H * H, 1, H, H) / 16.0
            pyr = [corr]
            for _ in range(L - 1):
                pyr.append(pool(pyr[-1]))
            return torch_lookup(pyr, flow, r)

        pyr_mod, look = CorrelationPyramid(num_levels=L), CorrLookup(radius=r)
        ref = torch_level()
        out = look(pyr_mod(f1, f2), flow)
        add("CorrelationPyramid + CorrLookup (20 x 256 x %d^2, L=%d, r=2)" % (H, L), timed(lambda: look(pyr_mod(f1, f2), flow), a.iters),
            timed(torch_level, a.iters), float((out - ref).abs().max()))
        xs = torch.arange(H, device=dev, dtype=torch.float32)
        grid = (torch.stack(torch.meshgrid(xs, xs, indexing="xy"), dim=0)[None] + flow).permute(0, 2, 3, 1).contiguous()

        def torch_warp():
            gg = torch.stack([2 * grid[..., 0] / (H - 1) - 1, 2 * grid[..., 1] / (H - 1) - 1], dim=-1)
            return F.grid_sample(f2, gg, mode="bilinear", padding_mode="zeros", align_corners=True)

        ref = torch_warp()
        out = bilinear_sample(f2, grid, align_corners=True)
        add("feature_sample / bilinear_sample (20 x 256 x %d^2)" % H, timed(lambda: bilinear_sample(f2, grid, align_corners=True), a.iters),
            timed(torch_warp, a.iters), float((out - ref).abs().max()))
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
