"""Mint golden vectors by RUNNING THE REFERENCE (build container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz

Imports the reference's own modules from /root/reference (read-only; nothing is
copied) and records inputs + outputs of

    utils.matching.matching_templates / matching_features_similarity
    utils.corr_lookup.CorrLookup / bilinear_sample / coords_grid
    utils.correspondence.compute_init_correspondences / compute_stage3_correspondences
    model.stage3.raft_decoder.CorrelationPyramid   (through a stub mmcv.cnn.ConvModule)
    model.picopose.Net.select_template_data, model.stage3.raft_decoder.MotionEncoder.corr_net[0],
    model.stage3.flow_decoder.FlowDecoder.forward   (`python oracle/make_golden.py r2` mints only these)
    utils.corr_lookup.bilinear_sample / CorrLookup with the other interpolation / padding modes and with
    non-finite flows                                (`python oracle/make_golden.py r2b` mints only these)

on seeded synthetic inputs (picopose_b200/synth.py).  /root/reference does not
exist on the GPU box, so tests only ever read the committed .npz files.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PICOPOSE_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference tree not found at {REF}")
    sys.path.insert(0, REF)
    # CorrelationPyramid's module imports mmcv.cnn.ConvModule (not installed);
    # it carries no arithmetic of this path, a stub is enough to import the file.
    mmcv = types.ModuleType("mmcv")
    cnn = types.ModuleType("mmcv.cnn")

    class ConvModule(torch.nn.Module):
        """Conv2d -> ReLU (act_cfg None: no activation); child named `conv` like mmcv's."""

        def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, act_cfg=dict(type="ReLU"), **kw):
            super().__init__()
            self.conv = torch.nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
            self.act = torch.nn.ReLU() if act_cfg is not None else torch.nn.Identity()

        def forward(self, x):
            return self.act(self.conv(x))

    cnn.ConvModule = ConvModule
    mmcv.cnn = cnn
    sys.modules.setdefault("mmcv", mmcv)
    sys.modules.setdefault("mmcv.cnn", cnn)
    warnings.filterwarnings("ignore")
    import utils.matching as ref_matching
    import utils.corr_lookup as ref_lookup
    import utils.correspondence as ref_corresp
    from model.stage3.raft_decoder import CorrelationPyramid
    return ref_matching, ref_lookup, ref_corresp, CorrelationPyramid


def _np(t):
    return t.detach().cpu().numpy()


def main_round2():
    """Fixtures added in round 2: hypothesis selection (model/picopose.py:52-70), lookup + first MotionEncoder conv
    (model/stage3/raft_decoder.py:113-116,157) and the whole FlowDecoder loop (model/stage3/flow_decoder.py:74-94)."""
    _import_reference()
    import json
    from model.picopose import Net
    from model.stage3.flow_decoder import FlowDecoder
    from model.stage3.raft_decoder import CorrelationPyramid, MotionEncoder
    from utils.corr_lookup import CorrLookup
    from oracle import flow_decoder_oracle as OF
    os.makedirs(OUT, exist_ok=True)

    # ---- Net.select_template_data: the method does not touch `self` ----
    g = torch.Generator().manual_seed(60)
    B, N, hm, hp = 3, 6, 8, 4
    ep = {"tem_pose": torch.randn(B, N, 4, 4, generator=g), "tem_K": torch.randn(B, N, 3, 3, generator=g),
          "tem_M": torch.randn(B, N, 3, 3, generator=g), "tem_mask": (torch.rand(B, N, hm, hm, generator=g) > 0.5).float(),
          "tem_rgb": torch.rand(B, N, 3, hm, hm, generator=g), "tem_pts3d": torch.randn(B, N, hp, hp, 3, generator=g),
          "real_pts2d": torch.randn(B, hp, hp, 2, generator=g), "real_K": torch.randn(B, 3, 3, generator=g),
          "real_M": torch.randn(B, 3, 3, generator=g), "real_mask": (torch.rand(B, hm, hm, generator=g) > 0.5).float(),
          "real_pose": torch.randn(B, 4, 4, generator=g)}
    pred_id = torch.stack([torch.randperm(N, generator=g)[:3] for _ in range(B)])
    d = {"in_" + k: _np(v) for k, v in ep.items()}
    d["pred_id"] = _np(pred_id)
    for k in range(3):
        sel = Net.select_template_data(None, ep, pred_id, k)
        for key, v in sel.items():
            d[f"out{k}_{key}"] = _np(v)
    np.savez_compressed(os.path.join(OUT, "hyp_select.npz"), **d)
    print("hyp_select: keys", sorted(sel))

    # ---- lookup followed by the motion encoder's first (1x1) convolution ----
    torch.manual_seed(61)
    enc = MotionEncoder(num_levels=2, radius=2, net_type="Basic", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU")).eval()
    g = torch.Generator().manual_seed(62)
    f1 = torch.randn(2, 32, 16, 16, generator=g)
    f2 = torch.randn(2, 32, 16, 16, generator=g)
    flow = 2.0 * torch.randn(2, 2, 16, 16, generator=g)
    with torch.no_grad():
        corr = CorrLookup(radius=2)(CorrelationPyramid(num_levels=2)(f1, f2), flow.clone())
        feat0 = enc.corr_net[0](corr)
    np.savez_compressed(os.path.join(OUT, "motion_conv.npz"), f1=_np(f1), f2=_np(f2), flow=_np(flow), corr=_np(corr),
                        weight=_np(enc.corr_net[0].conv.weight), bias=_np(enc.corr_net[0].conv.bias), out=_np(feat0))
    print("motion_conv: corr", tuple(corr.shape), "->", tuple(feat0.shape))

    # ---- the whole FlowDecoder (3 levels, radius 4 -> lookup radius 2), seeded weights, eval mode ----
    seed = 1234
    torch.manual_seed(seed)
    ref = FlowDecoder(num_levels=3, radius=4).eval()
    torch.manual_seed(seed)
    mine = OF.FlowDecoder(3, 4).eval()
    sd_ref, sd_mine = ref.state_dict(), mine.state_dict()
    assert list(sd_ref) == list(sd_mine), "parameter names / order differ"
    assert all(torch.equal(sd_ref[k], sd_mine[k]) for k in sd_ref), "seeded construction does not reproduce the reference's weights"
    render, real, flow0, cert0 = OF.decoder_inputs(seed + 1)
    with torch.no_grad():
        fr, cr = ref([t.clone() for t in render], [t.clone() for t in real], flow0.clone(), cert0.clone())
        fm, cm = mine(render, real, flow0, cert0)
    err = max(float((a - b).abs().max()) for a, b in zip(fr + cr, fm + cm))
    assert err < 1e-4, err
    print("flow_decoder: restatement vs reference max |diff| = %.2e" % err)
    sums = OF.weight_checksums(ref)
    np.savez_compressed(os.path.join(OUT, "flow_decoder.npz"), seed=seed,
                        checksums=json.dumps(sums), restatement_vs_reference=err,
                        **{f"flow{i}": _np(t) for i, t in enumerate(fr)}, **{f"cert{i}": _np(t) for i, t in enumerate(cr)})
    print("round-2 golden vectors written to", OUT)


def main_sampling_modes():
    """`python oracle/make_golden.py r2b`: the argument combinations of bilinear_sample / CorrLookup that PicoPose does
    not use (other interpolation and padding modes, align_corners=False) and non-finite flows, from the reference."""
    from picopose_b200 import synth
    _, ref_lookup, _, _ = _import_reference()
    g = torch.Generator().manual_seed(41)
    feat = torch.randn(2, 3, 6, 9, generator=g)
    # coordinates from well outside the map (several reflections) to inside, plus exact pixel centres and borders
    grid = torch.stack([torch.rand(2, 5, 8, generator=g) * 40 - 15, torch.rand(2, 5, 8, generator=g) * 30 - 12], dim=-1)
    grid[0, 0, :4, 0] = torch.tensor([0.0, 8.0, 4.0, -0.5])
    grid[0, 0, :4, 1] = torch.tensor([0.0, 5.0, 2.5, 5.5])
    d = dict(feat=_np(feat), grid=_np(grid))
    for mode in ("bilinear", "nearest", "bicubic"):
        for pad in ("zeros", "border", "reflection"):
            for ac in (True, False):
                d[f"{mode}_{pad}_{int(ac)}"] = _np(ref_lookup.bilinear_sample(feat, grid.clone(), mode, pad, ac))
    np.savez_compressed(os.path.join(OUT, "sample_modes.npz"), **d)
    print("sample_modes:", len(d) - 2, "combinations")

    pyr, flow = synth.lookup_inputs(1, 8, 2, seed=42, flow_sigma=3.0)
    d = dict(flow=_np(flow), **{f"pyr{i}": _np(v) for i, v in enumerate(pyr)})
    for mode, pad, ac in (("bilinear", "zeros", False), ("bilinear", "border", True), ("nearest", "zeros", True),
                          ("bicubic", "reflection", False), ("nearest", "reflection", False)):
        out = ref_lookup.CorrLookup(2, mode, pad, ac)([v.clone() for v in pyr], flow.clone())
        d[f"{mode}_{pad}_{int(ac)}"] = _np(out)
    np.savez_compressed(os.path.join(OUT, "lookup_modes.npz"), radius=2, **d)

    # non-finite flows: a NaN / infinite coordinate poisons every tap weight of the query (F.grid_sample), a finite
    # far-away one is padding; 3e38 overflows the reference's `* 2.` at level 0 only
    pyr, flow = synth.lookup_inputs(1, 8, 2, seed=43, flow_sigma=1.0)
    flow[0, 0, 1, 1] = float("nan")
    flow[0, 1, 2, 3] = float("inf")
    flow[0, 0, 4, 4] = float("-inf")
    flow[0, 1, 4, 4] = float("nan")
    flow[0, 0, 5, 0] = 3.0e38
    flow[0, 1, 6, 6] = -1.0e30
    out = ref_lookup.CorrLookup(2)([v.clone() for v in pyr], flow.clone())
    feat = torch.randn(1, 4, 8, 8, generator=g)
    grid = (ref_lookup.coords_grid(1, torch.arange(8), torch.arange(8)) + flow)
    warped = ref_lookup.bilinear_sample(feat, grid.clone(), align_corners=True)
    np.savez_compressed(os.path.join(OUT, "lookup_nonfinite.npz"), flow=_np(flow), out=_np(out), radius=2, feat=_np(feat),
                        warped=_np(warped), **{f"pyr{i}": _np(v) for i, v in enumerate(pyr)})
    print("lookup_nonfinite: NaN outputs", int(torch.isnan(out).sum()), "of", out.numel())


def main():
    from picopose_b200 import synth

    if len(sys.argv) > 1 and sys.argv[1] == "r2":
        return main_round2()
    if len(sys.argv) > 1 and sys.argv[1] == "r2b":
        return main_sampling_modes()
    ref_matching, ref_lookup, ref_corresp, CorrelationPyramid = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    # ---------------- stage-1 matching ----------------
    def match_case(name, B, N, C, H, k, seed, mask="disc", special=None):
        src, tar, planted = synth.planted_match_inputs(B, N, C, H, seed=seed)
        if mask == "disc":
            m = synth.disc_mask(B)
        elif mask == "bern":
            m = synth.bernoulli_mask(B, 224, 0.7, seed + 100)
        elif mask == "ones":
            m = torch.ones(B, 224, 224)
        elif mask == "zeros":
            m = torch.zeros(B, 224, 224)
        elif mask == "prefix":
            # detection 0: the first 257 patches unmasked, detection 1: the last 700 (patch 0 masked) -- unmasked counts
            # on both sides of the 256-column tile boundaries of the CUDA contraction (match_gemm.cu)
            m = torch.zeros(B, 224, 224)
            step = 224 // H
            for b, n_on in enumerate((257, 700)):
                on = torch.arange(H * H) < n_on
                if b == 1:
                    on = on.flip(0)
                m[b, ::step, ::step][:H, :H] = on.view(H, H).float()
        if special == "identical":          # query == template 0 exactly
            tar = src[:, 0].clone()
        src_masks = torch.ones(B, N, 224, 224)[:, :, :1, :1]  # unused by the reference
        score, idx = ref_matching.matching_templates(src.clone(), tar.clone(), src_masks, m.clone(), topk=k)
        # also the dense per-template scores (k = N) for a sharper check
        score_all, idx_all = ref_matching.matching_templates(src.clone(), tar.clone(), src_masks, m.clone(), topk=N)
        sim_avg = torch.zeros(B, N)
        sim_avg.scatter_(1, idx_all, score_all)
        np.savez_compressed(os.path.join(OUT, f"match_{name}.npz"),
                            src=_np(src), tar=_np(tar), mask=_np(m).astype(np.uint8), topk=k,
                            score=_np(score), idx=_np(idx), sim_avg=_np(sim_avg), planted=_np(planted))
        print(f"match_{name}: top-{k} idx {idx.tolist()}")

    match_case("small", 2, 7, 16, 4, 3, seed=0)
    match_case("medium", 1, 12, 64, 8, 5, seed=1)
    match_case("bern", 2, 9, 32, 8, 5, seed=2, mask="bern")
    match_case("ones", 1, 6, 16, 4, 5, seed=3, mask="ones")
    match_case("allmasked", 1, 6, 16, 4, 2, seed=4, mask="zeros")
    match_case("identical", 1, 6, 16, 4, 2, seed=5, mask="ones", special="identical")
    match_case("tiles", 2, 2, 16, 32, 2, seed=6, mask="prefix")

    # ---------------- stage-2 similarity volume ----------------
    for name, B, C, H, seed in (("small", 2, 16, 4, 10), ("medium", 1, 32, 8, 11)):
        g = torch.Generator().manual_seed(seed)
        src = torch.randn(B, C, H, H, generator=g)
        tar = torch.randn(B, C, H, H, generator=g)
        sm = synth.bernoulli_mask(B, 224, 0.7, seed + 1)
        tm = torch.ones(B, 224, 224)
        out = ref_matching.matching_features_similarity(src.clone(), tar.clone(), sm.clone(), tm)
        np.savez_compressed(os.path.join(OUT, f"sim_{name}.npz"), src=_np(src), tar=_np(tar),
                            src_mask=_np(sm).astype(np.uint8), out=_np(out))
        print(f"sim_{name}: out {tuple(out.shape)}")

    # ---------------- stage-3 lookup ----------------
    def lookup_case(name, B, H, L, r, seed, sigma, W=None, ramp=False, zero_flow=False):
        pyr, flow = synth.lookup_inputs(B, H, L, seed=seed, flow_sigma=sigma, W=W)
        if ramp:                     # KAT: volume value == x coordinate
            Wv = pyr[0].shape[-1]
            pyr = [torch.arange(Wv, dtype=torch.float32).view(1, 1, 1, Wv).expand_as(pyr[0]).contiguous()]
        if zero_flow:
            flow = torch.zeros_like(flow)
        mod = ref_lookup.CorrLookup(radius=r)
        out = mod([p.clone() for p in pyr], flow.clone())
        d = {f"pyr{i}": _np(p) for i, p in enumerate(pyr)}
        np.savez_compressed(os.path.join(OUT, f"lookup_{name}.npz"), flow=_np(flow), out=_np(out),
                            radius=r, levels=L, **d)
        print(f"lookup_{name}: out {tuple(out.shape)}")

    lookup_case("small", 2, 6, 2, 2, seed=20, sigma=2.0)
    lookup_case("ramp", 1, 8, 1, 1, seed=21, sigma=0.0, ramp=True, zero_flow=True)
    lookup_case("ladder", 1, 16, 3, 4, seed=22, sigma=4.0)
    lookup_case("rect", 1, 8, 2, 3, seed=23, sigma=3.0, W=12)
    lookup_case("intflow", 1, 8, 1, 2, seed=24, sigma=0.0, zero_flow=True)

    # generic bilinear_sample (FlowDecoder.feature_sample uses align_corners=True)
    g = torch.Generator().manual_seed(30)
    feat = torch.randn(2, 3, 5, 7, generator=g)
    grid = torch.stack([torch.rand(2, 4, 6, generator=g) * 9 - 1, torch.rand(2, 4, 6, generator=g) * 7 - 1], dim=-1)
    out_t = ref_lookup.bilinear_sample(feat, grid.clone(), align_corners=True)
    out_f = ref_lookup.bilinear_sample(feat, grid.clone(), align_corners=False)
    cg = ref_lookup.coords_grid(2, torch.arange(0, 7), torch.arange(0, 5))
    np.savez_compressed(os.path.join(OUT, "bilinear.npz"), feat=_np(feat), grid=_np(grid),
                        out_true=_np(out_t), out_false=_np(out_f), coords=_np(cg))

    # correlation pyramid
    g = torch.Generator().manual_seed(40)
    f1 = torch.randn(2, 8, 8, 8, generator=g)
    f2 = torch.randn(2, 8, 8, 8, generator=g)
    pyr = CorrelationPyramid(num_levels=3)(f1, f2)
    np.savez_compressed(os.path.join(OUT, "pyramid.npz"), f1=_np(f1), f2=_np(f2),
                        **{f"lvl{i}": _np(p) for i, p in enumerate(pyr)})

    # ---------------- correspondence glue ----------------
    B = 4
    Ms = synth.random_affines(B, seed=50)
    Ms[0] = torch.eye(3)
    Ms[1] = torch.eye(3)
    Ms[1, 0, 2] = 28.0
    tm = synth.bernoulli_mask(B, 224, 0.8, 51)
    tm[0] = 1.0
    tm[1] = 1.0
    flow0, cert0 = ref_corresp.compute_init_correspondences(Ms.clone(), tm.clone())
    np.savez_compressed(os.path.join(OUT, "corresp_init.npz"), Ms=_np(Ms), mask=_np(tm).astype(np.uint8),
                        flow=_np(flow0), cert=_np(cert0))

    flow = torch.zeros(1, 2, 4, 4)
    flow[:, 0] = 0.6
    flow[:, 1] = -0.4
    cert = torch.full((1, 1, 4, 4), 3.0)
    cert[0, 0, 1, 2] = -3.0      # (h=1, w=2) -> flat k = w*H + h = 9
    tar, src = ref_corresp.compute_stage3_correspondences(flow.clone(), cert.clone())
    g = torch.Generator().manual_seed(52)
    flow_r = 3.0 * torch.randn(3, 2, 16, 16, generator=g)
    cert_r = 2.0 * torch.randn(3, 1, 16, 16, generator=g)
    tar_r, src_r = ref_corresp.compute_stage3_correspondences(flow_r.clone(), cert_r.clone())
    np.savez_compressed(os.path.join(OUT, "corresp_stage3.npz"), flow=_np(flow), cert=_np(cert),
                        tar=_np(tar), src=_np(src), flow_r=_np(flow_r), cert_r=_np(cert_r),
                        tar_r=_np(tar_r), src_r=_np(src_r))
    print("golden vectors written to", OUT)
    main_round2()


if __name__ == "__main__":
    main()
