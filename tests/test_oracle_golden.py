"""Pin the oracle: replay the reference's own outputs (tests/golden/*.npz, minted by
oracle/make_golden.py from /root/reference) through oracle/*.py.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import correspondence_oracle as OC
from oracle import corr_lookup_oracle as OL
from oracle import matching_oracle as OM

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


T = torch.from_numpy


MATCH = ["small", "medium", "bern", "ones", "allmasked", "identical", "tiles"]   # tiles: 32x32 patches, prefix masks


@pytest.mark.parametrize("name", MATCH)
def test_matching_templates_matches_reference(name):
    g = load(f"match_{name}.npz")
    src, tar = torch.from_numpy(g["src"]), torch.from_numpy(g["tar"])
    mask = torch.from_numpy(g["mask"]).float()
    k = int(g["topk"])
    sim_avg = OM.template_scores(src, tar, mask)
    np.testing.assert_allclose(sim_avg.numpy(), g["sim_avg"], rtol=0, atol=2e-6)
    score, idx = OM.matching_templates(src, tar, None, mask, topk=k)
    np.testing.assert_allclose(score.numpy(), g["score"], rtol=0, atol=2e-6)
    if name != "allmasked":                      # all-zero scores: topk tie order is unspecified
        np.testing.assert_array_equal(idx.numpy(), g["idx"])
    assert idx.dtype == torch.int64 and score.dtype == torch.float32


@pytest.mark.parametrize("name", ["small", "ones", "allmasked", "identical"])
def test_matching_loops_agree(name):
    g = load(f"match_{name}.npz")
    out = OM.template_scores_loops(g["src"], g["tar"], g["mask"].astype(np.float32))
    np.testing.assert_allclose(out, g["sim_avg"], rtol=0, atol=3e-6)


def test_matching_kats():
    g = load("match_allmasked.npz")
    assert np.all(g["sim_avg"] == 0.0)                     # all-masked query scores exactly 0
    g = load("match_identical.npz")
    # identical query/template 0, full mask: every patch but index 0 is its own best match
    T = g["src"].shape[-1] ** 2
    H = g["src"].shape[-1]
    np.testing.assert_allclose(g["sim_avg"][0, 0], (T - 1) / float(H * H), atol=1e-5)
    g = load("match_medium.npz")
    np.testing.assert_array_equal(g["idx"][0], g["planted"][0][:5])  # planted ranking recovered


@pytest.mark.parametrize("name", ["small", "medium"])
def test_similarity_volume(name):
    g = load(f"sim_{name}.npz")
    src, tar = torch.from_numpy(g["src"]), torch.from_numpy(g["tar"])
    sm = torch.from_numpy(g["src_mask"]).float()
    out = OM.similarity_volume(src, tar, sm)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=2e-6)
    if name == "small":
        np.testing.assert_allclose(OM.similarity_volume_loops(g["src"], g["tar"], g["src_mask"]), g["out"], atol=3e-6)


LOOKUP = ["small", "ramp", "ladder", "rect", "intflow"]


@pytest.mark.parametrize("name", LOOKUP)
def test_corr_lookup(name):
    g = load(f"lookup_{name}.npz")
    L, r = int(g["levels"]), int(g["radius"])
    pyr = [torch.from_numpy(g[f"pyr{i}"]) for i in range(L)]
    flow = torch.from_numpy(g["flow"])
    out = OL.corr_lookup(pyr, flow, r)
    assert out.shape == g["out"].shape
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=5e-6)


@pytest.mark.parametrize("name", ["small", "ramp", "intflow"])
def test_corr_lookup_loops(name):
    g = load(f"lookup_{name}.npz")
    L, r = int(g["levels"]), int(g["radius"])
    out = OL.corr_lookup_loops([g[f"pyr{i}"] for i in range(L)], g["flow"], r)
    np.testing.assert_allclose(out, g["out"], rtol=0, atol=5e-6)


def test_corr_lookup_ramp_kat():
    g = load("lookup_ramp.npz")
    # x-ramp volume, zero flow, r=1: window is x-major -> channels [3,3,3,4,4,4,5,5,5] at (4,4)
    np.testing.assert_allclose(g["out"][0, :, 4, 4], [3, 3, 3, 4, 4, 4, 5, 5, 5], atol=1e-5)


def test_bilinear_and_grid():
    g = load("bilinear.npz")
    feat, grid = torch.from_numpy(g["feat"]), torch.from_numpy(g["grid"])
    np.testing.assert_allclose(OL.bilinear_sample(feat, grid.clone(), True).numpy(), g["out_true"], atol=3e-6)
    np.testing.assert_allclose(OL.bilinear_sample(feat, grid.clone(), False).numpy(), g["out_false"], atol=3e-6)
    np.testing.assert_array_equal(OL.coords_grid(2, 7, 5).numpy(), g["coords"])


def test_sampling_modes_and_nonfinite_flows():
    """The argument combinations PicoPose does not use (nearest / bicubic, border / reflection, align_corners=False) and
    NaN / infinite flows, against reference outputs (oracle/make_golden.py r2b)."""
    g = load("sample_modes.npz")
    feat, grid = T(g["feat"]), T(g["grid"])
    for key in g.files:
        if key in ("feat", "grid"):
            continue
        mode, pad, ac = key.rsplit("_", 2)
        out = OL.grid_sample(feat, grid.clone(), mode, pad, bool(int(ac)))
        np.testing.assert_allclose(out.numpy(), g[key], rtol=0, atol=5e-6, err_msg=key)
    g = load("lookup_modes.npz")
    pyr, flow = [T(g["pyr0"]), T(g["pyr1"])], T(g["flow"])
    for key in g.files:
        if key in ("flow", "pyr0", "pyr1", "radius"):
            continue
        mode, pad, ac = key.rsplit("_", 2)
        out = OL.corr_lookup_general(pyr, flow, int(g["radius"]), mode, pad, bool(int(ac)))
        np.testing.assert_allclose(out.numpy(), g[key], rtol=0, atol=5e-6, err_msg=key)
    g = load("lookup_nonfinite.npz")
    pyr, flow = [T(g["pyr0"]), T(g["pyr1"])], T(g["flow"])
    out = OL.corr_lookup(pyr, flow, int(g["radius"])).numpy()
    assert np.isnan(g["out"]).sum() == 175                            # 3 poisoned queries x 2 levels + one at level 0 only
    np.testing.assert_allclose(out, g["out"], rtol=0, atol=2e-6, equal_nan=True)
    warped = OL.bilinear_sample(T(g["feat"]), OL.coords_grid(1, 8, 8) + flow, align_corners=True).numpy()
    np.testing.assert_allclose(warped, g["warped"], rtol=0, atol=2e-6, equal_nan=True)


def test_correlation_pyramid():
    g = load("pyramid.npz")
    pyr = OL.correlation_pyramid(torch.from_numpy(g["f1"]), torch.from_numpy(g["f2"]), 3)
    for i, p in enumerate(pyr):
        np.testing.assert_allclose(p.numpy(), g[f"lvl{i}"], atol=2e-6)


def test_init_correspondences():
    g = load("corresp_init.npz")
    flow, cert = OC.init_correspondences(torch.from_numpy(g["Ms"]), torch.from_numpy(g["mask"]).float())
    np.testing.assert_allclose(flow.numpy(), g["flow"], atol=2e-5)
    np.testing.assert_array_equal(cert.numpy(), g["cert"])
    np.testing.assert_allclose(g["flow"][0], 0.5, atol=1e-6)          # identity: half-patch offset
    np.testing.assert_allclose(g["flow"][1, 0], 2.5, atol=1e-6)       # +28 px in x = +2 patches
    np.testing.assert_allclose(g["flow"][1, 1], 0.5, atol=1e-6)


def test_stage3_correspondences():
    g = load("corresp_stage3.npz")
    tar, src = OC.stage3_correspondences(torch.from_numpy(g["flow"]), torch.from_numpy(g["cert"]))
    np.testing.assert_array_equal(tar.numpy(), g["tar"])
    np.testing.assert_array_equal(src.numpy(), g["src"])
    assert tar.dtype == torch.int64
    # KAT from SURVEY 8(c): k = w*H + h ; k=1 -> src (0,1), tar (0,0); k in {0,4,8,9,12..15} -> -1
    np.testing.assert_array_equal(g["src"][0, 1], [0, 1])
    np.testing.assert_array_equal(g["tar"][0, 1], [0, 0])
    for k in (0, 4, 8, 9, 12, 13, 14, 15):
        np.testing.assert_array_equal(g["src"][0, k], [-1, -1])
    tar, src = OC.stage3_correspondences(torch.from_numpy(g["flow_r"]), torch.from_numpy(g["cert_r"]))
    np.testing.assert_array_equal(tar.numpy(), g["tar_r"])
    np.testing.assert_array_equal(src.numpy(), g["src_r"])


# ------------------------------------------------------------------------------------------------
# round 2: hypothesis selection, lookup + first motion-encoder conv, the FlowDecoder loop
# ------------------------------------------------------------------------------------------------

def _hyp_inputs(g):
    ep = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in_")}
    return ep, torch.from_numpy(g["pred_id"])


def test_select_template_data_matches_reference():
    from oracle import hypotheses_oracle as OH
    g = load("hyp_select.npz")
    ep, pred_id = _hyp_inputs(g)
    for k in range(pred_id.shape[1]):
        sel = OH.select_template_data(ep, pred_id, k)
        assert set(sel) == set(OH.TEMPLATE_KEYS) | set(OH.REAL_KEYS)
        for key, v in sel.items():
            np.testing.assert_array_equal(v.numpy(), g[f"out{k}_{key}"])


def test_lookup_then_first_motion_conv_matches_reference():
    """corr_net[0] of the reference's MotionEncoder on the reference's lookup output: a 1x1 conv + ReLU, i.e. a
    (256 x 50) matrix applied per query -- the step libpicopose_b200 fuses into the correlation kernel (f-3)."""
    g = load("motion_conv.npz")
    f1, f2, flow = (torch.from_numpy(g[k]) for k in ("f1", "f2", "flow"))
    corr = OL.corr_lookup(OL.correlation_pyramid(f1, f2, 2), flow, 2)
    np.testing.assert_allclose(corr.numpy(), g["corr"], rtol=0, atol=3e-6)
    out = OL.conv1x1_relu(corr, torch.from_numpy(g["weight"]), torch.from_numpy(g["bias"]))
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=1e-5)


def test_flow_decoder_restatement_matches_reference():
    """The seeded restatement reproduces the reference FlowDecoder's weights (checksums of all 27 M parameters) and its
    outputs on the seeded inputs (tests/golden/flow_decoder.npz was produced by the reference's own class)."""
    import json
    from oracle import flow_decoder_oracle as OF
    g = load("flow_decoder.npz")
    seed = int(g["seed"])
    torch.manual_seed(seed)
    dec = OF.FlowDecoder(3, 4).eval()
    want = json.loads(str(g["checksums"]))
    got = OF.weight_checksums(dec)
    assert sorted(got) == sorted(want)
    for k, (s, a) in want.items():
        assert got[k][0] == pytest.approx(s, rel=1e-9, abs=1e-9) and got[k][1] == pytest.approx(a, rel=1e-9, abs=1e-9), k
    render, real, flow0, cert0 = OF.decoder_inputs(seed + 1)
    with torch.no_grad():
        flows, certs = dec(render, real, flow0, cert0)
    for i in range(3):
        np.testing.assert_allclose(flows[i].numpy(), g[f"flow{i}"], rtol=0, atol=2e-4)
        np.testing.assert_allclose(certs[i].numpy(), g[f"cert{i}"], rtol=0, atol=2e-4)
