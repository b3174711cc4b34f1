"""GPU parity of the hypothesis hand-over: the single-launch gather against the reference's
Net.select_template_data outputs (tests/golden/hyp_select.npz) and the oracle, and the one-pass (K*B) hypothesis loop
against the reference's sequential loop, through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import hypotheses_oracle as OH
from picopose_b200 import _lib

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def _golden():
    g = np.load(os.path.join(GOLDEN, "hyp_select.npz"))
    ep = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in_")}
    return g, ep, torch.from_numpy(g["pred_id"])


def test_select_template_data_golden():
    from picopose_b200.hypotheses import select_all_hypotheses, select_template_data
    g, ep, pred_id = _golden()
    ep_d = {k: v.to(DEV) for k, v in ep.items()}
    B, K = pred_id.shape
    for k in range(K):
        sel = select_template_data(ep_d, pred_id.to(DEV), k)
        assert set(sel) == set(OH.TEMPLATE_KEYS) | set(OH.REAL_KEYS)
        for key, v in sel.items():
            assert v.dtype == ep[key].dtype
            np.testing.assert_array_equal(v.cpu().numpy(), g[f"out{k}_{key}"])
    allh = select_all_hypotheses(ep_d, pred_id.to(DEV))
    for k in range(K):
        for key in OH.TEMPLATE_KEYS + OH.REAL_KEYS:
            np.testing.assert_array_equal(allh[key][k * B:(k + 1) * B].cpu().numpy(), g[f"out{k}_{key}"])
    _lib.check_device_faults()
    with pytest.raises(IndexError):
        select_template_data(ep_d, pred_id.to(DEV), K)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        select_template_data(ep, pred_id, 0)


def test_select_template_data_native_shapes():
    """run_test.py's shapes: 4 detections x 162 views, 224x224 masks / rgb crops, 16x16x3 point maps, odd byte counts and
    a non-16-byte-aligned tensor (the scalar copy path), int64 and uint8 payloads."""
    from picopose_b200.hypotheses import _gather_views
    g = torch.Generator().manual_seed(4)
    B, N, K = 4, 162, 5
    tensors = [torch.randn(B, N, 4, 4, generator=g), torch.rand(B, N, 3, 224, 224, generator=g),
               (torch.rand(B, N, 224, 224, generator=g) > 0.5).float(), torch.randn(B, N, 16, 16, 3, generator=g),
               torch.randint(0, 255, (B, N, 7), generator=g, dtype=torch.uint8),
               torch.randint(-5, 5, (B, N, 3), generator=g, dtype=torch.int64)]
    pred = torch.stack([torch.randperm(N, generator=g)[:K] for _ in range(B)])
    bidx = torch.arange(B)
    outs = _gather_views([t.to(DEV) for t in tensors], pred.to(DEV), -1)
    for t, o in zip(tensors, outs):
        ref = torch.cat([t[bidx, pred[:, k]] for k in range(K)])
        assert torch.equal(o.cpu(), ref)
    outs = _gather_views([t.to(DEV) for t in tensors], pred.to(DEV), 3)
    for t, o in zip(tensors, outs):
        assert torch.equal(o.cpu(), t[bidx, pred[:, 3]])


class _TinyNet:
    """Stand-in with the interface forward_test_batched drives on the reference Net: a 'backbone' and a stage-2/3
    function of (selected template data, real features) that is batch-agnostic, like the reference's."""

    def __init__(self, C, H):
        g = torch.Generator().manual_seed(11)
        self.proj = torch.randn(C, 3, generator=g).to(DEV)
        self.H = H
        self.calls = 0

    def feature_extractor(self, rgb):
        f = torch.einsum("cj,bjhw->bchw", self.proj, torch.nn.functional.adaptive_avg_pool2d(rgb, self.H))
        return [f * 0.5, f]

    def forward_test_hyp(self, ep, features_real):
        self.calls += 1
        s = ep["tem_rgb"].mean(dim=(1, 2, 3)) + features_real[-1].mean(dim=(1, 2, 3)) + ep["real_K"].sum(dim=(1, 2))
        return {"tem_pose": ep["tem_pose"], "pred_poses": ep["tem_pose"] * s.view(-1, 1, 1),
                "pred_tar_pts": (ep["tem_mask"].sum(dim=(1, 2)) + s).view(-1, 1)}


def test_forward_test_batched_equals_the_sequential_loop():
    from picopose_b200.hypotheses import forward_test_batched, patch_net, select_template_data
    from picopose_b200.matching import matching_templates
    from picopose_b200 import synth
    import torch.nn.functional as F
    B, N, C, H, K = 3, 9, 64, 8, 4
    g = torch.Generator().manual_seed(12)
    net = _TinyNet(C, H)
    ep = {"real_rgb": torch.rand(B, 3, 56, 56, generator=g), "tem_mask": (torch.rand(B, N, 16, 16, generator=g) > 0.4).float(),
          "real_mask": synth.disc_mask(B), "tem_pose": torch.randn(B, N, 4, 4, generator=g),
          "tem_K": torch.randn(B, N, 3, 3, generator=g), "tem_M": torch.randn(B, N, 3, 3, generator=g),
          "tem_rgb": torch.rand(B, N, 3, 16, 16, generator=g), "tem_pts3d": torch.randn(B, N, 4, 4, 3, generator=g),
          "real_pts2d": torch.randn(B, 4, 4, 2, generator=g), "real_K": torch.randn(B, 3, 3, generator=g),
          "real_M": torch.randn(B, 3, 3, generator=g), "real_pose": torch.randn(B, 4, 4, generator=g)}
    ep = {k: v.to(DEV) for k, v in ep.items()}
    with torch.no_grad():
        real = net.feature_extractor(ep["real_rgb"])[-1]
    ep["template_feature"] = torch.randn(B, N, C, H, H, generator=g).to(DEV)
    ep["template_feature"][torch.arange(B), torch.tensor([5, 0, 7])] += 2.0 * real       # a clear best view per detection
    # the reference's loop (model/picopose.py:97-112) with our drop-ins
    feats = net.feature_extractor(ep["real_rgb"])
    _, pred = matching_templates(F.normalize(ep["template_feature"], dim=2), feats[-1], ep["tem_mask"], ep["real_mask"], topk=K)
    assert pred[:, 0].tolist() == [5, 0, 7]
    loop = OH.hypothesis_loop(select_template_data, net.forward_test_hyp, ep, pred, feats)
    calls = net.calls
    batched = forward_test_batched(net, ep, hyp=K)
    assert net.calls == calls + 1 and len(batched) == K                      # ONE stage-2/3 pass
    for a, b in zip(loop, batched):
        assert set(a) == set(b)
        for key in a:
            assert a[key].shape == b[key].shape
            torch.testing.assert_close(a[key], b[key], rtol=1e-6, atol=1e-6)
    patched = patch_net(net)
    out = patched.forward_test(ep, hyp=K)
    torch.testing.assert_close(out[2]["pred_poses"], loop[2]["pred_poses"], rtol=1e-6, atol=1e-6)
    _lib.check_device_faults()
