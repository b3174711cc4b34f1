"""Stage-3 integration on the GPU: the whole FlowDecoder refinement loop (three pyramid levels) with OUR CUDA modules
in the places the overlay puts them -- CorrelationPyramid, CorrLookup, bilinear_sample (feature_sample), coords_grid,
compute_stage3_correspondences -- against outputs of the REFERENCE's own FlowDecoder (tests/golden/flow_decoder.npz,
minted by oracle/make_golden.py from /root/reference with seeded weights).  The conv stacks are the reference's
architecture restated in oracle/flow_decoder_oracle.py (the reference tree does not exist on the GPU box) and run in
cuDNN fp32 (TF32 off); its seeded weights are verified against the fixture's checksums before use."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import correspondence_oracle as OC
from oracle import flow_decoder_oracle as OF
from picopose_b200 import _lib

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


class CudaOps:
    """The four call sites of model/stage3/flow_decoder.py:49-66 on libpicopose_b200."""

    def pyramid(self, f1, f2, num_levels):
        from picopose_b200.correlation import CorrelationPyramid
        return CorrelationPyramid(num_levels=num_levels)(f1, f2)

    def lookup(self, pyramid, flow, radius):
        from picopose_b200.corr_lookup import CorrLookup
        return CorrLookup(radius=radius)(pyramid, flow)

    def warp(self, feat, grid):
        from picopose_b200.corr_lookup import bilinear_sample
        return bilinear_sample(feat, grid, 'bilinear', 'zeros', True)

    def coords(self, B, H, W, device):
        from picopose_b200.corr_lookup import coords_grid
        return coords_grid(B, torch.arange(0, W, device=device), torch.arange(0, H, device=device))

    def encode(self, encoder, corr, flow):
        # what the overlay's MotionEncoder.forward does: first 1x1 conv of corr_net inside the lookup kernel when it can
        from picopose_b200.correlation import motion_encoder_forward
        return motion_encoder_forward(encoder, corr, flow)


@pytest.mark.parametrize("fused", ["conv", "1", "0"])
def test_flow_decoder_loop_against_the_reference(fused, monkeypatch):
    import picopose_b200.correlation as _corr
    # conv: windowed correlation + MotionEncoder's first 1x1 conv in one kernel (what the overlay runs);
    # 1: windowed correlation (no volume); 0: materialised (tiled) pyramid + lookup
    monkeypatch.setenv("PICOPOSE_B200_FUSED_CORR", "0" if fused == "0" else "1")
    monkeypatch.setattr(_corr, "ENCODER_FUSION", fused == "conv")
    g = np.load(os.path.join(GOLDEN, "flow_decoder.npz"))
    seed = int(g["seed"])
    torch.manual_seed(seed)
    dec = OF.FlowDecoder(3, 4, ops=CudaOps()).eval()
    want = json.loads(str(g["checksums"]))
    got = OF.weight_checksums(dec)
    for k, (s, a) in want.items():
        assert got[k][0] == pytest.approx(s, rel=1e-9, abs=1e-9) and got[k][1] == pytest.approx(a, rel=1e-9, abs=1e-9), k
    dec = dec.to(DEV)
    render, real, flow0, cert0 = OF.decoder_inputs(seed + 1)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            flows, certs = dec([t.to(DEV) for t in render], [t.to(DEV) for t in real], flow0.to(DEV), cert0.to(DEV))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    _lib.check_device_faults()
    worst = 0.0
    for i in range(3):
        worst = max(worst, float(np.abs(flows[i].cpu().numpy() - g[f"flow{i}"]).max()),
                    float(np.abs(certs[i].cpu().numpy() - g[f"cert{i}"]).max()))
        # refined correspondence coordinates: far inside the 0.05 px the north star allows
        np.testing.assert_allclose(flows[i].cpu().numpy(), g[f"flow{i}"], rtol=0, atol=2e-3)
        np.testing.assert_allclose(certs[i].cpu().numpy(), g[f"cert{i}"], rtol=0, atol=2e-3)
    print("FlowDecoder on the CUDA ops (fused=%s): max |diff| vs the reference = %.2e" % (fused, worst))
    # ... and the integer correspondences the reference derives from its final flow (utils/correspondence.py:28-59)
    from picopose_b200.correspondence import compute_stage3_correspondences
    tar, src = compute_stage3_correspondences(flows[-1], certs[-1])
    ref_flow, ref_cert = torch.from_numpy(g["flow2"]), torch.from_numpy(g["cert2"])
    tar_ref, src_ref = OC.stage3_correspondences(ref_flow, ref_cert)
    grid = OF.OL.coords_grid(1, 64, 64) + ref_flow
    frac = (grid - torch.round(grid)).abs()
    lo = torch.minimum(grid[:, 0], grid[:, 1])
    hi = torch.maximum(grid[:, 0], grid[:, 1])
    safe = (frac.min(dim=1).values > 1e-2) & ((torch.sigmoid(ref_cert[:, 0]) - 0.5).abs() > 1e-2) & \
           ((lo - 0).abs() > 1e-2) & ((hi - 63).abs() > 1e-2)                 # (B,H,W): decisions not within 0.01 of a boundary
    safe_k = safe.permute(0, 2, 1).reshape(1, -1)                             # flat index k = w*H + h
    assert float(safe_k.float().mean()) > 0.9
    assert torch.equal(tar.cpu()[safe_k], tar_ref[safe_k]) and torch.equal(src.cpu()[safe_k], src_ref[safe_k])
