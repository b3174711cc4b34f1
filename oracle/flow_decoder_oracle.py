"""Oracle for the stage-3 refinement loop around the lookup (TEST INFRASTRUCTURE, see oracle/__init__.py).

A restatement of the reference's FlowDecoder (model/stage3/flow_decoder.py:8-94) with its MotionEncoder ('Basic',
model/stage3/raft_decoder.py:56-161) and XHead (:251-289) sub-networks: per pyramid level a 1x1 projection + BatchNorm
of both feature maps, all-pairs correlation pyramid, windowed lookup around the current flow, motion encoder, feature
warp, flow / certainty heads; flow and certainty are upsampled x2 (bilinear, align_corners) between levels and the
flow is doubled.  mmcv's ConvModule is Conv2d -> ReLU here (no norm layer is configured in the reference's call,
flow_decoder.py:28-35).

The correlation / lookup / warp primitives are pluggable (`ops`): the default is the CPU oracle's own, the GPU
integration test plugs in picopose_b200's CUDA modules, so the loop runs exactly as the reference's FlowDecoder does
on the overlay.  Module and parameter names follow the reference's state_dict, so weights load across.

Parity pinning (oracle/make_golden.py, `flowdec`): the reference FlowDecoder is built under torch.manual_seed(seed),
its state_dict is loaded into this restatement and both are run on the same inputs (outputs equal to 1e-6); this
restatement built under the same seed reproduces the reference's weights bit for bit (same construction order), which
is what lets tests/golden/flow_decoder.npz carry outputs and weight checksums only (27 M parameters do not fit a fixture).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import corr_lookup_oracle as OL


class ConvAct(nn.Module):
    """mmcv.cnn.ConvModule as the reference configures it: conv (+ ReLU unless act is None); child named `conv`."""

    def __init__(self, cin, cout, k, padding=0, act=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, 1, padding)
        self.act = nn.ReLU() if act else None

    def forward(self, x):
        x = self.conv(x)
        return self.act(x) if self.act is not None else x


class MotionEncoder(nn.Module):
    """'Basic' motion encoder, raft_decoder.py:56-161."""

    def __init__(self, num_levels, radius):
        super().__init__()
        corr_inch = num_levels * (2 * radius + 1) ** 2                                          # :126
        self.corr_net = nn.Sequential(ConvAct(corr_inch, 256, 1, 0), ConvAct(256, 192, 3, 1))   # :127-129
        self.flow_net = nn.Sequential(ConvAct(2, 128, 7, 3), ConvAct(128, 64, 3, 1))            # :131-134
        self.out_net = nn.Sequential(ConvAct(192 + 64, 126, 3, 1))                              # :136-139

    def forward(self, corr, flow, corr_feat0=None):
        """corr_feat0: optional precomputed output of corr_net[0] (the fused lookup + 1x1 conv kernel)."""
        x = self.corr_net[0](corr) if corr_feat0 is None else corr_feat0
        corr_feat = self.corr_net[1](x)                                                         # :157
        flow_feat = self.flow_net(flow)                                                         # :158
        out = self.out_net(torch.cat([corr_feat, flow_feat], dim=1))                            # :160
        return torch.cat([out, flow], dim=1)                                                    # :161


class XHead(nn.Module):
    """raft_decoder.py:251-289."""

    def __init__(self, cin, feat_channels, cout, x):
        super().__init__()
        layers = []
        for ch in feat_channels:
            layers.append(ConvAct(cin, ch, 3, 1))
            cin = ch
        self.layers = nn.Sequential(*layers)
        self.predict_layer = nn.Conv2d(cin, cout, 3, padding=1) if x == "flow" else nn.Conv2d(cin, cout, 1, padding=0)

    def forward(self, x):
        return self.predict_layer(self.layers(x))


class DefaultOps:
    """CPU oracle primitives with the call shapes of the reference modules."""

    def pyramid(self, f1, f2, num_levels):
        return OL.correlation_pyramid(f1, f2, num_levels)

    def lookup(self, pyramid, flow, radius):
        return OL.corr_lookup(pyramid, flow, radius)

    def warp(self, feat, grid_nhw2):
        return OL.bilinear_sample(feat, grid_nhw2, align_corners=True)

    def coords(self, B, H, W, device):
        return OL.coords_grid(B, W, H).to(device)


class FlowDecoder(nn.Module):
    """flow_decoder.py:8-94."""

    def __init__(self, num_levels, radius, ops=None):
        super().__init__()
        self.num_levels = num_levels
        self.r = int(radius / 2)                                                                # :24
        self.ops = ops or DefaultOps()
        proj, enc, fp, mp = [], [], [], []
        for lvl in range(num_levels):                                                           # construction order = :18-40
            proj.append(nn.Sequential(nn.Conv2d(256, 256, 1, 1), nn.BatchNorm2d(256)))
            enc.append(MotionEncoder(lvl + 1, self.r))
            fp.append(XHead(2 * 256 + 128, [512, 256], 2, "flow"))
            mp.append(XHead(2 * 256 + 128, [512, 256], 1, "mask"))
        self.proj, self.encoder = nn.ModuleList(proj), nn.ModuleList(enc)
        self.flow_pred, self.mask_pred = nn.ModuleList(fp), nn.ModuleList(mp)

    def forward_flow(self, feat_render, feat_real, flow, level):
        pyr = self.ops.pyramid(feat_render, feat_real, level + 1)                               # :59
        corr = self.ops.lookup(pyr, flow, self.r)                                               # :61
        if hasattr(self.ops, "encode"):      # lets the CUDA ops fuse the lookup with the encoder's first 1x1 convolution
            motion = self.ops.encode(self.encoder[level], corr, flow)
        else:
            motion = self.encoder[level](corr, flow)                                            # :62
        B, _, H, W = flow.shape
        grid = (self.ops.coords(B, H, W, flow.device) + flow).permute(0, 2, 3, 1)               # :50-54
        warped = self.ops.warp(feat_real, grid)                                                 # :64
        x = torch.cat([feat_render, warped, motion], dim=1)                                     # :66
        return self.flow_pred[level](x), self.mask_pred[level](x)                               # :68-70

    def forward(self, feat_render_list, feat_real_list, init_flow, init_certainty):
        flows, certs = [], []
        flow, cert = init_flow, init_certainty
        for level in range(self.num_levels):
            fr, fl = self.proj[level](feat_render_list[level]), self.proj[level](feat_real_list[level])   # :78
            dflow, dcert = self.forward_flow(fr, fl, flow, level)                               # :80-83 (iters = 1)
            flow, cert = flow + dflow, cert + dcert
            flows.append(flow)
            certs.append(cert)
            if level != self.num_levels - 1:                                                    # :88-92
                flow = 2 * F.interpolate(flow, scale_factor=(2, 2), mode="bilinear", align_corners=True)
                cert = F.interpolate(cert, scale_factor=(2, 2), mode="bilinear", align_corners=True)
        return flows, certs


def decoder_inputs(seed, batch=1):
    """Seeded DPT-shaped inputs: three feature levels (B,256,{16,32,64}^2) per image, a smooth initial flow."""
    g = torch.Generator().manual_seed(seed)
    render = [0.5 * torch.randn(batch, 256, s, s, generator=g) for s in (16, 32, 64)]
    real = [0.5 * torch.randn(batch, 256, s, s, generator=g) for s in (16, 32, 64)]
    flow = 1.5 * torch.randn(batch, 2, 1, 1, generator=g) + 0.5 * torch.randn(batch, 2, 16, 16, generator=g)
    cert = (torch.rand(batch, 1, 16, 16, generator=g) > 0.2).float()
    return render, real, flow, cert


def weight_checksums(module):
    """name -> (sum, abs-sum) in float64: lets a fixture vouch for 27 M seeded parameters without storing them."""
    return {k: (float(v.double().sum()), float(v.double().abs().sum())) for k, v in module.state_dict().items()
            if v.dtype.is_floating_point}
