"""Oracle for the stage-3 correlation lookup (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates on the CPU (fp32 torch, explicit gathers -- no F.grid_sample) the
reference's ``utils/corr_lookup.py``:

* ``coords_grid``       <- utils/corr_lookup.py:9-26
* ``bilinear_sample``   <- utils/corr_lookup.py:29-65 (bilinear / zeros padding only)
* ``corr_lookup``       <- CorrLookup.forward, utils/corr_lookup.py:100-134
* ``correlation_pyramid`` <- model/stage3/raft_decoder.py:30-53
* ``grid_sample`` / ``corr_lookup_general`` <- the same two functions for the other interpolation / padding modes the
  reference hands through to F.grid_sample (ATen GridSampler rules written out)

and ``corr_lookup_loops`` (numpy scalar loops, tiny shapes).

Parity pinning: tests/golden/lookup_*.npz (reference outputs, made by
oracle/make_golden.py) are replayed by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def coords_grid(batch: int, W: int, H: int) -> torch.Tensor:
    """(batch, 2, H, W) float grid; channel 0 = x (column), channel 1 = y (row)."""
    ys = torch.arange(H, dtype=torch.float32).view(H, 1).expand(H, W)
    xs = torch.arange(W, dtype=torch.float32).view(1, W).expand(H, W)
    return torch.stack([xs, ys], dim=0)[None].repeat(batch, 1, 1, 1)


def _normalise(pix: torch.Tensor, size: int) -> torch.Tensor:
    # utils/corr_lookup.py:62-63 :  x * 2. / max(size - 1, 1) - 1.
    return (pix * 2.0) / float(max(size - 1, 1)) - 1.0


def _unnormalise(g: torch.Tensor, size: int, align_corners: bool) -> torch.Tensor:
    # ATen grid_sampler_unnormalize
    if align_corners:
        return ((g + 1.0) / 2.0) * float(size - 1)
    return ((g + 1.0) * float(size) - 1.0) / 2.0


def _axis_taps(pix, size, align_corners, scale=True):
    """pixel coords -> (i0 int64, w0, w1) per ATen's bilinear rule."""
    g = _normalise(pix, size) if scale else pix
    ix = _unnormalise(g, size, align_corners)
    i0 = torch.floor(ix)
    w1 = ix - i0
    w0 = (i0 + 1.0) - ix
    return i0.long(), w0, w1


def bilinear_sample(feat, grid, align_corners=False, scale=True):
    """feat (N,C,Hf,Wf); grid (N,Ho,Wo,2) or (N,2,Ho,Wo), pixel coords (x,y) if scale."""
    N, C, Hf, Wf = feat.shape
    if grid.shape[-1] != 2:
        grid = grid.permute(0, 2, 3, 1)
    x0, wx0, wx1 = _axis_taps(grid[..., 0].float(), Wf, align_corners, scale)
    y0, wy0, wy1 = _axis_taps(grid[..., 1].float(), Hf, align_corners, scale)
    flat = feat.reshape(N, C, Hf * Wf)
    out = torch.zeros(N, C, *grid.shape[1:3])

    def tap(xi, yi, w):
        ok = (xi >= 0) & (xi < Wf) & (yi >= 0) & (yi < Hf)
        lin = (yi.clamp(0, Hf - 1) * Wf + xi.clamp(0, Wf - 1)).view(N, 1, -1).expand(N, C, -1)
        v = torch.gather(flat, 2, lin).view(N, C, *xi.shape[1:])
        return v * (w * ok).unsqueeze(1)

    out = tap(x0, y0, wx0 * wy0) + tap(x0 + 1, y0, wx1 * wy0) \
        + tap(x0, y0 + 1, wx0 * wy1) + tap(x0 + 1, y0 + 1, wx1 * wy1)
    return out


# ---- the other argument combinations (utils/corr_lookup.py:29-65 hands mode / padding_mode to F.grid_sample) ----------
# ATen GridSampler rules written out: unnormalise -> padding transform (border: clip; reflection: reflect then clip) ->
# taps.  Bicubic pads every one of its 4 x 4 taps separately and uses the A = -0.75 convolution coefficients; nearest
# rounds half to even.

def _reflect(c: torch.Tensor, twice_low: int, twice_high: int) -> torch.Tensor:
    if twice_low == twice_high:
        return torch.zeros_like(c)
    mn, span = twice_low / 2.0, (twice_high - twice_low) / 2.0
    c = (c - mn).abs()
    extra = torch.fmod(c, span)
    flips = torch.floor(c / span)
    return torch.where(flips % 2 == 0, extra + mn, span - extra + mn)


def _pad_coord(c: torch.Tensor, size: int, padding_mode: str, align_corners: bool) -> torch.Tensor:
    if padding_mode == "border":
        return c.clamp(0, size - 1)
    if padding_mode == "reflection":
        c = _reflect(c, 0, 2 * (size - 1)) if align_corners else _reflect(c, -1, 2 * size - 1)
        return c.clamp(0, size - 1)
    return c


def _fetch(feat: torch.Tensor, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """feat (N,C,Hf,Wf); x, y (N,Ho,Wo) float pixel indices (already integral) -> (N,C,Ho,Wo), 0 outside."""
    N, C, Hf, Wf = feat.shape
    ok = (x >= 0) & (x <= Wf - 1) & (y >= 0) & (y <= Hf - 1)
    xi = x.clamp(0, Wf - 1).long()
    yi = y.clamp(0, Hf - 1).long()
    lin = (yi * Wf + xi).view(N, 1, -1).expand(N, C, -1)
    v = torch.gather(feat.reshape(N, C, Hf * Wf), 2, lin).view(N, C, *x.shape[1:])
    return v * ok.unsqueeze(1)


def _cubic_weights(t: torch.Tensor):
    A = -0.75
    def c1(x): return ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0
    def c2(x): return ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A
    return [c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)]


def grid_sample(feat, grid, mode="bilinear", padding_mode="zeros", align_corners=False, scale=True):
    """bilinear_sample with every mode / padding_mode F.grid_sample has (finite coordinates)."""
    if mode == "bilinear" and padding_mode == "zeros":
        return bilinear_sample(feat, grid, align_corners, scale)
    feat = feat.float()
    N, C, Hf, Wf = feat.shape
    if grid.shape[-1] != 2:
        grid = grid.permute(0, 2, 3, 1)
    gx, gy = grid[..., 0].float(), grid[..., 1].float()
    if scale:
        gx, gy = _normalise(gx, Wf), _normalise(gy, Hf)
    ix, iy = _unnormalise(gx, Wf, align_corners), _unnormalise(gy, Hf, align_corners)
    if mode == "nearest":
        x = torch.round(_pad_coord(ix, Wf, padding_mode, align_corners))    # torch.round: half to even, as nearbyint
        y = torch.round(_pad_coord(iy, Hf, padding_mode, align_corners))
        return _fetch(feat, x, y)
    if mode == "bicubic":
        fx, fy = torch.floor(ix), torch.floor(iy)
        wx, wy = _cubic_weights(ix - fx), _cubic_weights(iy - fy)
        out = 0.0
        for j in range(4):
            y = torch.floor(_pad_coord(fy - 1.0 + j, Hf, padding_mode, align_corners))    # int() truncation of a value >= 0
            row = 0.0
            for k in range(4):
                x = torch.floor(_pad_coord(fx - 1.0 + k, Wf, padding_mode, align_corners))
                xx = fx - 1.0 + k if padding_mode == "zeros" else x
                yy = fy - 1.0 + j if padding_mode == "zeros" else y
                row = row + _fetch(feat, xx, yy) * wx[k].unsqueeze(1)
            out = out + row * wy[j].unsqueeze(1)
        return out
    ix, iy = _pad_coord(ix, Wf, padding_mode, align_corners), _pad_coord(iy, Hf, padding_mode, align_corners)
    fx, fy = torch.floor(ix), torch.floor(iy)
    wx1, wy1 = ix - fx, iy - fy
    wx0, wy0 = (fx + 1.0) - ix, (fy + 1.0) - iy
    return (_fetch(feat, fx, fy) * (wx0 * wy0).unsqueeze(1) + _fetch(feat, fx + 1, fy) * (wx1 * wy0).unsqueeze(1)
            + _fetch(feat, fx, fy + 1) * (wx0 * wy1).unsqueeze(1) + _fetch(feat, fx + 1, fy + 1) * (wx1 * wy1).unsqueeze(1))


def corr_lookup_general(corr_pyramid, flow, radius, mode="bilinear", padding_mode="zeros", align_corners=True):
    """CorrLookup.forward (utils/corr_lookup.py:100-134) for any mode / padding_mode / align_corners."""
    B, _, H, W = flow.shape
    r = int(radius)
    D = 2 * r + 1
    Q = B * H * W
    g = coords_grid(B, W, H) + flow.float()
    centre = g.permute(0, 2, 3, 1).reshape(Q, 1, 1, 2)
    d = torch.linspace(-r, r, D)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, D, D, 2)     # [a, b] = (d[a], d[b]) on (x, y)
    outs = []
    for i, corr in enumerate(corr_pyramid):
        coords = centre / float(2 ** i) + delta
        outs.append(grid_sample(corr, coords, mode, padding_mode, align_corners).view(B, H, W, D * D))
    return torch.cat(outs, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def corr_lookup(corr_pyramid, flow, radius):
    """CorrLookup.forward: pyramid[i] (B*H*W,1,Hi,Wi), flow (B,2,H,W) -> (B, L*D*D, H, W).

    Window channel k = a*D + b samples x + (a-r), y + (b-r): the FIRST window
    axis moves x (utils/corr_lookup.py:116-121,126 -- delta = stack(meshgrid(dy,dx))
    is added to an (x, y) centroid).
    """
    B, _, H, W = flow.shape
    r = int(radius)
    D = 2 * r + 1
    Q = B * H * W
    g = coords_grid(B, W, H) + flow.float()                       # :110-113
    cx = g[:, 0].reshape(Q, 1)
    cy = g[:, 1].reshape(Q, 1)
    d = torch.linspace(-r, r, D).view(1, D)                       # :116-119
    outs = []
    for i, corr in enumerate(corr_pyramid):
        Hi, Wi = corr.shape[-2:]
        px = cx / float(2 ** i) + d                               # (Q,D)  :125-126
        py = cy / float(2 ** i) + d
        x0, wx0, wx1 = _axis_taps(px, Wi, True)
        y0, wy0, wy1 = _axis_taps(py, Hi, True)
        flat = corr.reshape(Q, Hi * Wi).float()

        def tap(xi, yi, wx, wy):
            # xi,wx: (Q,D) over a ; yi,wy: (Q,D) over b  -> (Q,D,D)
            X = xi.view(Q, D, 1).expand(Q, D, D)
            Y = yi.view(Q, 1, D).expand(Q, D, D)
            ok = (X >= 0) & (X < Wi) & (Y >= 0) & (Y < Hi)
            lin = (Y.clamp(0, Hi - 1) * Wi + X.clamp(0, Wi - 1)).reshape(Q, D * D)
            v = torch.gather(flat, 1, lin).view(Q, D, D)
            return v * (wx.view(Q, D, 1) * wy.view(Q, 1, D)) * ok

        o = tap(x0, y0, wx0, wy0) + tap(x0 + 1, y0, wx1, wy0) \
            + tap(x0, y0 + 1, wx0, wy1) + tap(x0 + 1, y0 + 1, wx1, wy1)
        outs.append(o.view(B, H, W, D * D))                       # :130
    out = torch.cat(outs, dim=-1)                                 # :133
    return out.permute(0, 3, 1, 2).contiguous().float()           # :134


def correlation_pyramid(feat1, feat2, num_levels):
    """model/stage3/raft_decoder.py:30-53: all-pairs dot / sqrt(C), 2x2 average pools."""
    N, C, H, W = feat1.shape
    corr = torch.matmul(feat1.reshape(N, C, H * W).transpose(1, 2), feat2.reshape(N, C, H * W))
    corr = corr.reshape(N * H * W, 1, H, W) / torch.sqrt(torch.tensor(float(C)))
    pyr = [corr]
    for _ in range(num_levels - 1):
        c = pyr[-1]
        h2, w2 = c.shape[-2] // 2, c.shape[-1] // 2
        c = c[..., : 2 * h2, : 2 * w2]
        pooled = (c[..., 0::2, 0::2] + c[..., 0::2, 1::2] + c[..., 1::2, 0::2] + c[..., 1::2, 1::2]) * 0.25
        pyr.append(pooled)
    return pyr


def conv1x1_relu(x, weight, bias=None, relu=True):
    """First layer of MotionEncoder.corr_net (model/stage3/raft_decoder.py:113-116,127-129,157): a 1x1 convolution
    corr_inch -> Cout (+ ReLU) over the lookup output x (B, Cin, H, W), written as the per-pixel matrix product it is.
    weight (Cout, Cin) or (Cout, Cin, 1, 1)."""
    w = weight.reshape(weight.shape[0], -1).float()
    y = torch.einsum("oc,bchw->bohw", w, x.float())
    if bias is not None:
        y = y + bias.float().view(1, -1, 1, 1)
    return torch.clamp(y, min=0.0) if relu else y


# ----------------------------------------------------------------------------
# explicit-loop second opinion (tiny shapes only)
# ----------------------------------------------------------------------------

def corr_lookup_loops(corr_pyramid, flow, radius):
    flow = np.asarray(flow, dtype=np.float32)
    B, _, H, W = flow.shape
    r = int(radius)
    D = 2 * r + 1
    L = len(corr_pyramid)
    out = np.zeros((B, L * D * D, H, W), dtype=np.float32)
    f32 = np.float32
    for lvl, corr in enumerate(corr_pyramid):
        vol = np.asarray(corr, dtype=np.float32)
        Hi, Wi = vol.shape[-2:]
        for b in range(B):
            for h in range(H):
                for w in range(W):
                    q = (b * H + h) * W + w
                    cx = f32(f32(w) + flow[b, 0, h, w]) / f32(2 ** lvl)
                    cy = f32(f32(h) + flow[b, 1, h, w]) / f32(2 ** lvl)
                    for a in range(D):
                        for bb in range(D):
                            px = f32(cx + f32(a - r))
                            py = f32(cy + f32(bb - r))
                            gx = f32(f32(f32(px * f32(2)) / f32(max(Wi - 1, 1))) - f32(1))
                            gy = f32(f32(f32(py * f32(2)) / f32(max(Hi - 1, 1))) - f32(1))
                            ix = f32(f32(f32(gx + f32(1)) / f32(2)) * f32(Wi - 1))
                            iy = f32(f32(f32(gy + f32(1)) / f32(2)) * f32(Hi - 1))
                            x0 = int(np.floor(ix))
                            y0 = int(np.floor(iy))
                            acc = f32(0)
                            for (xx, yy, wgt) in (
                                (x0, y0, f32(f32(x0 + 1) - ix) * f32(f32(y0 + 1) - iy)),
                                (x0 + 1, y0, f32(ix - f32(x0)) * f32(f32(y0 + 1) - iy)),
                                (x0, y0 + 1, f32(f32(x0 + 1) - ix) * f32(iy - f32(y0))),
                                (x0 + 1, y0 + 1, f32(ix - f32(x0)) * f32(iy - f32(y0))),
                            ):
                                if 0 <= xx < Wi and 0 <= yy < Hi:
                                    acc = f32(acc + vol[q, 0, yy, xx] * f32(wgt))
                            out[b, lvl * D * D + a * D + bb, h, w] = acc
    return out
