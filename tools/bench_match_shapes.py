"""Stage-1 matching across the shapes of BASELINE.json / the reference's native operating point, resident banks.

    python tools/bench_match_shapes.py [--json out.json]

Reports ms per call of matching_templates(TemplateBank, ...) (query prologue + contraction + finalisation + top-k) and
the dense-formula rate 2*B*N*T*S*C / t."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SHAPES = [  # (label, B, N, C, H)
    ("native run_test.py: bs 4 x 162 views, 16x16, C=1024", 4, 162, 1024, 16),
    ("config 1 shape: 1 x 42, 32x32, C=384", 1, 42, 384, 32),
    ("config 2: 1 x 162, 32x32, C=1024", 1, 162, 1024, 32),
    ("8 x 642, 32x32, C=1024", 8, 642, 1024, 32),
    ("native grid, big batch: 64 x 162, 16x16, C=1024", 64, 162, 1024, 16),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    from picopose_b200 import matching as M
    from picopose_b200 import synth
    dev = "cuda:0"
    rows = []
    for label, B, N, C, H in SHAPES:
        g = torch.Generator(device=dev).manual_seed(0)
        bank_f = torch.randn(1, N, C, H, H, device=dev, generator=g)
        tar = bank_f[0, :B].clone() + 0.5 * torch.randn(B, C, H, H, device=dev, generator=g)
        bank = M.TemplateBank.from_features(bank_f)
        del bank_f
        bidx = torch.zeros(B, dtype=torch.int32, device=dev)
        for mask_name, mask in (("disc", synth.disc_mask(B).to(dev)), ("ones", torch.ones(B, 224, 224, device=dev))):
            for _ in range(3):
                s, i = M.matching_templates(bank, tar, None, mask, topk=5, bank_index=bidx)
            torch.cuda.synchronize()
            assert i[:, 0].tolist() == list(range(B))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                M.matching_templates(bank, tar, None, mask, topk=5, bank_index=bidx)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.iters
            T = H * H
            row = {"shape": label, "mask": mask_name, "ms_per_call": ms, "detections_per_s": B * 1e3 / ms,
                   "dense_formula_TFLOPs": 2.0 * B * N * T * T * C / ms / 1e9}
            rows.append(row)
            print(json.dumps(row), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
