"""Stage-by-stage bring-up of the CUDA kernels on a GPU box, each stage in its own process so that a
trapped kernel cannot poison the next stage (test infrastructure: uses the oracle as checker).   python tests/gpu_debug.py [stage ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["lookup", "prep", "emit1", "emit2", "match1", "match2", "match_big"]


def errmap(diff, blk=32):
    import torch
    T, S = diff.shape
    rows = []
    for i in range(0, T, blk):
        rows.append(" ".join("%8.1e" % float(diff[i:i + blk, j:j + blk].max()) for j in range(0, S, blk)))
    return "\n".join(rows)


def run_stage(stage):
    import numpy as np
    import torch
    from picopose_b200 import _lib, synth
    from picopose_b200 import matching as M
    from oracle import matching_oracle as OM, corr_lookup_oracle as OL
    dev = "cuda:0"
    torch.manual_seed(0)
    if stage == "lookup":
        from picopose_b200.corr_lookup import corr_lookup
        for (B, H, L, r) in [(2, 16, 2, 2), (1, 64, 1, 4), (1, 64, 3, 4), (1, 64, 1, 8)]:
            pyr, flow = synth.lookup_inputs(B, H, L, seed=r, flow_sigma=4.0)
            out = corr_lookup([p.to(dev) for p in pyr], flow.to(dev), r).cpu()
            ref = OL.corr_lookup(pyr, flow, r)
            print(f"lookup B={B} H={H} L={L} r={r}: max err {float((out - ref).abs().max()):.3e}")
    elif stage == "prep":
        x = torch.randn(2, 3, 128, 8, 8)
        ref = torch.nn.functional.normalize(x, dim=2).reshape(2, 3, 128, 64).transpose(2, 3)
        out, rn = M.prepare_features(x.to(dev), "bf16")
        out = out.float().cpu() * rn.cpu().unsqueeze(-1)
        print("prep bf16 max err", float((out - ref).abs().max()), "shape", tuple(out.shape))
    elif stage in ("emit1", "emit2"):
        cl = 1 if stage == "emit1" else 2
        os.environ["PICOPOSE_B200_CLUSTER"] = str(cl)
        for (B, C, H) in [(1, 64, 16), (2, 128, 16), (1, 256, 32), (1, 64, 4)]:
            src = torch.randn(B, C, H, H)
            tar = torch.randn(B, C, H, H)
            T = H * H
            q, qn = M.prepare_features(tar.to(dev), "bf16", True)
            s, sn = M.prepare_features(src.to(dev), "bf16", False)
            q = q.float().cpu() * qn.cpu().unsqueeze(-1)
            s = s.float().cpu() * sn.cpu().unsqueeze(-1)
            ref = torch.clamp(torch.matmul(q, s.transpose(1, 2)), min=0)      # (B,T,S)
            out = M.matching_features_similarity(src.to(dev), tar.to(dev), torch.ones(B, 224, 224, device=dev), None,
                                                 mode="bf16")
            _lib.check_device_faults()
            got = out.view(B, T, H, H).permute(0, 3, 2, 1).reshape(B, T, T).cpu()
            d = (got - ref).abs()
            print(f"{stage} B={B} C={C} H={H}: max err {float(d.max()):.3e} (ref max {float(ref.max()):.3f})")
            if float(d.max()) > 1e-3:
                print(errmap(d[0], 32 if T > 64 else 4))
    elif stage in ("match1", "match2"):
        cl = 1 if stage == "match1" else 2
        for (B, N, C, H) in [(1, 3, 64, 4), (2, 4, 64, 16), (1, 2, 128, 32)]:
            src, tar, planted = synth.planted_match_inputs(B, N, C, H, seed=1)
            mask = synth.disc_mask(B)
            ref, sc_r, it_r, is_r = OM.template_scores(src, tar, mask, want_indices=True)
            for mode in ("bf16", "fp32"):
                got, sc, it, is_ = M.template_scores(src.to(dev), tar.to(dev), mask.to(dev), mode=mode,
                                                     want_indices=True, cluster=cl)
                _lib.check_device_faults()
                print(f"{stage} {mode} B={B} N={N} C={C} H={H}: sim_avg err {float((got.cpu() - ref).abs().max()):.3e} "
                      f"score err {float((sc.cpu() - sc_r).abs().max()):.3e} "
                      f"idx_t2s mismatches {int((it.cpu().long() != it_r).sum())}/{it_r.numel()} "
                      f"idx_s2t mismatches {int((is_.cpu().long() != is_r).sum())}/{is_r.numel()}")
    elif stage == "match_big":
        src, tar, planted = synth.planted_match_inputs(1, 162, 1024, 32, seed=0)
        mask = synth.disc_mask(1).to(dev)
        src, tar = src.to(dev), tar.to(dev)
        bank = M.TemplateBank.from_features(src, "bf16")
        for cl in (1, 2):
            sc, idx = None, None
            sim = M.template_scores(bank, tar, mask, cluster=cl)
            _lib.check_device_faults()
            sc, idx = M.topk_scores(sim, 5)
            print(f"match_big cluster={cl}: top5 {idx[0].tolist()} planted {planted[0, :5].tolist()} scores {sc[0].tolist()}")
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                M.template_scores(bank, tar, mask, cluster=cl)
            t0.record()
            for _ in range(10):
                M.template_scores(bank, tar, mask, cluster=cl)
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / 10
            print(f"match_big cluster={cl}: {ms:.3f} ms per call (warm bank) -> {2 * 162 * 1024 * 1024 * 1024 / ms / 1e9:.1f} TFLOP/s incl. query prep/finalize")
    print(f"[{stage}] done")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_stage(sys.argv[2])
        sys.exit(0)
    stages = sys.argv[1:] or STAGES
    for st in stages:
        print(f"================ {st} ================", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", st], timeout=300,
                               capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print(f"[{st}] FAILED rc={r.returncode}\n{r.stderr[-3000:]}")
        except subprocess.TimeoutExpired:
            print(f"[{st}] TIMEOUT")
