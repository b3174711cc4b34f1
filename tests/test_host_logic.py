"""Host-side logic that needs no GPU: the reference-facing Python surface, argument checking, the overlay
mechanism, sharding arithmetic and the synthetic generators."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_public_surface_mirrors_the_reference():
    import inspect
    from picopose_b200 import corr_lookup, correspondence, matching
    assert list(inspect.signature(matching.matching_templates).parameters)[:5] == \
        ["src_feats", "tar_feat", "src_masks", "tar_mask", "topk"]
    assert inspect.signature(matching.matching_templates).parameters["topk"].default == 5
    assert list(inspect.signature(matching.matching_features_similarity).parameters)[:4] == \
        ["src_feat", "tar_feat", "src_mask", "tar_mask"]
    sig = inspect.signature(corr_lookup.bilinear_sample).parameters
    assert [sig[k].default for k in ("mode", "padding_mode", "align_corners", "scale")] == ["bilinear", "zeros", False, True]
    mod = corr_lookup.CorrLookup()
    assert isinstance(mod, torch.nn.Module) and mod.r == 4 and mod.align_corners is True
    assert len(list(mod.parameters())) == 0 and len(mod.state_dict()) == 0      # checkpoints unaffected
    assert inspect.signature(correspondence.compute_init_correspondences).parameters["size"].default == (16, 16)
    assert inspect.signature(correspondence.compute_stage3_correspondences).parameters["threshold"].default == 0.5


def test_coords_grid_matches_golden():
    from picopose_b200.corr_lookup import coords_grid
    g = np.load(os.path.join(ROOT, "tests", "golden", "bilinear.npz"))
    out = coords_grid(2, torch.arange(0, 7), torch.arange(0, 5))
    np.testing.assert_array_equal(out.numpy(), g["coords"])
    assert out.dtype == torch.float32


def test_cpu_tensors_are_rejected_not_silently_computed():
    from picopose_b200 import corr_lookup, correspondence, matching
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        matching.matching_templates(torch.zeros(1, 2, 8, 4, 4), torch.zeros(1, 8, 4, 4), None, torch.zeros(1, 224, 224))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        matching.matching_features_similarity(torch.zeros(1, 8, 4, 4), torch.zeros(1, 8, 4, 4), torch.zeros(1, 224, 224), None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        corr_lookup.CorrLookup(2)([torch.zeros(16, 1, 4, 4)], torch.zeros(1, 2, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        correspondence.compute_stage3_correspondences(torch.zeros(1, 2, 4, 4), torch.zeros(1, 1, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        corr_lookup.bilinear_sample(torch.zeros(1, 1, 2, 2), torch.zeros(1, 2, 2, 2), mode="nearest")
    with pytest.raises(ValueError):                                          # F.grid_sample rejects unknown modes too
        corr_lookup.bilinear_sample(torch.zeros(1, 1, 2, 2), torch.zeros(1, 2, 2, 2), mode="lanczos")


def test_overlay_shadows_exactly_three_reference_modules(tmp_path):
    """Namespace-package overlay: a fake reference tree with utils/{matching,torch_utils}.py; after
    install_overlay, utils.matching comes from the overlay and utils.torch_utils still from the 'reference'."""
    ref = tmp_path / "ref"
    (ref / "utils").mkdir(parents=True)
    (ref / "utils" / "matching.py").write_text("ORIGIN = 'reference'\n")
    (ref / "utils" / "torch_utils.py").write_text("ORIGIN = 'reference'\n")
    (ref / "model" / "stage3").mkdir(parents=True)
    (ref / "model" / "stage3" / "raft_decoder.py").write_text(
        "ORIGIN = 'reference'\nclass CorrelationPyramid: pass\nclass MotionEncoder: pass\n")
    (ref / "model" / "stage3" / "flow_decoder.py").write_text("ORIGIN = 'reference'\n")
    from picopose_b200 import launcher
    def ours(k):
        return k in ("utils", "model") or k.startswith("utils.") or k.startswith("model.")
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if ours(k)}
    try:
        launcher.install_overlay(str(ref))
        for k in list(sys.modules):
            if ours(k):
                del sys.modules[k]
        m = importlib.import_module("utils.matching")
        t = importlib.import_module("utils.torch_utils")
        c = importlib.import_module("utils.corr_lookup")
        assert m.__file__.startswith(launcher.OVERLAY) and hasattr(m, "matching_templates")
        assert c.__file__.startswith(launcher.OVERLAY) and hasattr(c, "CorrLookup")
        assert t.ORIGIN == "reference"
        # model/stage3/raft_decoder.py: reference module re-exported, CorrelationPyramid swapped
        rd = importlib.import_module("model.stage3.raft_decoder")
        fd = importlib.import_module("model.stage3.flow_decoder")
        assert rd.__file__.startswith(launcher.OVERLAY) and rd.ORIGIN == "reference" and hasattr(rd, "MotionEncoder")
        assert rd.CorrelationPyramid.__module__ == "picopose_b200.correlation"
        assert rd.MotionEncoder.__mro__[1].__module__ == "model.stage3._reference_raft_decoder"   # subclass of the reference's
        assert fd.ORIGIN == "reference"
    finally:
        import picopose_b200.correlation as _c
        _c.ENCODER_FUSION = False
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if ours(k):
                del sys.modules[k]
        sys.modules.update(saved_mods)


def test_shard_range_covers_all_views():
    from picopose_b200.sharded import shard_range
    for n in (42, 162, 642, 5, 3):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(642, r, 4) for r in range(4)] == [(0, 161), (161, 322), (322, 482), (482, 642)]


def test_planted_generator_is_seeded_and_ranked():
    from oracle import matching_oracle as OM
    from picopose_b200 import synth
    a = synth.planted_match_inputs(2, 9, 32, 8, seed=3)
    b = synth.planted_match_inputs(2, 9, 32, 8, seed=3)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    src, tar, planted = a
    score, idx = OM.matching_templates(src, tar, None, synth.disc_mask(2), topk=5)
    assert idx.tolist() == planted[:, :5].tolist()
    assert bool((score[:, :-1] - score[:, 1:] > 1e-3).all())       # gaps far above the 1e-3 exactness band
    m = synth.disc_mask(1)
    assert m[0, 0, 0] == 0 and 0.55 < float(m.mean()) < 0.7


def test_bench_reference_arm_runs_on_cpu():
    import json
    import subprocess
    env = dict(os.environ, PICOPOSE_BENCH_SMALL="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, env=env, check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["metric"] == "detections/sec"


REFERENCE = os.environ.get("PICOPOSE_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "model", "stage3")),
                    reason="reference tree not present (it never is on the GPU box)")
def test_reference_flow_decoder_builds_on_the_overlay():
    """With the real reference tree: the UNMODIFIED model/stage3/flow_decoder.py imports through the overlay and its
    FlowDecoder ends up holding our CorrLookup / CorrelationPyramid modules (mmcv is stubbed: conv stacks only)."""
    import types
    from picopose_b200 import launcher

    def ours(k):
        return k in ("utils", "model", "mmcv") or k.startswith(("utils.", "model.", "mmcv."))
    saved_path, saved_mods = list(sys.path), {k: v for k, v in sys.modules.items() if ours(k)}
    try:
        for k in list(sys.modules):
            if ours(k):
                del sys.modules[k]
        mmcv, cnn = types.ModuleType("mmcv"), types.ModuleType("mmcv.cnn")

        class ConvModule(torch.nn.Module):
            def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, act_cfg=dict(type="ReLU"), **kw):
                super().__init__()
                self.conv = torch.nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding)
                self.act = torch.nn.ReLU() if act_cfg is not None else torch.nn.Identity()

            def forward(self, x):
                return self.act(self.conv(x))

        cnn.ConvModule = ConvModule
        mmcv.cnn = cnn
        sys.modules["mmcv"], sys.modules["mmcv.cnn"] = mmcv, cnn
        launcher.install_overlay(REFERENCE)
        fd = importlib.import_module("model.stage3.flow_decoder")
        assert fd.__file__.startswith(REFERENCE)                         # the reference's own file
        dec = fd.FlowDecoder(num_levels=3, radius=4)
        assert [type(m).__module__ for m in dec.corr_lookup] == ["picopose_b200.corr_lookup"] * 3
        assert [type(m).__module__ for m in dec.corr_block] == ["picopose_b200.correlation"] * 3
        assert dec.corr_lookup[0].r == 2                                  # flow_decoder.py:24 halves the radius
        enc = type(dec.encoder[0])       # our subclass (lookup + first 1x1 conv fusion) of the reference's MotionEncoder
        assert enc.__module__ == "model.stage3.raft_decoder" and enc.__mro__[1].__module__ == "model.stage3._reference_raft_decoder"
        ref_keys = [k for k, _ in enc.__mro__[1](num_levels=1, radius=2, net_type="Basic", conv_cfg=None, norm_cfg=None,
                                                 act_cfg=dict(type="ReLU")).state_dict().items()]
        assert list(dec.encoder[0].state_dict()) == ref_keys                                  # checkpoints load unchanged
        m = importlib.import_module("utils.matching")
        c = importlib.import_module("utils.correspondence")
        assert m.matching_templates.__module__ == "picopose_b200.matching"
        assert c.compute_stage3_correspondences.__module__ == "picopose_b200.correspondence"
        tu = importlib.import_module("utils.torch_utils")
        assert tu.__file__.startswith(REFERENCE)
    finally:
        import picopose_b200.correlation as _c
        _c.ENCODER_FUSION = False
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if ours(k):
                del sys.modules[k]
        sys.modules.update(saved_mods)


def test_bank_handle_passes_the_reference_loop_operations_through():
    """BankHandle: what run_test.py:161-162 and model/picopose.py:99 do to the cached template features
    is index bookkeeping on the handle; anything else raises instead of computing on data it does not have."""
    import torch.nn.functional as F
    from picopose_b200.serving import BankHandle

    class FakeBank:
        n_banks, n_views, C, H, W = 5, 7, 8, 4, 4
        device = torch.device("cpu")

    h = BankHandle(FakeBank())
    assert tuple(h.shape) == (5, 7, 8, 4, 4) and h.dim() == 5 and h.dtype == torch.float32 and h.bank_index is None
    templates_data = {"template_feature": h}
    obj_idx = torch.tensor([[3, 1, 3]]).reshape(-1)
    picked = templates_data["template_feature"][obj_idx].contiguous()        # run_test.py:161-162
    assert isinstance(picked, BankHandle) and tuple(picked.shape) == (3, 7, 8, 4, 4)
    assert picked.bank_index.tolist() == [3, 1, 3] and picked.bank is h.bank
    assert F.normalize(picked, dim=2) is picked                              # model/picopose.py:99
    assert picked[torch.tensor([2, 0])].bank_index.tolist() == [3, 3]        # indices compose
    assert h[1:3].bank_index.tolist() == [1, 2] and picked.to("cpu") is picked and "BankHandle" in repr(picked)
    for bad in (lambda: picked + 1, lambda: F.normalize(picked, dim=1), lambda: picked.sum(), lambda: picked[:, 0],
                lambda: picked[0], lambda: torch.cat([picked, picked])):
        with pytest.raises(NotImplementedError):
            bad()


def test_column_tiling_rule_and_bench_accounting():
    """The contraction cuts tv unmasked query patches into ceil(tv/256) column tiles of round_up(tv/tiles, 32) columns
    (csrc/match_gemm.cu, decode_tile_prefix).  The rule must cover every patch, never exceed the UMMA N limit, never
    produce an empty tile -- and bench.py's issued-FLOP accounting (executed_rows) must follow the same rule."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for tv in range(1, 1025):
        tiles = (tv + 255) // 256
        ncols = (-(-tv // tiles) + 31) // 32 * 32
        assert 32 <= ncols <= 256 and ncols % 32 == 0
        assert tiles * ncols >= tv > (tiles - 1) * ncols
    H, step = 32, 224 // 32
    masks = []
    for tv in (1, 256, 257, 655, 1024):
        m = torch.zeros(224, 224)
        m[::step, ::step][:H, :H] = (torch.arange(H * H) < tv).view(H, H).float()
        masks.append(m)
    rows_exec, rows_unmasked = bench.executed_rows(torch.stack(masks), H)
    assert rows_unmasked == 1 + 256 + 257 + 655 + 1024
    assert rows_exec == 32 + 256 + 2 * 160 + 3 * 224 + 4 * 256


def test_peer_gather_buffer_layout():
    """PeerGather.layout: flag array, then two parities of (tar slots, mask slots); regions are 256-byte aligned, do
    not overlap and the slot offsets used by `gather` stay inside their regions."""
    from picopose_b200.sharded import PeerGather
    for world, tar_b, mask_b in [(2, 4 * 1024 * 32 * 32, 4 * 224 * 224), (8, 4 * 1024 * 32 * 32, 4 * 224 * 224), (3, 4 * 48, 4 * 20)]:
        lay = PeerGather.layout(world, tar_b, mask_b)
        assert lay["flag_bytes"] % 256 == 0 and lay["flag_bytes"] >= 2 * world * 4
        assert lay["mask_off"] % 256 == 0 and lay["mask_off"] >= world * tar_b
        assert lay["parity_bytes"] % 256 == 0 and lay["parity_bytes"] >= lay["mask_off"] + world * mask_b
        assert lay["total"] == lay["flag_bytes"] + 2 * lay["parity_bytes"]
        for par in (0, 1):
            base = lay["flag_bytes"] + par * lay["parity_bytes"]
            assert base + (world - 1) * tar_b + tar_b <= base + lay["mask_off"]
            assert base + lay["mask_off"] + world * mask_b <= lay["flag_bytes"] + (par + 1) * lay["parity_bytes"]


def test_lookup_path_choice_by_radius_and_map_size():
    """CorrLookup on a lazy pyramid: radius 1-2 always takes the no-volume path (TMA-tiled kernel); from radius 3 on the
    per-query kernel is only chosen where it is not slower than pyramid + lookup (launch-bound 16^2, memory-heavy 64^2)."""
    from picopose_b200.correlation import LazyCorrelationPyramid
    def lazy(h, L):
        return LazyCorrelationPyramid(torch.zeros(1, 256, h, h), torch.zeros(1, 256, h, h), L)
    for h, L in ((16, 1), (32, 2), (64, 3)):
        assert lazy(h, L).fusable(1) and lazy(h, L).fusable(2)
    assert lazy(16, 1).fusable(4) and lazy(64, 3).fusable(4)
    assert not lazy(32, 2).fusable(4) and not lazy(32, 2).fusable(3)
    assert not lazy(64, 3).fusable(9)

