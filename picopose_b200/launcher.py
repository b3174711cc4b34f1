"""Runs the UNMODIFIED reference entry point on top of the B200 hot path.

    python -m picopose_b200.launcher /path/to/PicoPose run_test.py --gpus 0 --dataset ycbv ...

`utils` and `model` are namespace packages in the reference (no __init__.py), so putting
picopose_b200/overlay ahead of the reference root on sys.path makes
utils.matching / utils.corr_lookup / utils.correspondence resolve to the overlay while every other
utils.* / model.* module still comes from the reference.
"""
from __future__ import annotations

import os
import runpy
import sys

OVERLAY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "overlay")
PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install_overlay(reference_root: str) -> None:
    """Puts the overlay (and this package) ahead of the reference on sys.path."""
    reference_root = os.path.abspath(reference_root)
    for p in (reference_root, PKG_ROOT, OVERLAY):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in ("utils", "utils.matching", "utils.corr_lookup", "utils.correspondence", "model", "model.stage3",
                 "model.stage3.raft_decoder"):
        sys.modules.pop(name, None)


def main(argv=None) -> None:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        raise SystemExit("usage: python -m picopose_b200.launcher <reference_root> <script.py> [script args...]")
    ref, script = os.path.abspath(argv[0]), argv[1]
    install_overlay(ref)
    os.chdir(ref)  # the reference reads config/, data/, log/ relative to its root
    sys.argv = [script] + argv[2:]
    runpy.run_path(os.path.join(ref, script), run_name="__main__")


if __name__ == "__main__":
    main()
