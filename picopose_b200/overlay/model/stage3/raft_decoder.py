"""Overlay for the reference's model/stage3/raft_decoder.py.

Everything (MotionEncoder, XHead, ConvGRU, RAFTDecoder, ...) is the reference's own code, loaded from the
next `model/stage3/raft_decoder.py` on sys.path; only `CorrelationPyramid` -- the all-pairs product that
feeds CorrLookup -- is replaced by the tensor-core implementation (same constructor, no parameters).
"""
import importlib.util
import os
import sys

_HERE = os.path.abspath(__file__)


def _load_reference_module():
    for base in sys.path:
        cand = os.path.abspath(os.path.join(base or ".", "model", "stage3", "raft_decoder.py"))
        if cand != _HERE and os.path.isfile(cand):
            spec = importlib.util.spec_from_file_location("model.stage3._reference_raft_decoder", cand)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = mod
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("reference model/stage3/raft_decoder.py not found on sys.path (run through picopose_b200.launcher)")


_ref = _load_reference_module()
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})

from picopose_b200.correlation import CorrelationPyramid  # noqa: E402,F401  (overrides the reference class)
