"""Overlay for the reference's model/stage3/raft_decoder.py.

Everything (XHead, ConvGRU, RAFTDecoder, the MotionEncoder's layers and weights, ...) is the reference's own code,
loaded from the next `model/stage3/raft_decoder.py` on sys.path.  Two things change:

* `CorrelationPyramid` -- the all-pairs product that feeds CorrLookup -- is replaced by ours (same constructor, no
  parameters): it returns a lazy pyramid, so the volume is never built;
* `MotionEncoder` is a subclass of the reference's class (same constructor, same parameters and state_dict keys) whose
  forward lets the first 1x1 convolution of `corr_net` run inside the lookup kernel; every other layer
  runs as the reference's own module.
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

import picopose_b200.correlation as _corr  # noqa: E402
from picopose_b200.correlation import CorrelationPyramid  # noqa: E402,F401  (overrides the reference class)


class MotionEncoder(_ref.MotionEncoder):
    """The reference's MotionEncoder (model/stage3/raft_decoder.py:56-161) with the lookup + corr_net[0] fusion."""

    def forward(self, corr, flow):
        return _corr.motion_encoder_forward(self, corr, flow)


_corr.ENCODER_FUSION = True   # CorrLookup may now hand an un-evaluated lookup to the encoder
