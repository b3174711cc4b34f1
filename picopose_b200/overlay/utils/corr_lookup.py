"""Overlay for the reference's utils/corr_lookup.py: same names, B200 kernels underneath."""
from picopose_b200.corr_lookup import CorrLookup, bilinear_sample, coords_grid  # noqa: F401
