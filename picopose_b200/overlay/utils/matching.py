"""Overlay for the reference's utils/matching.py: same names, B200 kernels underneath."""
from picopose_b200.matching import matching_features_similarity, matching_templates  # noqa: F401
