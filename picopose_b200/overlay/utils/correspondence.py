"""Overlay for the reference's utils/correspondence.py: same names, B200 kernels underneath."""
from picopose_b200.correspondence import (  # noqa: F401
    compute_init_correspondences,
    compute_stage3_correspondences,
    coords_grid,
)
