"""``models.fit_SMPL`` of the reference (src/models/fit_SMPL.py:17-269), served by etch_b200."""
from etch_b200.models.fit_SMPL import fit_smpl, get_markers  # noqa: F401
