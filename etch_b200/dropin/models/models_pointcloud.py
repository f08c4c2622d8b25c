"""``models.models_pointcloud`` of the reference (src/models/models_pointcloud.py:18-221), served by etch_b200."""
from etch_b200.models.models_pointcloud import GT_network_equiv  # noqa: F401
