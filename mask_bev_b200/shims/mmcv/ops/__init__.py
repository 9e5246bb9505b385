from mask_bev_b200.voxelize import Voxelization  # noqa: F401  (mask_bev_encoders.py:5)

__all__ = ["Voxelization"]
