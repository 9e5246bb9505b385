"""Shim package: only `mmcv.ops.Voxelization` (the one mmcv symbol on MaskBEV's encoder path)."""
__version__ = "2.0.0+mask_bev_b200"
