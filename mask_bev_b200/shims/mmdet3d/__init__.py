"""Shim package: only `mmdet3d.models.{PillarFeatureNet, PointPillarsScatter}` (mask_bev_encoders.py:6)."""
__version__ = "1.1.0+mask_bev_b200"
