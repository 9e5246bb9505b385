from mask_bev_b200.pillar_encoder import PFNLayer, PillarFeatureNet  # noqa: F401
from mask_bev_b200.scatter import PointPillarsScatter  # noqa: F401

__all__ = ["PillarFeatureNet", "PointPillarsScatter", "PFNLayer"]
