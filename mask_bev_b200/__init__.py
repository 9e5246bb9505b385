"""mask_bev_b200 — B200-native (sm_100a) point-cloud -> BEV front end for MaskBEV.

Public surface mirrors what /root/reference/mask_bev/models/encoders/mask_bev_encoders.py uses:
``Voxelization`` (mmcv.ops), ``PillarFeatureNet`` / ``PointPillarsScatter`` (mmdet3d.models) and the
``MaskBevEncoder`` that chains them. Everything computes in libmask_bev_b200.so (hand-written CUDA behind a
C ABI, include/mask_bev_b200.h); there is no CPU fallback.
"""
from ._lib import MbevError, launch_count  # noqa: F401
from .voxelize import Voxelization  # noqa: F401
from .pillar_encoder import PFNLayer, PillarFeatureNet  # noqa: F401
from .scatter import PointPillarsScatter  # noqa: F401
from .encoder import EncodingType, MaskBevEncoder  # noqa: F401
from .patch_embed import PillarPatchEmbed  # noqa: F401

__all__ = ["Voxelization", "PFNLayer", "PillarFeatureNet", "PointPillarsScatter", "MaskBevEncoder",
           "EncodingType", "PillarPatchEmbed", "MbevError", "launch_count"]
