"""depth_from_motion_b200 -- B200-native DfM plane-sweep cost-volume path.

Host-side mirror of the reference's plugin interface (registry + module
classes) over the C-ABI CUDA library ``libdfm_b200.so`` (``include/dfm_b200.h``).
"""
from . import capi, checkpoint  # noqa: F401
from .modules import (CostLogits, DepthHead, DfMBackbone,  # noqa: F401
                      DfMNeck, FrustumToVoxel, OutdoorImVoxelNeck, build_dfm_cost, conv3d,
                      multiview_lift)
from .registry import (BACKBONES, HEADS, NECKS, Config,  # noqa: F401
                       build_backbone, build_head, build_neck,
                       register_into_mmdet)

__version__ = '0.1.0'
