"""Drop-in mirror of the reference's point-voxel `modules` package
(/root/reference/experiments/model/pvcnn/modules/__init__.py:1-8 == experiments/pvd/modules/):
the same thirteen class names with the same constructor arguments, parameter names (state_dict
keys) and forward signatures, so reference checkpoints load and the reference's network builders
(pvcnn_utils.py:72-168) work unchanged on top of the B200 kernels."""
from . import functional  # noqa: F401  (reference modules do `from . import functional as F`)
from .layers import SE3d, Attention, KLLoss, SharedMLP, Swish
from .point_voxel import PVConv, PVConvReLU, Voxelization
from .pointnet2 import BallQuery, FrustumPointNetLoss, PointNetAModule, PointNetFPModule, PointNetSAModule

__all__ = ['BallQuery', 'FrustumPointNetLoss', 'KLLoss', 'PointNetAModule', 'PointNetSAModule',
           'PointNetFPModule', 'PVConv', 'Attention', 'Swish', 'PVConvReLU', 'SE3d', 'SharedMLP', 'Voxelization']
