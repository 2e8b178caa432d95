"""Unrolled reconstruction networks (callers of the SENSE hot path): mirrors of dl_cs/models/{unrolledswin,unrolled,swin3D,resnet3d}.py."""
from .layers import set_conv_precision, get_conv_precision  # noqa: F401
from .resnet3d import ResNet  # noqa: F401
from .swin3d import SwinTransformer3DNet  # noqa: F401
from .unrolled import (HalfQuadraticSplitting, ProximalGradientDescent, SwinHalfQuadraticSplitting,  # noqa: F401
                       SwinProximalGradientDescent, UnrolledNet, UnrolledSwinNet, build_unrolled)
