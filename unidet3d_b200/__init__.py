"""unidet3d_b200: Blackwell-native (sm_100a) implementation of UniDet3D's forward hot path.

Importing the package registers ``SpConvUNet``, ``UniDet3DEncoder``, ``UniDet3DCriterion`` and ``UniDet3D`` under the
reference's names (reference: unidet3d/__init__.py:1-19).  The CUDA library is loaded lazily by the
first op; it is mandatory -- there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"

from .registry import MODELS  # noqa: F401
from .structures import SparseConvTensor, Det3DDataSample, PointData, InstanceData, DepthInstance3DBoxes  # noqa: F401
from .spconv_unet import SpConvUNet, ResidualBlock  # noqa: F401
from .encoder import UniDet3DEncoder  # noqa: F401
from .criterion import UniDet3DCriterion  # noqa: F401
from .detector import UniDet3D  # noqa: F401
