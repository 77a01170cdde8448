"""Host-side mirror of the reference's hot-path plugin interface (reference daod/modeling/*), backed by
libsfod_b200.so.  Importing this package registers the plugins under the reference's registry names."""
from .box_regression import Box2BoxTransform  # noqa: F401
from .anchor_generator import DefaultAnchorGenerator  # noqa: F401
from .batch_norm import SfodBatchNorm2d, convert_batchnorm  # noqa: F401
from .poolers import ROIPooler  # noqa: F401
from .proposal_generator import RPN, DARPN, PseudoLabRPN, StandardRPNHead  # noqa: F401
from .fast_rcnn import FastRCNNOutputLayers, SourceFreeFastRCNNOutputLayers  # noqa: F401
from .roi_heads import (FastRCNNConvFCHead, SourceFreeAdaptiveTeacherStandardROIHeads,  # noqa: F401
                        SourceFreeAdaptiveTeacherEvalStandardROIHeads, AdaptiveTeacherStandardROIHeads, build_box_head)
from .vgg import build_vgg_backbone, vgg_backbone  # noqa: F401
from .resnet import build_resnet_backbone, ResNet, BottleneckBlock, BasicStem, FrozenBatchNorm2d  # noqa: F401
from .meta_arch import SourceFreeAdaptiveTeacherGeneralizedRCNN  # noqa: F401
