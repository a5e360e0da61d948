"""ldic_b200 -- B200 (sm_100a) rate-distortion forward path of
xiaobucc/learning-driven-image-compression-algorithm behind the reference's nn.Module surface.

The directory name carries the reference's (hyphenated) name; import it as ``ldic_b200``
(repo-root shim ``ldic_b200.py``).
"""
from . import _lib, ops                                     # noqa: F401
from . import torch_ops                                     # noqa: F401  (registers torch.ops.ldic.*)
from ._lib import LdicError, EXPORTED_SYMBOLS, lib_path     # noqa: F401
from .layers import (GDN, GaussianConditional, GaussianModel, LowerBound, ModelGDN, ModelIGDN,   # noqa: F401
                     NonNegativeParametrizer, WinBasedAttention, WindowAttention, bypass_round, ste_round,
                     psnr_from_sq_err)
from .transforms import (analysisTransformModel, synthesisTransformModel, h_analysisTransformModel,  # noqa: F401
                         h_synthesisTransformModel)
from .net import Net                                        # noqa: F401
from . import eval as evaluation                            # noqa: F401  (eval_net.py driver)
from .graph import GraphedEvaluator                         # noqa: F401

__all__ = ["GDN", "ModelGDN", "ModelIGDN", "LowerBound", "NonNegativeParametrizer", "GaussianModel",
           "GaussianConditional", "WindowAttention", "WinBasedAttention", "bypass_round", "ste_round", "analysisTransformModel", "synthesisTransformModel",
           "h_analysisTransformModel", "h_synthesisTransformModel", "Net", "ops", "LdicError"]
