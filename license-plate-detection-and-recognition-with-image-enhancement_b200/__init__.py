"""lpsr_b200 -- B200-native (sm_100a) LPSR forward pass behind the reference's own ``LPSR`` nn.Module interface.

The importable name of this package is ``lpsr_b200`` (the directory name contains hyphens; the tiny shim package
``lpsr_b200/`` at the repository root points Python at this directory).
"""
from . import capi  # noqa: F401
from .model import LPSR  # noqa: F401
from .ops import conv2d, enhance_plates, non_max_suppression, pixel_shuffle2, pixel_unshuffle2, preprocess_for_sr_batch  # noqa: F401
from .parallel import forward_sharded, gather_outputs, shard_bounds  # noqa: F401
from . import evaluation, pipeline  # noqa: F401
from .pipeline import enhance_frames, format_long_plate, restack_to_square  # noqa: F401

__all__ = ["LPSR", "capi", "conv2d", "pixel_shuffle2", "pixel_unshuffle2", "preprocess_for_sr_batch", "enhance_plates", "non_max_suppression", "forward_sharded", "gather_outputs", "shard_bounds", "pipeline", "evaluation", "enhance_frames", "format_long_plate", "restack_to_square"]
