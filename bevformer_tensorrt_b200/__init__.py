"""bevformer_tensorrt_b200 — BEVFormer's attention-sampling hot path (multi-scale deformable attention, grid sampler,
modulated deformable conv) as hand-written sm_100a CUDA behind the operator interface of DerryHub/BEVFormer_tensorrt.

Importing the package loads lib/libb200_bev_ops.so (built in-tree by ``python -m bevformer_tensorrt_b200.build``) and
fails loudly if it is missing — there is no CPU or PyTorch fallback on this path.
"""
from . import _lib

_lib.load()

from .functions import (  # noqa: E402
    TRT_FUNCTIONS,
    multi_scale_deformable_attn,
    multi_scale_deformable_attn2,
    multi_scale_deformable_attn_int8,
)

__all__ = [
    "TRT_FUNCTIONS",
    "multi_scale_deformable_attn",
    "multi_scale_deformable_attn2",
    "multi_scale_deformable_attn_int8",
]
__version__ = "0.1.0"
