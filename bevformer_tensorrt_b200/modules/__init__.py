"""mmcv-free integration harness: the three callers of the MSDA op, with the tensor plumbing of the reference's
``forward_trt`` methods and plain ``nn.Linear`` layers (mmcv / its registries are absent here). They exist to show the
operator is called *unchanged* — bound once at construction, invoked positionally — and to measure the op in situ
(BASELINE configs[1]: tiny SpatialCrossAttention). They are not a port of the model graph."""
from .attention import MSDeformableAttention3DTRTP, SpatialCrossAttentionTRTP, TemporalSelfAttentionTRTP
from .cnn import CONV_LAYERS, ModulatedDeformConv2dPackPlugin, ModulatedDeformConv2dPackPlugin2
from .encoder import BEVFormerEncoderPrologueTRTP

__all__ = ["BEVFormerEncoderPrologueTRTP", "CONV_LAYERS", "MSDeformableAttention3DTRTP",
           "ModulatedDeformConv2dPackPlugin", "ModulatedDeformConv2dPackPlugin2", "SpatialCrossAttentionTRTP",
           "TemporalSelfAttentionTRTP"]
