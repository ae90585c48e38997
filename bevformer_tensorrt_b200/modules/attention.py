"""Callers of ``multi_scale_deformable_attn`` (reference: det2trt/models/modules/spatial_cross_attention.py:200-273 and
:694-768, temporal_self_attention.py:350-457). Only ``forward_trt`` is mirrored; weights are plain nn.Linear."""
import torch
from torch import nn

from ..registry import TRT_FUNCTIONS


class MSDeformableAttention3DTRTP(nn.Module):
    """value_proj -> [cams, keys, heads, ch]; Linear(query) -> offsets / logits (no softmax here, it is in the op);
    reference points reshaped to [cams, nq, 1, 2*Z] (spatial_cross_attention.py:747-766)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=8, num_cams=6,
                 op="multi_scale_deformable_attn"):  # fmt: skip
        super().__init__()
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.num_cams = num_cams
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.multi_scale_deformable_attn = TRT_FUNCTIONS.get(op) if isinstance(op, str) else op  # bound once (:692)

    def forward_trt(self, query, value, reference_points, spatial_shapes, bev_mask=None):
        """With ``bev_mask`` (and the fused op registered) returns the camera-summed slots [1, nq, embed_dims] directly:
        MSDA + ``(queries * bev_mask).sum(0)`` in one kernel, the per-camera output is never materialised."""
        value = self.value_proj(value).view(self.num_cams, -1, self.num_heads, self.embed_dims // self.num_heads)
        sampling_offsets = self.sampling_offsets(query)
        attention_weights = self.attention_weights(query)
        reference_points = reference_points.reshape(self.num_cams, -1, 1, reference_points.shape[-2] * 2)
        sampling_offsets = sampling_offsets.view(*sampling_offsets.shape[:2], self.num_heads, -1)
        attention_weights = attention_weights.view(*attention_weights.shape[:2], self.num_heads, -1)
        if bev_mask is not None:
            fused = TRT_FUNCTIONS.get("multi_scale_deformable_attn_sca")
            slots = fused(value, spatial_shapes, reference_points, sampling_offsets, attention_weights, bev_mask)
            return slots.to(query.dtype).unsqueeze(0)
        return self.multi_scale_deformable_attn(value, spatial_shapes, reference_points, sampling_offsets,
                                                attention_weights).flatten(2)  # fmt: skip


class SpatialCrossAttentionTRTP(nn.Module):
    """query repeated per camera, per-camera MSDA, bev_mask-weighted camera sum, output projection, residual
    (spatial_cross_attention.py:248-273)."""

    def __init__(self, embed_dims=256, num_cams=6, fused=False, **attn):
        super().__init__()
        self.embed_dims, self.num_cams, self.fused = embed_dims, num_cams, fused
        self.deformable_attention = MSDeformableAttention3DTRTP(embed_dims=embed_dims, num_cams=num_cams, **attn)
        self.output_proj = nn.Linear(embed_dims, embed_dims)

    def forward_trt(self, query, value, reference_points_cam, bev_mask, spatial_shapes, query_pos=None):
        inp_residual = query
        if query_pos is not None:
            query = query + query_pos
        query = query.repeat(self.num_cams, 1, 1)
        reference_points_cam = reference_points_cam.view(self.num_cams, -1, int(reference_points_cam.size(3)), 2)
        value = value.view(self.num_cams, -1, self.embed_dims)
        if self.fused:  # MSDA + bev_mask camera-sum in one kernel (multi_scale_deformable_attn_sca)
            slots = self.deformable_attention.forward_trt(query, value, reference_points_cam, spatial_shapes, bev_mask)
            return self.output_proj(slots) + inp_residual
        queries = self.deformable_attention.forward_trt(query, value, reference_points_cam, spatial_shapes)
        slots = (queries * bev_mask).sum(0, keepdims=True)
        return self.output_proj(slots) + inp_residual


class TemporalSelfAttentionTRTP(nn.Module):
    """prev/current BEV queue of 2, one level, mean over the queue (temporal_self_attention.py:400-457)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=1, num_points=4, num_bev_queue=2,
                 op="multi_scale_deformable_attn"):  # fmt: skip
        super().__init__()
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.num_bev_queue = num_bev_queue
        self.sampling_offsets = nn.Linear(embed_dims * num_bev_queue, num_bev_queue * num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims * num_bev_queue, num_bev_queue * num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.multi_scale_deformable_attn = TRT_FUNCTIONS.get(op) if isinstance(op, str) else op

    def forward_trt(self, query, reference_points, spatial_shapes, value=None, query_pos=None):
        if value is None:
            value = query.repeat(2, 1, 1)
        identity = query
        if query_pos is not None:
            query = query + query_pos
        query = torch.cat([value[:1], query], -1)
        value = self.value_proj(value).view(self.num_bev_queue, -1, self.num_heads, self.embed_dims // self.num_heads)
        sampling_offsets = self.sampling_offsets(query).view(1, -1, self.num_heads, self.num_bev_queue, self.num_levels,
                                                             self.num_points, 2)  # fmt: skip
        attention_weights = self.attention_weights(query).view(1, -1, self.num_heads, self.num_bev_queue,
                                                               self.num_levels, self.num_points)  # fmt: skip
        attention_weights = attention_weights.permute(0, 3, 1, 2, 4, 5).contiguous()
        sampling_offsets = sampling_offsets.permute(0, 3, 1, 2, 4, 5, 6).contiguous()
        attention_weights = attention_weights.view(*attention_weights.shape[1:3], self.num_heads, -1)
        sampling_offsets = sampling_offsets.view(*sampling_offsets.shape[1:3], self.num_heads, -1)
        output = self.multi_scale_deformable_attn(value, spatial_shapes, reference_points, sampling_offsets,
                                                  attention_weights).flatten(2)  # fmt: skip
        output = torch.mean(output, keepdim=True, dim=0)
        return self.output_proj(output) + identity
