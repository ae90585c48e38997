"""The encoder-side prologue that feeds the MSDA callers (reference: det2trt/models/modules/encoder.py:163-330 and
det2trt/models/modules/transformer.py:290-304): prev_bev rotation, pillar grid, camera projection, visibility weights,
2-D reference points. Only the tensor plumbing of ``forward_trt`` up to the layer loop is mirrored; every op on it is
bound once through the function registry, as the reference modules bind theirs."""
import torch
from torch import nn

from ..registry import TRT_FUNCTIONS


class BEVFormerEncoderPrologueTRTP(nn.Module):
    def __init__(self, pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), num_points_in_pillar=4, embed_dims=256,
                 rotate_center=(100, 100)):  # fmt: skip
        super().__init__()
        self.pc_range, self.num_points_in_pillar, self.embed_dims = list(pc_range), num_points_in_pillar, embed_dims
        self.rotate_center = list(rotate_center)
        self.rotate = TRT_FUNCTIONS.get("rotate")
        self.get_reference_points_3d = TRT_FUNCTIONS.get("get_reference_points_3d")
        self.point_sampling_trt = TRT_FUNCTIONS.get("point_sampling_trt")

    def rotate_prev_bev(self, prev_bev, rotation_angle, bev_h, bev_w):
        """transformer.py:296-304: prev_bev [H*W, 1, C] -> rotated, same layout (the permuted view is rotated in its
        channels-last memory layout, the permute back is free)."""
        out = self.rotate(prev_bev.view(bev_h, bev_w, -1).permute(2, 0, 1), rotation_angle,
                          center=prev_bev.new_tensor(self.rotate_center))  # fmt: skip
        return out.permute(1, 2, 0).reshape(bev_h * bev_w, 1, -1)

    def forward_trt(self, bev_query, lidar2img, bev_h, bev_w, image_shape, shift, use_prev_bev):
        """encoder.py:281-303. Returns (hybird_ref_2d [2, Q, 1, 2], reference_points_cam [cams, 1, Q, D, 2],
        bev_mask [cams, Q, 1])."""
        ref_3d = self.get_reference_points_3d(bev_h, bev_w, self.pc_range[5] - self.pc_range[2],
                                              self.num_points_in_pillar, bs=1, device=bev_query.device,
                                              dtype=bev_query.dtype)  # fmt: skip
        ref_2d = ref_3d[0, 0, :, :2].view(1, -1, 1, 2).clone()
        reference_points_cam, bev_mask = self.point_sampling_trt(ref_3d, self.pc_range, lidar2img, image_shape)
        shift_ref_2d = ref_2d + shift.view(1, 1, 1, 2) * use_prev_bev
        return torch.cat([shift_ref_2d, ref_2d], dim=0), reference_points_cam, bev_mask
