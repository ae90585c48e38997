"""Synthetic NuScenes-shaped inputs for the attention-sampling path (host logic, CPU-generated, seeded).

Shapes follow SURVEY.md §8(a)/Appendix D (reference: configs/bevformer/bevformer_base.py:30-35,84-101,
configs/bevformer/bevformer_tiny.py:44-58 and det2trt/models/utils/test_trt_ops/test_multi_scale_deformable_attn.py:7-33).
Two input distributions for the MSDA op:

* ``"U"`` — the reference unit test's: value/offsets/logits ~ N(0,1), reference points ~ U[0,1)
  (test_multi_scale_deformable_attn.py:25-33). Every sampling point lands inside the image: worst case for the gather.
* ``"G"`` — geometry-realistic: reference points are the BEV pillar grid of ``get_reference_points_3d`` projected
  through a synthetic 6-camera ring with the formulas of ``point_sampling_trt``
  (det2trt/models/modules/encoder.py:170-259), so only about one camera in six sees a given BEV query; offsets ~ N(0, 2 px).

Everything is generated on the CPU with a seeded ``torch.Generator`` so vectors are reproducible anywhere.
"""
import math
from dataclasses import dataclass, field
from typing import List, Tuple

import torch


@dataclass(frozen=True)
class MSDAConfig:
    name: str
    batch: int  # cameras for SCA, bev-queue length for TSA, 1 for the decoder
    num_query: int
    num_heads: int
    channels: int  # per head
    spatial_shapes: Tuple[Tuple[int, int], ...]  # (h, w) per level
    num_points: int  # per level
    points_per_group: int  # Z anchors (4 in SCA, 1 in TSA / decoder)
    bev_hw: Tuple[int, int] = field(default=(0, 0))

    @property
    def num_levels(self):
        return len(self.spatial_shapes)

    @property
    def spatial_size(self):
        return sum(h * w for h, w in self.spatial_shapes)

    def algorithmic_bytes(self, elem_bytes: int, ref_bytes: int = None) -> int:
        """Each input read once, output written once; no credit for gather re-reads."""
        ref_bytes = elem_bytes if ref_bytes is None else ref_bytes
        B, Q, M, C = self.batch, self.num_query, self.num_heads, self.channels
        NP = self.num_levels * self.num_points
        value = B * self.spatial_size * M * C * elem_bytes
        ref = B * Q * self.points_per_group * 2 * ref_bytes
        off = B * Q * M * NP * 2 * elem_bytes
        logit = B * Q * M * NP * elem_bytes
        out = B * Q * M * C * elem_bytes
        return value + ref + off + logit + out + self.num_levels * 2 * 4


BASE_LEVELS = ((116, 200), (58, 100), (29, 50), (15, 25))

CONFIGS = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable plumbing case
    "cpu_plumbing": MSDAConfig("cpu_plumbing", 1, 100, 8, 32, ((64, 64),), 4, 1),
    # configs[1]: BEVFormer-tiny SCA op shapes (50x50 BEV, 6 cams, 1 level 15x25, 8 points)
    "tiny_sca": MSDAConfig("tiny_sca", 6, 2500, 8, 32, ((15, 25),), 8, 4, (50, 50)),
    # configs[2]: BEVFormer-base SCA op shapes — the headline (200x200 BEV, 6 cams, 4 levels, 4x8 points)
    "base_sca": MSDAConfig("base_sca", 6, 40000, 8, 32, BASE_LEVELS, 8, 4, (200, 200)),
    # the other MSDA call sites of BEVFormer-base
    "base_tsa": MSDAConfig("base_tsa", 2, 40000, 8, 32, ((200, 200),), 4, 1, (200, 200)),
    "base_decoder": MSDAConfig("base_decoder", 1, 900, 8, 32, ((200, 200),), 4, 1, (200, 200)),
    # small multi-level case for oracle-speed parity tests
    "small_sca": MSDAConfig("small_sca", 2, 333, 8, 32, ((12, 20), (6, 10), (3, 5), (2, 3)), 8, 4, (0, 0)),
}


def camera_ring_lidar2img(num_cams=6, img_hw=(928, 1600), focal=1260.0, radius=1.0, cam_height=1.5):
    """A synthetic NuScenes-like rig: ``num_cams`` pinhole cameras looking outward at 360/num_cams degree yaw spacing.
    Returns lidar2img [num_cams,4,4] (float64 -> float32) such that  (u*d, v*d, d, 1) = lidar2img @ (x, y, z, 1)."""
    H, W = img_hw
    K = torch.tensor([[focal, 0, W / 2, 0], [0, focal, H / 2, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=torch.float64)
    mats = []
    for i in range(num_cams):
        yaw = 2 * math.pi * i / num_cams
        fwd = torch.tensor([math.cos(yaw), math.sin(yaw), 0.0], dtype=torch.float64)
        up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
        right = torch.linalg.cross(fwd, up)  # camera x axis (image right)
        down = -up  # camera y axis (image down)
        R = torch.stack([right, down, fwd])  # world -> camera rotation
        c = radius * fwd + torch.tensor([0.0, 0.0, cam_height], dtype=torch.float64)
        E = torch.eye(4, dtype=torch.float64)
        E[:3, :3] = R
        E[:3, 3] = -R @ c
        mats.append(K @ E)
    return torch.stack(mats).float()


def bev_reference_points_cam(bev_hw, lidar2img, img_hw=(928, 1600), Z=8.0, pillars=4,
                             pc_range=(-51.2, -51.2, -5.0, 51.2, 51.2, 3.0)):  # fmt: skip
    """encoder.py:170-259 restated: pillar grid -> per-camera normalised image coordinates + visibility weights.
    Returns (reference_points_cam [cams, nq, pillars, 2], bev_mask [cams, nq, 1])."""
    H, W = bev_hw
    dev = lidar2img.device
    zs = (torch.linspace(0.5, Z - 0.5, pillars, device=dev) / Z).view(pillars, 1, 1).expand(pillars, H, W)
    xs = (torch.linspace(0.5, W - 0.5, W, device=dev) / W).view(1, 1, W).expand(pillars, H, W)
    ys = (torch.linspace(0.5, H - 0.5, H, device=dev) / H).view(1, H, 1).expand(pillars, H, W)
    ref3d = torch.stack((xs, ys, zs), -1).reshape(pillars, H * W, 3)
    lo = torch.tensor(pc_range[:3], device=dev)
    ext = torch.tensor(pc_range[3:], device=dev) - lo
    pts = torch.cat((ref3d * ext + lo, torch.ones(pillars, H * W, 1, device=dev)), -1)  # [pillars, nq, 4]
    cam = torch.einsum("cij,pqj->cqpi", lidar2img, pts)  # [cams, nq, pillars, 4]
    eps = 1e-5
    depth = cam[..., 2:3]
    vis = (depth > eps).float()
    uv = cam[..., 0:2] / torch.clamp(depth, min=eps)
    uv = uv / torch.tensor([img_hw[1], img_hw[0]], dtype=uv.dtype, device=dev)
    vis = vis * ((uv[..., 0:1] > 0) & (uv[..., 0:1] < 1) & (uv[..., 1:2] > 0) & (uv[..., 1:2] < 1)).float()
    seen = 1 - (1 - vis).prod(2)  # [cams, nq, 1]: any pillar visible
    bev_mask = seen / torch.clamp(seen.sum(0, keepdim=True), min=1e-4)
    return uv, bev_mask


def make_msda_inputs(cfg: MSDAConfig, dist: str = "U", seed: int = 0, dtype=torch.float32):
    """Returns CPU tensors (value[B,S,M,C], spatial_shapes int32[L,2], reference_points[B,Q,1,2G],
    sampling_offsets[B,Q,M,L*P*2], attention logits[B,Q,M,L*P]) in ``dtype`` (float32 or float16)."""
    g = torch.Generator().manual_seed(seed)
    B, Q, M, C = cfg.batch, cfg.num_query, cfg.num_heads, cfg.channels
    L, P, G = cfg.num_levels, cfg.num_points, cfg.points_per_group
    value = torch.randn(B, cfg.spatial_size, M, C, generator=g)
    logits = torch.randn(B, Q, M, L * P, generator=g)
    if dist == "U":
        ref = torch.rand(B, Q, 1, 2 * G, generator=g)
        off = torch.randn(B, Q, M, L * P * 2, generator=g)
    elif dist == "G":
        assert cfg.bev_hw[0] * cfg.bev_hw[1] == Q and G == 4, "distribution G needs an SCA-shaped config"
        uv, _ = bev_reference_points_cam(cfg.bev_hw, camera_ring_lidar2img(B), pillars=G)
        ref = uv.reshape(B, Q, 1, 2 * G).clone()
        # keep fp16 finite: points behind the camera divide by eps and overflow; any value far outside [0,1] is
        # equivalent for the op (every sample is out of range), so clamp like a calibrated fp16 engine would.
        ref = ref.clamp(-60000.0, 60000.0)
        off = 2.0 * torch.randn(B, Q, M, L * P * 2, generator=g)
    elif dist == "edge":
        # stress the range gate and the per-tap validity rules: refs on/over the borders, large offsets
        ref = torch.rand(B, Q, 1, 2 * G, generator=g) * 1.4 - 0.2
        snap = torch.rand(B, Q, 1, 2 * G, generator=g)
        ref = torch.where(snap < 0.15, torch.zeros_like(ref), ref)
        ref = torch.where(snap > 0.85, torch.ones_like(ref), ref)
        off = torch.randn(B, Q, M, L * P * 2, generator=g) * 3.0
        half = torch.rand(B, Q, M, L * P * 2, generator=g) < 0.3
        off = torch.where(half, torch.round(off) + 0.5, off)  # lands exactly on pixel centres / edges
    else:
        raise ValueError(dist)
    shapes = torch.tensor(cfg.spatial_shapes, dtype=torch.int32)
    return value.to(dtype), shapes, ref.to(dtype), off.to(dtype), logits.to(dtype)


def quantize_per_tensor(x: torch.Tensor):
    """MinMax PTQ convention of the reference's op tests (det2trt/models/utils/test_trt_ops/utils.py:18-50):
    scale = amax/127, q = clamp(round(x/scale), -128, 127), real = q*scale."""
    scale = float(x.abs().max()) / 127.0
    scale = scale if scale > 0 else 1.0
    q = torch.clamp(torch.round(x.float() / scale), -128, 127).to(torch.int8)
    return q, scale
