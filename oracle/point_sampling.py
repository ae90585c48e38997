"""oracle/point_sampling.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

numpy float32 restatement of the encoder prologue that produces the MSDA reference points and the camera visibility
weights:
    BEVFormerEncoderTRTP.get_reference_points_3d   det2trt/models/modules/encoder.py:168-194
    BEVFormerEncoderTRTP.point_sampling_trt         det2trt/models/modules/encoder.py:196-259
Every step is a separately rounded float32 operation in the reference's order; the 4x4 product is summed left to
right (the reference's comes out of a batched matmul whose order is the backend's, hence a tolerance, not equality).

Pinning: tests/golden/make_golden_point_sampling.py executes the reference's own two methods (their source is read from
/root/reference at generation time, the class around them needs mmcv) and stores their outputs;
tests/test_oracle_golden.py compares this file against them.
"""
import numpy as np

F = np.float32


def linspace_f32(start, end, steps):
    """torch.linspace for float32 (ATen RangeFactories.cpp): step in float, first half from start, second from end."""
    start, end = F(start), F(end)
    if steps == 1:
        return np.array([start], F)
    step = F((end - start) / F(steps - 1))
    i = np.arange(steps)
    lo = (start + step * i.astype(F)).astype(F)
    hi = (end - step * (steps - i - 1).astype(F)).astype(F)
    return np.where(i < steps // 2, lo, hi).astype(F)


def get_reference_points_3d(H, W, Z=8, num_points_in_pillar=4):
    """[1, D, H*W, 3] float32 (encoder.py:168-194)."""
    D = num_points_in_pillar
    zs = (linspace_f32(0.5, F(Z) - F(0.5), D) / F(Z)).astype(F).reshape(D, 1, 1) * np.ones((D, H, W), F)
    xs = (linspace_f32(0.5, F(W) - F(0.5), W) / F(W)).astype(F).reshape(1, 1, W) * np.ones((D, H, W), F)
    ys = (linspace_f32(0.5, F(H) - F(0.5), H) / F(H)).astype(F).reshape(1, H, 1) * np.ones((D, H, W), F)
    return np.stack((xs, ys, zs), -1).reshape(1, D, H * W, 3).astype(F)


def point_sampling(reference_points, pc_range, lidar2img, image_shape, return_cam=False):
    """reference_points [1, D, Q, 3]; lidar2img [cams, 4, 4]; image_shape (h, w).
    Returns reference_points_cam [cams, 1, Q, D, 2], bev_mask [cams, Q, 1] (encoder.py:196-259)."""
    ref = np.asarray(reference_points, F)[0]  # [D, Q, 3]
    ext = np.array([pc_range[3] - pc_range[0], pc_range[4] - pc_range[1], pc_range[5] - pc_range[2]], F)
    lo = np.array(pc_range[:3], F)
    p = ((ref * ext).astype(F) + lo).astype(F)  # [D, Q, 3]
    m = np.asarray(lidar2img, F).reshape(-1, 4, 4)  # [cams, 4, 4]
    x, y, z = p[None, ..., 0], p[None, ..., 1], p[None, ..., 2]  # [1, D, Q]

    def row(r):
        a = m[:, r, :, None, None]  # [cams, 4, 1, 1]
        acc = (a[:, 0] * x).astype(F)
        acc = (acc + (a[:, 1] * y).astype(F)).astype(F)
        acc = (acc + (a[:, 2] * z).astype(F)).astype(F)
        return (acc + a[:, 3]).astype(F)  # [cams, D, Q]

    cx, cy, cz = row(0), row(1), row(2)
    eps = F(1e-5)
    vis = cz > eps
    den = np.maximum(cz, eps)
    u = ((cx / den).astype(F) / F(image_shape[1])).astype(F)
    v = ((cy / den).astype(F) / F(image_shape[0])).astype(F)
    vis &= (v > 0) & (v < 1) & (u < 1) & (u > 0)
    ref_cam = np.stack((u, v), -1).transpose(0, 2, 1, 3)[:, None]  # [cams, 1, Q, D, 2]
    seen = vis.any(1).astype(F)  # [cams, Q]
    bev_mask = (seen / np.maximum(seen.sum(0, keepdims=True), F(1e-4))).astype(F)[..., None]
    if return_cam:
        return np.ascontiguousarray(ref_cam), bev_mask, np.stack((cx, cy, cz), -1)
    return np.ascontiguousarray(ref_cam), bev_mask
