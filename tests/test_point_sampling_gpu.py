"""GPU parity tests for the on-device BEV point sampling (pytest -m gpu). Checkers: oracle/point_sampling.py (the
reference's encoder prologue restated in float32) and golden outputs of the reference's own methods. Tolerances:
reference_points_cam 2e-5 relative (the 4x4 product's summation order is the backend's in the reference); bev_mask
identical except for queries with a pillar point within that rounding of an image border."""
import ctypes
import os

import numpy as np
import pytest
import torch

import bevformer_tensorrt_b200 as bt
from bevformer_tensorrt_b200 import _lib
from bevformer_tensorrt_b200.workloads import CONFIGS, make_msda_inputs
from oracle import point_sampling as ops
from tests.helpers import GOLDEN, POINT_SAMPLING_CASES, make_point_sampling_inputs

pytestmark = pytest.mark.gpu


def close(cam, mask, want_cam, want_mask, rtol=2e-5):
    assert cam.shape == want_cam.shape and mask.shape == want_mask.shape
    rel = np.abs(cam - want_cam) / np.maximum(1.0, np.abs(want_cam))
    assert rel.max() < rtol, rel.max()
    assert (mask != want_mask).any(0).mean() < 2e-3


@pytest.mark.parametrize("case", list(POINT_SAMPLING_CASES))
def test_matches_oracle_and_reference_golden(case):
    z = np.load(os.path.join(GOLDEN, "point_sampling_ref.npz"))
    H, W, D, img_hw, pc_range = POINT_SAMPLING_CASES[case]
    l2i = make_point_sampling_inputs(case)
    # the encoder's call sequence (encoder.py:281-295) with the reference's function names
    # built on the CPU it is the golden tensor bit for bit; built on the GPU, torch divides by a Python scalar as a
    # multiplication with its reciprocal (1 ulp), which is what the reference's eager path does on a GPU as well
    ref3d = bt.get_reference_points_3d(H, W, pc_range[5] - pc_range[2], D, device="cpu")
    assert np.array_equal(ref3d.numpy().reshape(-1), z[f"{case}_ref3d"].reshape(-1))
    on_gpu = bt.get_reference_points_3d(H, W, pc_range[5] - pc_range[2], D, device="cuda")
    assert (on_gpu.cpu() - ref3d).abs().max().item() < 1e-6
    ref3d = ref3d.cuda()
    cam, mask = bt.point_sampling_trt(ref3d, pc_range, l2i.cuda(), img_hw)
    assert cam.shape == (6, 1, H * W, D, 2) and mask.shape == (6, H * W, 1)
    want_cam, want_mask = ops.point_sampling(ref3d.cpu().numpy(), pc_range, l2i.numpy(), img_hw)
    # same operations in the same order as the oracle: bit-identical
    assert np.array_equal(cam.cpu().numpy(), want_cam) and np.array_equal(mask.cpu().numpy(), want_mask)
    close(cam.cpu().numpy(), mask.cpu().numpy(), z[f"{case}_cam"], z[f"{case}_mask"])
    # fused form: pillar grid generated in registers -> bit-identical to reading the tensor
    cam2, mask2 = bt.bev_point_sampling(H, W, pc_range, l2i.cuda(), img_hw, D)
    assert torch.equal(cam2, cam) and torch.equal(mask2, mask)


def test_fp16_outputs_are_rounded_fp32_and_finite():
    H, W, D, img_hw, pc_range = POINT_SAMPLING_CASES["ring_small"]
    l2i = make_point_sampling_inputs("ring_small").cuda()
    cam32, mask32 = bt.bev_point_sampling(H, W, pc_range, l2i, img_hw, D)
    cam16, mask16 = bt.bev_point_sampling(H, W, pc_range, l2i, img_hw, D, dtype=torch.float16)
    assert cam16.dtype == torch.float16 and mask16.dtype == torch.float16
    assert torch.isfinite(cam16).all()  # behind-camera points (|u| ~ 1e8) saturate instead of becoming inf
    assert torch.equal(cam16, cam32.clamp(-65504.0, 65504.0).half())
    assert torch.equal(mask16, mask32.half())
    # half reference_points in (an fp16 model's ref_3d): same kernel, inputs rounded to half first
    ref3d = bt.get_reference_points_3d(H, W, 8.0, D, device="cuda", dtype=torch.float16)
    cam_h, mask_h = bt.point_sampling_trt(ref3d, pc_range, l2i, img_hw)
    want_cam, want_mask = ops.point_sampling(ref3d.float().cpu().numpy(), pc_range, l2i.cpu().numpy(), img_hw)
    want_h = np.clip(want_cam, -65504, 65504)
    assert (np.abs(cam_h.float().cpu().numpy() - want_h) <= 1e-3 * np.maximum(1.0, np.abs(want_h))).all()  # half ulp
    assert (mask_h.float().cpu().numpy() != want_mask).any(0).mean() < 2e-3


def test_base_size_feeds_the_fused_sca_kernel():
    """BEVFormer-base: 200x200 BEV, 6 cameras, 4 pillar points. The kernel's outputs go straight into the fused SCA
    sampling op; the result equals the one computed from the oracle's reference points and mask."""
    cfg = CONFIGS["base_sca"]
    H, W = cfg.bev_hw
    pc_range, img_hw = (-51.2, -51.2, -5.0, 51.2, 51.2, 3.0), (928, 1600)
    from bevformer_tensorrt_b200.workloads import camera_ring_lidar2img

    l2i = camera_ring_lidar2img(6)
    cam, mask = bt.bev_point_sampling(H, W, pc_range, l2i.cuda(), img_hw, 4, dtype=torch.float16)
    want_cam, want_mask = ops.point_sampling(ops.get_reference_points_3d(H, W, 8.0, 4), pc_range, l2i.numpy(), img_hw)
    vis = (mask[..., 0] > 0).float().mean().item()
    assert 0.15 < vis < 0.30  # ~21 % of camera x query pairs are visible (distribution G)
    assert ((mask.float().cpu().numpy() > 0) != (want_mask > 0)).mean() < 1e-4
    value, shapes, _, off, logits = (t.cuda() for t in make_msda_inputs(cfg, "U", 0, torch.float16))
    ref_dev = cam.view(6, H * W, 1, 8)
    ref_orc = torch.from_numpy(np.clip(want_cam, -65504, 65504)).half().cuda().view(6, H * W, 1, 8)
    got = bt.multi_scale_deformable_attn_sca(value, shapes, ref_dev, off, logits, mask.view(6, -1))
    want = bt.multi_scale_deformable_attn_sca(value, shapes, ref_orc, off, logits,
                                              torch.from_numpy(want_mask).half().cuda().view(6, -1))
    torch.cuda.synchronize()
    # identical except at the handful of queries whose visibility or reference point differs by rounding
    diff = (got.float() - want.float()).abs().amax(-1).view(-1)
    assert (diff > 1e-3).float().mean().item() < 1e-3


def test_errors():
    l2i = make_point_sampling_inputs("ring_small")
    with pytest.raises(RuntimeError):
        bt.point_sampling_trt(torch.rand(1, 4, 10, 3), (-1, -1, -1, 1, 1, 1), l2i, (928, 1600))  # CPU tensor
    with pytest.raises(ValueError):
        bt.point_sampling_trt(torch.rand(4, 10, 3).cuda(), (-1, -1, -1, 1, 1, 1), l2i, (928, 1600))
    with pytest.raises(ValueError):
        bt.bev_point_sampling(5, 5, (-1, -1, 1, 1), l2i.cuda(), (928, 1600))
    with pytest.raises(_lib.B200OpsError):  # more than 8 pillar points
        bt.bev_point_sampling(5, 5, (-1, -1, -1, 1, 1, 1), l2i.cuda(), (928, 1600), 9)
    with pytest.raises(_lib.B200OpsError):
        bt.bev_point_sampling(5, 5, (-1, -1, -1, 1, 1, 1), l2i.cuda(), (928, 1600), 4, dtype=torch.float64)
    lib = _lib.load()
    pcr = (ctypes.c_double * 6)(-1, -1, -1, 1, 1, 1)
    x = torch.empty(64, device="cuda")
    assert lib.b200_bev_point_sampling(None, pcr, None, 6, 928, 1600, 5, 5, 4, 0, x.data_ptr(), x.data_ptr(), None) == 2
    assert lib.b200_bev_point_sampling(None, pcr, x.data_ptr(), 6, 928, 1600, 0, 5, 4, 0, x.data_ptr(), x.data_ptr(),
                                       None) == 2  # fmt: skip
