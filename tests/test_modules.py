"""Integration harness: the SCA / TSA callers drive the op unchanged. CPU part: the plumbing with the oracle port as
the op; GPU part (BASELINE configs[1], tiny SpatialCrossAttention FP16): the sm_100a op inside the module vs the same
module with the oracle op."""
import numpy as np
import pytest
import torch

import bevformer_tensorrt_b200 as bt
from bevformer_tensorrt_b200.modules import SpatialCrossAttentionTRTP, TemporalSelfAttentionTRTP
from bevformer_tensorrt_b200.workloads import bev_reference_points_cam, camera_ring_lidar2img
from oracle import msda as omsda


def _oracle_op(value, shapes, ref, off, logits):
    dev, dt = value.device, value.dtype
    out = omsda.msda_f32(*(t.detach().float().cpu().numpy() for t in (value, shapes.to(torch.float32), ref, off, logits)))
    return torch.from_numpy(out).to(device=dev, dtype=dt)


def _tiny_sca_inputs(dtype=torch.float32, device="cpu"):
    g = torch.Generator().manual_seed(0)
    nq, cams = 2500, 6
    query = torch.randn(1, nq, 256, generator=g)
    value = torch.randn(cams, 375, 256, generator=g)  # 6 x (15*25) x 256
    uv, mask = bev_reference_points_cam((50, 50), camera_ring_lidar2img(cams), img_hw=(480, 800), focal=630.0) \
        if False else bev_reference_points_cam((50, 50), camera_ring_lidar2img(cams))
    ref_cam = uv.view(cams, 1, nq, 4, 2).clamp(-60000, 60000)
    shapes = torch.tensor([[15, 25]], dtype=torch.int64)
    return [t.to(device=device, dtype=dtype) for t in (query, value, ref_cam, mask)] + [shapes.to(device)]


def test_sca_and_tsa_plumbing_cpu():
    torch.manual_seed(0)
    sca = SpatialCrossAttentionTRTP(num_levels=1, num_points=8, op=_oracle_op)
    query, value, ref_cam, mask, shapes = _tiny_sca_inputs()
    out = sca.forward_trt(query, value, ref_cam, mask, shapes)
    assert out.shape == (1, 2500, 256) and torch.isfinite(out).all()
    # the op saw exactly the plugin-signature tensors of BASELINE configs[1]
    seen = {}

    def spy(value, shapes, ref, off, w):
        seen.update(value=value.shape, ref=ref.shape, off=off.shape, w=w.shape)
        return _oracle_op(value, shapes, ref, off, w)

    sca.deformable_attention.multi_scale_deformable_attn = spy
    sca.forward_trt(query, value, ref_cam, mask, shapes)
    assert seen == dict(value=(6, 375, 8, 32), ref=(6, 2500, 1, 8), off=(6, 2500, 8, 16), w=(6, 2500, 8, 8))

    tsa = TemporalSelfAttentionTRTP(op=spy)
    q = torch.randn(1, 400, 256)
    ref2d = torch.rand(2, 400, 1, 2)
    out = tsa.forward_trt(q, ref2d, torch.tensor([[20, 20]]))
    assert out.shape == (1, 400, 256)
    assert seen == dict(value=(2, 400, 8, 32), ref=(2, 400, 1, 2), off=(2, 400, 8, 8), w=(2, 400, 8, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float16, 4e-3)])
def test_tiny_sca_module_on_gpu_matches_oracle_op(dtype, tol):
    torch.manual_seed(0)
    sca = SpatialCrossAttentionTRTP(num_levels=1, num_points=8).cuda().to(dtype)  # op looked up in TRT_FUNCTIONS
    assert sca.deformable_attention.multi_scale_deformable_attn is bt.multi_scale_deformable_attn
    ins = _tiny_sca_inputs(dtype, "cuda")
    with torch.no_grad():
        got = sca.forward_trt(*ins)
        sca.deformable_attention.multi_scale_deformable_attn = _oracle_op
        want = sca.forward_trt(*ins)
    assert got.dtype == dtype and got.shape == (1, 2500, 256)
    assert (got.float() - want.float()).abs().max().item() < tol


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.float16, 4e-3)])
def test_tiny_sca_module_fused_equals_unfused(dtype, tol):
    torch.manual_seed(0)
    sca = SpatialCrossAttentionTRTP(num_levels=1, num_points=8).cuda().to(dtype)
    ins = _tiny_sca_inputs(dtype, "cuda")
    with torch.no_grad():
        want = sca.forward_trt(*ins)
        sca.fused = True
        got = sca.forward_trt(*ins)
    assert got.shape == want.shape and (got.float() - want.float()).abs().max().item() < tol


@pytest.mark.gpu
def test_base_tsa_module_on_gpu_matches_oracle_op():
    torch.manual_seed(1)
    tsa = TemporalSelfAttentionTRTP().cuda()
    q = torch.randn(1, 40 * 40, 256, device="cuda")
    ref2d = torch.rand(2, 40 * 40, 1, 2, device="cuda")
    shapes = torch.tensor([[40, 40]], device="cuda")
    with torch.no_grad():
        got = tsa.forward_trt(q, ref2d, shapes)
        tsa.multi_scale_deformable_attn = _oracle_op
        want = tsa.forward_trt(q, ref2d, shapes)
    assert (got - want).abs().max().item() < 2e-5
