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


@pytest.mark.gpu
def test_encoder_prologue_feeds_sca_on_gpu():
    """rotate(prev_bev) + on-device point sampling + fused SCA, all through registry-bound ops, against the same chain
    fed with the eager-torch geometry (workloads.bev_reference_points_cam) and the unfused SCA module."""
    from bevformer_tensorrt_b200.modules import BEVFormerEncoderPrologueTRTP

    torch.manual_seed(0)
    pro = BEVFormerEncoderPrologueTRTP(rotate_center=(25, 25)).cuda()
    query, value, ref_cam_host, mask_host, shapes = _tiny_sca_inputs(torch.float32, "cuda")
    l2i = camera_ring_lidar2img(6).cuda()
    hyb, ref_cam, bev_mask = pro.forward_trt(query, l2i, 50, 50, (928, 1600), torch.tensor([0.01, -0.02]).cuda(), 1.0)
    assert hyb.shape == (2, 2500, 1, 2) and ref_cam.shape == (6, 1, 2500, 4, 2) and bev_mask.shape == (6, 2500, 1)
    assert torch.allclose(hyb[0] - hyb[1], torch.tensor([0.01, -0.02]).cuda().expand(2500, 1, 2), atol=1e-6)
    assert ((bev_mask > 0) != (mask_host > 0)).float().mean().item() < 1e-3
    sca_f = SpatialCrossAttentionTRTP(num_levels=1, num_points=8, fused=True).cuda()
    sca_u = SpatialCrossAttentionTRTP(num_levels=1, num_points=8).cuda()
    sca_u.load_state_dict(sca_f.state_dict())
    got = sca_f.forward_trt(query, value, ref_cam.clamp(-60000, 60000), bev_mask, shapes)
    want = sca_u.forward_trt(query, value, ref_cam_host, mask_host, shapes)
    bad = ((got - want).abs().amax(-1) > 1e-3).float().mean().item()
    assert bad < 2e-3, bad  # identical but for queries whose visibility flips within rounding of an image border
    # prev_bev rotation through the prologue keeps BEVFormer's [H*W, 1, C] layout and equals the planar op
    prev = torch.randn(2500, 1, 256, device="cuda")
    rot = pro.rotate_prev_bev(prev, torch.tensor(3.0, device="cuda"), 50, 50)
    assert rot.shape == prev.shape
    planar = bt.rotate(prev.view(50, 50, 256).permute(2, 0, 1).contiguous(), torch.tensor(3.0, device="cuda"),
                       torch.tensor([25.0, 25.0], device="cuda"))
    assert torch.equal(rot.view(50, 50, 256).permute(2, 0, 1), planar)


# ---------------------------------------------------------------------------------------------------------------
# DCNv2P / DCNv2P2 conv layers (det2trt/models/modules/cnn/dcn.py:31-164)
# ---------------------------------------------------------------------------------------------------------------
def test_dcnv2p_layer_structure_and_checkpoint_keys():
    from bevformer_tensorrt_b200.modules import CONV_LAYERS, ModulatedDeformConv2dPackPlugin2

    layer = CONV_LAYERS["DCNv2P"](16, 12, 3, stride=1, padding=1, groups=2, deform_groups=2)
    assert sorted(k for k, _ in layer.named_parameters()) == ["bias", "conv_offset.bias", "conv_offset.weight", "weight"]
    assert layer.weight.shape == (12, 8, 3, 3) and layer.conv_offset.out_channels == 2 * 3 * 9
    assert layer.conv_offset.weight.abs().max() == 0 and layer.conv_offset.bias.abs().max() == 0  # dcn.py:62-66
    assert layer.bias.abs().max() == 0 and 0 < layer.weight.abs().max() <= 1.0 / (16 * 9) ** 0.5 + 1e-6
    assert CONV_LAYERS["DCNv2P2"] is ModulatedDeformConv2dPackPlugin2 and ModulatedDeformConv2dPackPlugin2._version == 2
    other = CONV_LAYERS["DCNv2P2"](16, 12, 3, padding=1, groups=2, deform_groups=2, bias=False)
    assert other.bias is None
    missing = other.load_state_dict({k: v for k, v in layer.state_dict().items() if k != "bias"})
    assert not missing.missing_keys and torch.equal(other.weight, layer.weight)
    # pre-version-2 checkpoints named the offset conv "<layer>_offset" (dcn.py:101-118)
    sd = {"weight": layer.weight.data, "bias": layer.bias.data, "_offset.weight": torch.ones(54, 16, 3, 3),
          "_offset.bias": torch.ones(54)}
    layer._load_from_state_dict(sd, ".", {}, True, [], [], [])
    assert ".conv_offset.weight" in sd and ".conv_offset.bias" in sd and "_offset.weight" not in sd  # renamed in place
    with pytest.raises(ValueError):
        CONV_LAYERS["DCNv2P"](10, 12, 3, groups=4)
