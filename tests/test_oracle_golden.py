"""CPU gate: the oracle (oracle/msda_oracle.c and the torch port) against golden vectors produced by the
reference's own Python code (tests/golden/make_golden_msda.py)."""
import numpy as np
import pytest
import torch

from oracle import msda as omsda
from tests.helpers import golden_msda_cases, load_golden_msda


@pytest.mark.parametrize("path", golden_msda_cases())
def test_c_oracle_matches_reference_python(path):
    cfg, (value, shapes, ref, off, logits), want = load_golden_msda(path)
    got = omsda.msda_f32(value.numpy(), shapes.numpy(), ref.numpy(), off.numpy(), logits.numpy())
    # fp32 vs fp32, two formulations of the same math (grid_sample normalises then un-normalises coordinates)
    assert np.abs(got - want).max() < 2e-5, np.abs(got - want).max()


@pytest.mark.parametrize("path", golden_msda_cases())
def test_torch_port_matches_reference_python(path):
    cfg, (value, shapes, ref, off, logits), want = load_golden_msda(path)
    got = omsda.msda_torch_port(value, shapes, ref, off, logits).numpy()
    assert np.abs(got - want).max() < 2e-6, np.abs(got - want).max()


def test_index_records_are_consistent():
    cfg, (value, shapes, ref, off, logits), _ = load_golden_msda(golden_msda_cases()[0])
    out, idx = omsda.msda_f32(value.numpy(), shapes.numpy(), ref.numpy(), off.numpy(), logits.numpy(), True)
    assert idx.shape == (cfg.batch, cfg.num_query, cfg.num_heads, cfg.num_levels * cfg.num_points)
    inr = idx["in_range"].astype(bool)
    assert inr.any()
    H, W = cfg.spatial_shapes[0]
    assert (idx["h_low"][inr] >= -1).all() and (idx["h_low"][inr] <= H - 1).all()
    assert (idx["w_low"][inr] >= -1).all() and (idx["w_low"][inr] <= W - 1).all()
    assert (idx["tap_mask"][~inr] == 0).all()


# ---------------------------------------------------------------------------------------------------------------
# golden vectors produced ON A B200 by the reference's own CUDA kernels (oracle/_ref, tests/golden/make_golden_ref_gpu.py)
# ---------------------------------------------------------------------------------------------------------------
import glob as _glob  # noqa: E402
import os as _os  # noqa: E402

from bevformer_tensorrt_b200.workloads import MSDAConfig, make_msda_inputs, quantize_per_tensor  # noqa: E402
from tests.helpers import GOLDEN as _GOLDEN, input_digest  # noqa: E402

_REF_GPU = sorted(_glob.glob(_os.path.join(_GOLDEN, "ref_gpu_*.npz")))


@pytest.mark.skipif(not _REF_GPU, reason="no tests/golden/ref_gpu_*.npz yet (generated on a GPU box)")
@pytest.mark.parametrize("path", _REF_GPU)
def test_c_oracle_matches_reference_cuda_kernels(path):
    """The FP32 restatement against ms_deformable_im2col_cuda<float>, and the quantised-intermediate emulation against
    ms_deformable_im2col_cuda_int8<float> (…Kernel.cu:1106-1128, :1172-1194): outputs of the reference's own kernels,
    compiled unmodified and run on a B200, committed as fixtures."""
    z = np.load(path)
    B, Q, M, C, L, P, G, seed = (int(x) for x in z["meta"])
    cfg = MSDAConfig("g", B, Q, M, C, tuple((int(h), int(w)) for h, w in z["shapes"]), P, G)
    value, shapes, ref, off, logits = make_msda_inputs(cfg, str(z["dist"]), seed, torch.float32)
    assert input_digest(value, shapes, ref, off, logits) == str(z["digest"])
    got = omsda.msda_f32(value.numpy(), shapes.numpy(), ref.numpy(), off.numpy(), logits.numpy())
    assert np.abs(got - z["out_f32"]).max() < 1e-5, np.abs(got - z["out_f32"]).max()
    sv, so, sw, sout = (float(x) for x in z["scales"])
    vq, sv2 = quantize_per_tensor(value)
    oq, so2 = quantize_per_tensor(off)
    wq, sw2 = quantize_per_tensor(logits)
    assert (sv, so, sw) == (sv2, so2, sw2)
    emu = omsda.msda_i8_refemu(vq.numpy(), sv, shapes.numpy(), ref.numpy(), oq.numpy(), so, wq.numpy(), sw, sout)
    d = np.abs(emu.astype(np.int32) - z["out_i8_f32ref"].astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 0.02, (d.max(), (d != 0).mean())
    # our INT8 definition (in-register dequantisation, one requantisation) vs the reference kernel, which requantises every
    # bilinear sample and the softmax weights to int8 first (…Kernel.cu:926-947): at most 2 output LSB apart
    deq = omsda.msda_i8_dequant(vq.numpy(), sv, shapes.numpy(), ref.numpy(), oq.numpy(), so, wq.numpy(), sw, sout)
    assert np.abs(deq.astype(np.int32) - z["out_i8_f32ref"].astype(np.int32)).max() <= 2


# ---------------------------------------------------------------------------------------------------------------
# grid sampler
# ---------------------------------------------------------------------------------------------------------------
import itertools  # noqa: E402
import os  # noqa: E402

from oracle import grid_sampler as ogs  # noqa: E402
from tests.helpers import GOLDEN, make_grid_sampler_inputs  # noqa: E402

_GS_MODES = list(itertools.product(["bilinear", "nearest", "bicubic"], ["zeros", "border", "reflection"], [False, True]))
_IM = {"bilinear": 0, "nearest": 1, "bicubic": 2}
_PM = {"zeros": 0, "border": 1, "reflection": 2}


@pytest.mark.parametrize("interp,pad,align", _GS_MODES)
def test_grid_sampler_c_oracle_matches_reference_python(interp, pad, align):
    z = np.load(os.path.join(GOLDEN, "grid_sampler_ref.npz"))
    inp, grid = make_grid_sampler_inputs(2, 6, 11, 13, 17, 19, seed=3)
    want = z[f"{interp}_{pad}_{int(align)}"]
    got = ogs.grid_sample_2d(inp.numpy(), grid.numpy(), _IM[interp], _PM[pad], align)
    d = np.abs(got - want)
    if interp == "nearest":
        # kernel: ::round (half away from zero) on an index un-normalised from [-10,10]; binding: nearbyint on /10 then
        # [-1,1]. They may pick different pixels exactly at half-way points (reference test delta 0.1, :146-147).
        assert (d > 1e-6).mean() < 0.01
    else:
        assert d.max() < 3e-5, d.max()


def test_grid_sampler_torch_port_matches_reference_python():
    z = np.load(os.path.join(GOLDEN, "grid_sampler_ref.npz"))
    inp, grid = make_grid_sampler_inputs(2, 6, 11, 13, 17, 19, seed=3)
    for interp, pad, align in _GS_MODES:
        got = ogs.grid_sampler_torch_port(inp, grid, _IM[interp], _PM[pad], align).numpy()
        assert np.array_equal(got, z[f"{interp}_{pad}_{int(align)}"])


# ---------------------------------------------------------------------------------------------------------------
# RotateTRT: golden vectors = the reference's own Python binding (tests/golden/make_golden_rotate.py)
# ---------------------------------------------------------------------------------------------------------------
from oracle import rotate as orot  # noqa: E402
from tests.helpers import ROTATE_CASES, make_rotate_inputs  # noqa: E402


@pytest.mark.parametrize("case", list(ROTATE_CASES))
@pytest.mark.parametrize("interp", ["bilinear", "nearest"])
def test_rotate_c_oracle_matches_reference_python(case, interp):
    z = np.load(os.path.join(GOLDEN, "rotate_ref.npz"))
    img, angle, center = make_rotate_inputs(case)
    want = z[f"{case}_{interp}"]
    got = orot.rotate(img.numpy(), float(angle[0]), center.numpy(), 0 if interp == "bilinear" else 1)
    d = np.abs(got - want)
    if interp == "nearest":
        # kernel: ::round (half away from zero); binding: aten nearbyint (half to even) on a grid built through bmm.
        # They pick different pixels only where the source index is within rounding of x.5 (right_angle / zero cases
        # sit exactly there). The reference's own test absorbs this in a mean-abs delta (test_rotate.py:98-105).
        frac = (d > 1e-6).mean()
        assert frac < (0.08 if case in ("right_angle", "zero") else 0.005), frac
    else:
        assert d.max() < 5e-5, d.max()


def test_rotate_torch_port_matches_reference_python():
    z = np.load(os.path.join(GOLDEN, "rotate_ref.npz"))
    for case in ROTATE_CASES:
        img, angle, center = make_rotate_inputs(case)
        for interp, name in ((0, "bilinear"), (1, "nearest")):
            got = orot.rotate_torch_port(img, angle[0], center, interp).numpy()
            assert np.array_equal(got, z[f"{case}_{name}"]), (case, name)


def test_rotate_oracle_properties():
    """Size-independent properties: zero angle about the image centre is the identity; a 90-degree turn of a square
    image about its centre is an exact transpose-flip (both modes); rotation is linear in the image."""
    rng = np.random.default_rng(5)
    img = rng.standard_normal((3, 24, 24)).astype(np.float32)
    assert np.array_equal(orot.rotate(img, 0.0, (12.0, 12.0), 1), img)
    assert np.abs(orot.rotate(img, 0.0, (12.0, 12.0), 0) - img).max() < 1e-5  # (gx + 1) * W / 2 rounds in fp32
    a, b = img, rng.standard_normal(img.shape).astype(np.float32)
    lin = orot.rotate(a + 2 * b, 17.0, (11.0, 13.0), 0)
    assert np.abs(lin - (orot.rotate(a, 17.0, (11.0, 13.0), 0) + 2 * orot.rotate(b, 17.0, (11.0, 13.0), 0))).max() < 1e-5
    r90 = orot.rotate(img, 90.0, (12.0, 12.0), 0)
    want = np.rot90(img, k=1, axes=(1, 2))  # counter-clockwise: out[h, w] = img[w, W-1-h]
    assert np.abs(r90 - want).max() < 1e-4


def test_rotate_t2int8_rounds_half_away_from_zero():
    got = orot.t2int8(np.array([0.5, -0.5, 1.5, -1.5, 126.5, 127.4, 300.0, -128.6, -300.0, 0.49], np.float32))
    assert got.tolist() == [1, -1, 2, -2, 127, 127, 127, -128, -128, 0]


# ---------------------------------------------------------------------------------------------------------------
# BEV point sampling (encoder prologue): golden = the reference's own two methods (make_golden_point_sampling.py)
# ---------------------------------------------------------------------------------------------------------------
from oracle import point_sampling as ops  # noqa: E402
from tests.helpers import POINT_SAMPLING_CASES, make_point_sampling_inputs  # noqa: E402


def assert_point_sampling_close(cam, mask, want_cam, want_mask, rtol=2e-5):
    """reference_points_cam: 2e-5 relative (coordinates of points behind a camera are ~1e8); bev_mask: identical except
    for queries with a pillar point within rounding of an image border or of depth eps."""
    assert cam.shape == want_cam.shape and mask.shape == want_mask.shape
    rel = np.abs(cam - want_cam) / np.maximum(1.0, np.abs(want_cam))
    assert rel.max() < rtol, rel.max()
    assert (mask != want_mask).any(0).mean() < 2e-3


@pytest.mark.parametrize("case", list(POINT_SAMPLING_CASES))
def test_point_sampling_oracle_matches_reference_python(case):
    z = np.load(os.path.join(GOLDEN, "point_sampling_ref.npz"))
    H, W, D, img_hw, pc_range = POINT_SAMPLING_CASES[case]
    ref3d = ops.get_reference_points_3d(H, W, pc_range[5] - pc_range[2], D)
    # linspace / divide restated bit for bit. The reference hard-codes `.view(1, 4, -1, 3)` (encoder.py:193), so for
    # D != 4 its tensor has the same memory under a different shape; point_sampling_trt re-views it by D (:208-210).
    assert np.array_equal(ref3d.reshape(-1), z[f"{case}_ref3d"].reshape(-1))
    cam, mask = ops.point_sampling(ref3d, pc_range, make_point_sampling_inputs(case).numpy(), img_hw)
    assert_point_sampling_close(cam, mask, z[f"{case}_cam"], z[f"{case}_mask"])
    # the fixture is not degenerate: every camera sees some queries, none sees all, weights sum to 1 where seen
    m = z[f"{case}_mask"][..., 0]
    assert ((m > 0).mean(1) > 0.02).all() and ((m > 0).mean(1) < 0.6).all()
    s = m.sum(0)
    assert np.all((np.abs(s - 1) < 1e-6) | (s == 0))


def test_point_sampling_workload_generator_matches_reference_python():
    """bevformer_tensorrt_b200.workloads.bev_reference_points_cam (the bench's distribution-G input generator) is the
    same computation."""
    from bevformer_tensorrt_b200.workloads import bev_reference_points_cam

    z = np.load(os.path.join(GOLDEN, "point_sampling_ref.npz"))
    H, W, D, img_hw, pc_range = POINT_SAMPLING_CASES["ring_small"]
    uv, mask = bev_reference_points_cam((H, W), make_point_sampling_inputs("ring_small"), img_hw=img_hw,
                                        Z=pc_range[5] - pc_range[2], pillars=D, pc_range=pc_range)  # fmt: skip
    # einsum sums the 4x4 product in another order: where the depth nearly cancels the quotient moves by ~1e-4 relative
    assert_point_sampling_close(uv.numpy()[:, None], mask.numpy(), z["ring_small_cam"], z["ring_small_mask"], rtol=3e-4)


# ---------------------------------------------------------------------------------------------------------------
# DCNv2: mmcv (the reference binding's forward) is absent, so the oracle is pinned to torchvision's independent
# implementation of the same definition (same offset channel order: 2*(i*kw+j) = dh, +1 = dw).
# ---------------------------------------------------------------------------------------------------------------
from oracle import dcn as odcn  # noqa: E402
from tests.helpers import make_dcn_inputs  # noqa: E402


@pytest.mark.parametrize("case", ["k3_s1_p1_g2_dg2", "k3_s2_p1_g1_dg1", "k3_s1_p2_d2_g1_dg4", "k1_s1_p0_g1_dg1",
                                  "k3x5_s1_p1_g1_dg1"])  # fmt: skip
def test_dcn_oracle_matches_torchvision(case):
    import torchvision

    x, off, mask, w, b, kw = make_dcn_inputs(case)
    want = torchvision.ops.deform_conv2d(x, off, w, b, stride=kw["stride"], padding=kw["padding"],
                                         dilation=kw["dilation"], mask=mask).numpy()  # fmt: skip
    got = odcn.modulated_deformable_conv2d(x.numpy(), off.numpy(), mask.numpy(), w.numpy(), b.numpy(), **kw)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 1e-4 * max(1.0, np.abs(want).max())
