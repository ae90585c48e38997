"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/b200_bev_ops.h declares
(no compute calls without a GPU), argument validation, the registry, workload generators, and that the product package
never imports oracle/."""
import ast
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import bevformer_tensorrt_b200 as bt
from bevformer_tensorrt_b200 import _lib
from bevformer_tensorrt_b200.workloads import (CONFIGS, bev_reference_points_cam, camera_ring_lidar2img,
                                               make_msda_inputs, quantize_per_tensor)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b200_bev_ops.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200_bev_ops.h but not exported"
    # and every symbol the Python binding uses is declared in the header
    assert set(_lib.SIGNATURES) <= set(syms), set(_lib.SIGNATURES) - set(syms)


def test_version_status_and_argument_validation_without_gpu():
    lib = _lib.load()
    assert b"sm_100a" in lib.b200_bev_ops_version()
    assert lib.b200_status_string(0) == b"ok" and lib.b200_status_string(2) == b"bad parameter"
    # null pointers / bad dims are rejected before any CUDA call
    assert lib.b200_msda_f32(None, None, None, None, None, 1, 1, 1, 32, 1, 1, 4, 1, None, None) == 2
    assert lib.b200_msda_i8(None, 1.0, None, None, 0, None, 1.0, None, 1.0, 1, 1, 1, 30, 1, 1, 4, 1, None, 1.0, None) == 1
    assert lib.b200_msda_i8(None, 1.0, None, None, 0, None, 1.0, None, 1.0, 1, 1, 1, 32, 1, 1, 4, 1, None, 0.0, None) == 2
    dims = (ctypes.c_int * 4)(1, 1, 1, 1)
    assert lib.b200_grid_sample_f32(None, None, None, dims, dims, dims, 4, 0, 0, 0, None) == 2
    assert lib.b200_msda_enqueue(None, None, None, None, None, None, 0) == 2
    assert lib.b200_msda_supports_format(0, None, 5, 1) == 0


def test_supports_format_mirror():
    lib = _lib.load()

    def desc(shape, t, fmt=0):
        d = _lib.TensorDesc()
        d.dims.nbDims = len(shape)
        for i, s in enumerate(shape):
            d.dims.d[i] = s
        d.type, d.format, d.scale = t, fmt, 1.0
        return d

    def io(vt, rt, ch=32, pts=32):
        return (_lib.TensorDesc * 6)(desc((6, 100, 8, ch), vt), desc((4, 2), 3), desc((6, 10, 1, 8), rt),
                                     desc((6, 10, 8, 2 * pts), vt), desc((6, 10, 8, pts), vt), desc((6, 10, 8, ch), vt))

    for pos in range(6):
        assert lib.b200_msda_supports_format(pos, io(0, 0), 5, 1) == 1
        assert lib.b200_msda_supports_format(pos, io(1, 1), 5, 1) == 1
        assert lib.b200_msda_supports_format(pos, io(2, 1), 5, 1) == 1  # int8 value, fp16 ref points
        assert lib.b200_msda_supports_format(pos, io(2, 0), 5, 1) == 1  # int8 value, fp32 ref points
    assert lib.b200_msda_supports_format(2, io(1, 0), 5, 1) == 0  # fp16 value needs fp16 ref points
    assert lib.b200_msda_supports_format(0, io(2, 1, ch=30), 5, 1) == 0  # int8 needs channels % 4 == 0
    assert lib.b200_msda_supports_format(0, io(2, 1, pts=12), 5, 1) == 0  # 12/4 levels = 3 points: % 4 != 0
    assert lib.b200_msda_supports_format(1, io(0, 0), 5, 1) == 1 and lib.b200_msda_supports_format(6, io(0, 0), 5, 1) == 0


def test_registry_and_names():
    for name in ("multi_scale_deformable_attn", "multi_scale_deformable_attn2", "grid_sampler", "grid_sampler2"):
        assert bt.TRT_FUNCTIONS.get(name) is getattr(bt, name)
    with pytest.raises(KeyError):
        bt.TRT_FUNCTIONS.register_module(module=bt.grid_sampler)
    with pytest.raises(TypeError):
        bt.TRT_FUNCTIONS.register_module(force=1, module=bt.grid_sampler)

    @bt.TRT_FUNCTIONS.register_module(name="_tmp_fn")
    def f():
        return 1

    assert "_tmp_fn" in bt.TRT_FUNCTIONS and bt.TRT_FUNCTIONS.get("_tmp_fn") is f


def test_cpu_tensors_are_rejected_loudly():
    cfg = CONFIGS["cpu_plumbing"]
    ins = make_msda_inputs(cfg, "U", 0)
    with pytest.raises(AssertionError):
        bt.multi_scale_deformable_attn(*ins)
    with pytest.raises(RuntimeError):
        bt.grid_sampler(torch.zeros(1, 1, 2, 2), torch.zeros(1, 2, 2, 2), "bilinear", "zeros", False)


def test_onnx_symbolic_names_are_the_reference_plugin_names():
    import importlib

    gs = importlib.import_module("bevformer_tensorrt_b200.functions.grid_sampler")
    m = importlib.import_module("bevformer_tensorrt_b200.functions.multi_scale_deformable_attn")

    class G:
        def op(self, name, *a, **k):
            return name, a, k

    assert m._MultiScaleDeformableAttnFunction.symbolic(G(), 1, 2, 3, 4, 5)[0] == "MultiScaleDeformableAttnTRT"
    assert m._MultiScaleDeformableAttnFunction2.symbolic(G(), 1, 2, 3, 4, 5)[0] == "MultiScaleDeformableAttnTRT2"
    n, a, k = gs._GridSampler2D.symbolic(G(), 1, 2, 0, 1, True)
    assert n == "GridSampler2DTRT" and k == {"interpolation_mode_i": 0, "padding_mode_i": 1, "align_corners_i": True}
    assert gs._GridSampler3D2.symbolic(G(), 1, 2, 0, 0, False)[0] == "GridSampler3DTRT2"
    rot = importlib.import_module("bevformer_tensorrt_b200.functions.rotate")
    n, a, k = rot._Rotate.symbolic(G(), 1, 2, 3, 1)  # rotate.py:9-10: g.op("RotateTRT", img, angle, center, interpolation_i=…)
    assert n == "RotateTRT" and a == (1, 2, 3) and k == {"interpolation_i": 1}
    assert rot._Rotate2.symbolic(G(), 1, 2, 3, 0)[0] == "RotateTRT2"
    dcn = importlib.import_module("bevformer_tensorrt_b200.functions.modulated_deformable_conv2d")
    names = {c.symbolic(G(), 1, 2, 3, 4, 5, 1, 1, 1, 1, 1)[0] for c in vars(dcn).values()
             if isinstance(c, type) and hasattr(c, "symbolic") and c.__module__ == dcn.__name__}
    assert names == {"ModulatedDeformableConv2dTRT", "ModulatedDeformableConv2dTRT2"}


def test_workload_generators():
    cfg = CONFIGS["base_sca"]
    assert cfg.spatial_size == 30825 and cfg.algorithmic_bytes(2) == 590054432
    assert cfg.algorithmic_bytes(1, 2) == 296947232
    assert CONFIGS["tiny_sca"].algorithmic_bytes(2) == 14832000 + 8
    uv, mask = bev_reference_points_cam((50, 50), camera_ring_lidar2img(6))
    assert uv.shape == (6, 2500, 4, 2) and mask.shape == (6, 2500, 1)
    seen = (mask > 0).float().sum(0)
    assert 0.9 < seen.mean() < 2.1  # about one or two cameras see a BEV pillar
    assert torch.allclose(mask.sum(0)[seen > 0], torch.ones_like(mask.sum(0)[seen > 0]))
    a = make_msda_inputs(CONFIGS["tiny_sca"], "G", 5, torch.float16)
    b = make_msda_inputs(CONFIGS["tiny_sca"], "G", 5, torch.float16)
    assert all(torch.equal(x, y) for x, y in zip(a, b)) and torch.isfinite(a[2].float()).all()
    q, s = quantize_per_tensor(torch.tensor([-2.0, 0.5, 1.0]))
    assert q.tolist() == [-127, 32, 64] and abs(s - 2 / 127) < 1e-9


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bevformer_tensorrt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                tree = ast.parse(open(os.path.join(dirpath, f)).read())
                for node in ast.walk(tree):
                    names = []
                    if isinstance(node, ast.Import):
                        names = [a.name for a in node.names]
                    elif isinstance(node, ast.ImportFrom) and node.module:
                        names = [node.module]
                    assert not any(n == "oracle" or n.startswith("oracle.") for n in names), (f, names)
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                for line in open(os.path.join(dirpath, f)):
                    if line.lstrip().startswith("#include") or "dlopen" in line:
                        assert "oracle" not in line, (f, line)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU/PyTorch fallback"):
        _lib.load()


def test_tensorrt_plugin_shell_compiles_against_the_mock_api():
    """TensorRT is not installed: the plugin shells (csrc/trt_plugin) are at least compiled against a declaration-only
    mock of the TensorRT plugin API, which checks override signatures, the PluginTensorDesc layout assertion and the C-ABI
    calls. With the real NvInfer.h the same file builds the registrable plugins."""
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        pytest.skip("no g++")
    src = os.path.join(ROOT, "bevformer_tensorrt_b200", "csrc", "trt_plugin", "b200_trt_plugins.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "mock_trt"),
                        "-I", os.path.join(ROOT, "include"), src], capture_output=True, text=True)  # fmt: skip
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    for name in ("MultiScaleDeformableAttnTRT", "MultiScaleDeformableAttnTRT2", "GridSampler2DTRT", "GridSampler2DTRT2",
                 "GridSampler3DTRT", "GridSampler3DTRT2", "ModulatedDeformableConv2dTRT", "ModulatedDeformableConv2dTRT2",
                 "RotateTRT", "RotateTRT2"):  # fmt: skip
        assert f'"{name}"' in text


def test_bench_stdout_redirect_keeps_native_output_off_stdout():
    """bench.py at N>1 must print exactly one JSON line on stdout although NCCL writes a banner to fd 1."""
    import subprocess
    import sys

    code = (
        "import importlib.util, os, sys\n"
        f"spec = importlib.util.spec_from_file_location('bench', {os.path.join(ROOT, 'bench.py')!r})\n"
        "bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)\n"
        "with bench.stdout_to_stderr():\n"
        "    os.write(1, b'native banner\\n'); print('python inside')\n"
        "print('{\"only\": 1}', flush=True)\n"
    )
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == '{"only": 1}'
    assert "native banner" in r.stderr and "python inside" in r.stderr


def test_reference_points_3d_matches_reference_golden_on_cpu():
    """get_reference_points_3d is plain tensor construction: on the CPU it reproduces the reference's tensor bit for
    bit (the GPU-side consumers are covered by tests/test_point_sampling_gpu.py)."""
    from tests.helpers import GOLDEN, POINT_SAMPLING_CASES

    z = np.load(os.path.join(GOLDEN, "point_sampling_ref.npz"))
    for case, (H, W, D, _, pc_range) in POINT_SAMPLING_CASES.items():
        got = bt.get_reference_points_3d(H, W, pc_range[5] - pc_range[2], D, device="cpu")
        assert got.shape == (1, D, H * W, 3)
        assert np.array_equal(got.numpy().reshape(-1), z[f"{case}_ref3d"].reshape(-1)), case


def test_rotate_recognises_channels_last_views_only():
    from bevformer_tensorrt_b200.functions.rotate import _is_hwc_view

    bev = torch.zeros(50 * 46, 1, 64)
    assert _is_hwc_view(bev.view(50, 46, 64).permute(2, 0, 1))           # BEVFormer's call-site expression
    assert not _is_hwc_view(bev.view(50, 46, 64).permute(2, 0, 1).contiguous())
    assert not _is_hwc_view(torch.zeros(64, 50, 46))
    assert not _is_hwc_view(torch.zeros(50, 46, 64).permute(2, 1, 0))     # transposed spatial dims: not [H, W, C] memory
    assert not _is_hwc_view(torch.zeros(50, 46, 128)[:, :, ::2].permute(2, 0, 1))  # strided channels


def test_cpu_tensors_are_refused_by_the_new_ops():
    for call in (lambda: bt.rotate(torch.zeros(4, 5, 5), torch.tensor(1.0), torch.tensor([2.0, 2.0])),
                 lambda: bt.rotate_hwc(torch.zeros(5, 5, 4), torch.tensor(1.0), torch.tensor([2.0, 2.0])),
                 lambda: bt.point_sampling_trt(torch.zeros(1, 4, 9, 3), (-1, -1, -1, 1, 1, 1), torch.eye(4)[None], (8, 8)),
                 lambda: bt.bev_point_sampling(3, 3, (-1, -1, -1, 1, 1, 1), torch.eye(4)[None], (8, 8)),
                 lambda: bt.multi_scale_deformable_attn_sca_shared(torch.zeros(1, 4, 1, 32), torch.tensor([[2, 2]]),
                                                                   torch.zeros(1, 3, 1, 2), torch.zeros(3, 1, 8),
                                                                   torch.zeros(3, 1, 4), torch.zeros(1, 3))):  # fmt: skip
        with pytest.raises(RuntimeError):
            call()


def test_tensorrt_plugin_shell_behaviour_through_the_mock_api(tmp_path):
    """Builds tests/mock_trt/shell_harness.cpp (the shells + the mock plugin API + libb200_bev_ops.so) and runs it: creator
    names / fields, create -> serialise -> deserialise -> clone round trips, output dimensions, the format negotiation
    tables of all five plugins, workspace sizes, and enqueue() forwarding (bad arguments -> status). No GPU needed."""
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        pytest.skip("no g++")
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    cuda_lib = "/usr/local/cuda/lib64"
    exe = str(tmp_path / "shell_harness")
    cmd = ["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "tests", "mock_trt"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "mock_trt", "shell_harness.cpp"), "-L", lib_dir, "-lb200_bev_ops", "-L", cuda_lib,
           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{cuda_lib}", "-o", exe]  # fmt: skip
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    assert r.stdout.startswith("OK ") and int(r.stdout.split()[1]) > 100


def test_c_header_is_plain_c_and_the_example_host_runs(tmp_path):
    """include/b200_bev_ops.h must be usable from C (the drop-in boundary is a C ABI): examples/minimal_host.c is
    compiled as strict C99 and run; it exercises version, format negotiation and argument validation without a GPU."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "minimal_host")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "minimal_host.c"), "-L", lib_dir, "-lb200_bev_ops", "-L", "/usr/local/cuda/lib64",
           f"-Wl,-rpath,{lib_dir}", "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe]  # fmt: skip
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "accepted for all six tensors" in r.stdout and "status 2" in r.stdout


# ---------------------------------------------------------------------------------------------------------------
# launch-shape switches and the autotuner (host logic: no kernel is launched here)
# ---------------------------------------------------------------------------------------------------------------
def test_launch_shape_setters_round_trip_without_gpu():
    lib = _lib.load()
    prev_units, prev_var = bt.get_msda_batch_units(), bt.get_msda_gather_variant()
    try:
        assert bt.set_msda_batch_units(4, True) == prev_units
        assert bt.get_msda_batch_units() == (4, True)
        assert lib.b200_msda_set_batch_units(3, 0) == (4 | 0x100)  # 3 is not a launch shape: query only
        assert bt.get_msda_batch_units() == (4, True)
        assert bt.set_msda_batch_units(1) == (4, True) and bt.get_msda_batch_units() == (1, False)
        with pytest.raises(ValueError):
            bt.set_msda_batch_units(8)
        assert bt.set_msda_gather_variant(2) == prev_var and bt.get_msda_gather_variant() == 2
        assert lib.b200_msda_set_gather_variant(7) == 2 and bt.get_msda_gather_variant() == 2  # out of range: query only
        with pytest.raises(ValueError):
            bt.set_msda_gather_variant(4)
        for name, (units, strided, variant) in bt.MSDA_LAUNCH_SHAPES.items():
            bt.set_msda_launch_shape(name)
            assert bt.get_msda_batch_units() == (units, strided) and bt.get_msda_gather_variant() == variant
    finally:
        bt.set_msda_batch_units(*prev_units)
        bt.set_msda_gather_variant(prev_var)


def test_launch_shape_environment_variables(tmp_path):
    """B200_MSDA_BATCH / B200_MSDA_VARIANT are read once, by the first query of a fresh process."""
    import subprocess
    import sys

    code = ("import bevformer_tensorrt_b200 as bt; "
            "print(bt.get_msda_batch_units(), bt.get_msda_gather_variant())")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def run(env):
        e = dict(os.environ, PYTHONPATH=root, **env)
        return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=e, cwd=root).stdout.strip()

    assert run({}) == "(1, False) 0"
    assert run({"B200_MSDA_BATCH": "4s", "B200_MSDA_VARIANT": "1"}) == "(4, True) 1"
    assert run({"B200_MSDA_BATCH": "2", "B200_MSDA_VARIANT": "2"}) == "(2, False) 2"
    assert run({"B200_MSDA_BATCH": "3", "B200_MSDA_VARIANT": "9"}) == "(1, False) 0"  # not launch shapes: ignored
    assert run({"B200_MSDA_BATCH": "1s"}) == "(1, False) 0"


class _FakeCudaTensor:
    """Stands in for a CUDA tensor in the autotuner's argument checks (no kernel runs in this test)."""

    is_cuda, dtype, device = True, torch.float16, "cpu"


def _run_fake_autotune(monkeypatch, speeds, wrong=()):
    """Drives autotune_msda with a fake op: `speeds` = simulated ms per launch shape name, `wrong` = shapes whose output
    differs from the default's."""
    import sys

    mod = sys.modules["bevformer_tensorrt_b200.functions.multi_scale_deformable_attn"]  # the module, not the op of that name
    clock = {"t": 0.0}
    shapes_by_setting = {v: k for k, v in bt.MSDA_LAUNCH_SHAPES.items()}

    def current():
        units, strided = bt.get_msda_batch_units()
        return shapes_by_setting[(units, strided, bt.get_msda_gather_variant())]

    def fake_op(*_a):
        name = current()
        clock["t"] += speeds[name]
        return torch.full((4,), 1.0 if name in wrong else 0.0)

    class FakeEvent:
        def __init__(self, enable_timing=False):
            self.t = None

        def record(self):
            self.t = clock["t"]

        def elapsed_time(self, other):
            return other.t - self.t

    monkeypatch.setattr(mod, "multi_scale_deformable_attn", fake_op)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *_a, **_k: None)
    t = _FakeCudaTensor()
    return mod.autotune_msda(t, t, t, t, t)


def test_autotuner_picks_the_fastest_identical_shape(monkeypatch):
    prev_units, prev_var = bt.get_msda_batch_units(), bt.get_msda_gather_variant()
    try:
        # a clearly faster shape wins and becomes the process-wide setting
        rep = _run_fake_autotune(monkeypatch, {"default": 1.0, "batch2_strided": 0.97, "deep_gather": 0.80,
                                               "deep_gather_explicit": 0.85, "mid_gather": 0.9, "deep_batch2_strided": 0.95,
                                               "deep_batch4_strided": 0.96})
        assert rep["chosen"] == "deep_gather" and rep["rejected"] == []
        assert abs(rep["ms"]["deep_gather"] - 0.80) < 1e-9 and abs(rep["ms"]["default"] - 1.0) < 1e-9
        assert bt.get_msda_batch_units() == (1, False) and bt.get_msda_gather_variant() == 1
        # a faster shape whose bits differ is never taken, whatever its speed
        rep = _run_fake_autotune(monkeypatch, {"default": 1.0, "batch2_strided": 0.9, "deep_gather": 0.1,
                                               "deep_gather_explicit": 0.95, "mid_gather": 0.93, "deep_batch2_strided": 0.97,
                                               "deep_batch4_strided": 0.98}, wrong=("deep_gather",))
        assert rep["chosen"] == "batch2_strided" and rep["rejected"] == ["deep_gather"] and "deep_gather" not in rep["ms"]
        assert bt.get_msda_batch_units() == (2, True) and bt.get_msda_gather_variant() == 0
        # gains inside the noise band (min_gain = 2 %) leave the default in place
        rep = _run_fake_autotune(monkeypatch, {"default": 1.0, "batch2_strided": 0.99, "deep_gather": 0.985,
                                               "deep_gather_explicit": 1.2, "mid_gather": 1.01, "deep_batch2_strided": 1.1,
                                               "deep_batch4_strided": 1.3})
        assert rep["chosen"] == "default"
        assert bt.get_msda_batch_units() == (1, False) and bt.get_msda_gather_variant() == 0
        # CPU tensors are refused before anything is touched
        with pytest.raises(ValueError):
            bt.autotune_msda(*(torch.zeros(1),) * 5)
    finally:
        bt.set_msda_batch_units(*prev_units)
        bt.set_msda_gather_variant(prev_var)


def test_fused_autotuner_logic(monkeypatch):
    """autotune_msda_fused with a fake step: picks the 2-CTA shape only when it is faster AND agrees with the default."""
    import sys

    mod = sys.modules["bevformer_tensorrt_b200.functions.multi_scale_deformable_attn"]
    clock = {"t": 0.0}

    class FakeEvent:
        def __init__(self, enable_timing=False):
            self.t = None

        def record(self):
            self.t = clock["t"]

        def elapsed_time(self, other):
            return other.t - self.t

    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *_a, **_k: None)
    prev = bt.get_msda_gather_variant()
    box = {"out": None}

    def make_run(ms, err):
        def run():
            v = bt.get_msda_gather_variant()
            clock["t"] += ms[v]
            box["out"] = torch.full((8,), 2.0 + (err if v else 0.0))

        return run

    try:
        rep = mod.autotune_msda_fused(make_run({0: 1.0, 1: 0.7}, 1e-6), lambda: box["out"])
        assert rep["chosen"] == "deep_gather" and bt.get_msda_gather_variant() == 1 and rep["rejected"] == []
        rep = mod.autotune_msda_fused(make_run({0: 1.0, 1: 0.7}, 1e-2), lambda: box["out"])  # wrong result: never taken
        assert rep["chosen"] == "default" and rep["rejected"] == ["deep_gather"] and bt.get_msda_gather_variant() == 0
        rep = mod.autotune_msda_fused(make_run({0: 1.0, 1: 0.99}, 0.0), lambda: box["out"])  # inside the noise band
        assert rep["chosen"] == "default" and bt.get_msda_gather_variant() == 0
    finally:
        bt.set_msda_gather_variant(prev)
