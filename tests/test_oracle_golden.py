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
