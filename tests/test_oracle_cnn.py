"""Pin oracle/cnn.py against golden vectors produced by the reference's own src/models code
(tests/golden/make_golden.py)."""
import json
import os

import pytest
import safetensors.torch as st
import torch

from oracle import cnn as OC


@pytest.fixture(scope="module")
def golden(golden_dir):
    return st.load_file(os.path.join(golden_dir, "cnn_golden.safetensors"))


def test_mlacfd_matches_reference_output(golden, golden_dir):
    cfg = json.load(open(os.path.join(golden_dir, "MLACFD.json")))
    sd = st.load_file(os.path.join(golden_dir, "mlacfd_weights.safetensors"))
    # weights contract: 14 fp64 tensors, keys models.0.layers.{i}.weight|bias
    assert len(sd) == 14 and all(t.dtype == torch.float64 for t in sd.values())
    assert sd["models.0.layers.0.weight"].shape == (64, 2, 3, 3)
    assert sd["models.0.layers.6.weight"].shape == (2, 64, 3, 3)
    assert sum(t.numel() for t in sd.values()) == 187010
    y = OC.model_forward(golden["mlacfd_x"], cfg, sd)
    assert (y - golden["mlacfd_y"]).abs().max().item() <= 1e-15
    ys = OC.model_forward(golden["mlacfd_x"] / 7.0, cfg, sd)
    assert (ys - golden["mlacfd_y_scaled"]).abs().max().item() <= 1e-15


def test_bn_variant_matches_reference_output(golden, golden_dir):
    cfg = json.load(open(os.path.join(golden_dir, "cnn1.json")))
    sd = st.load_file(os.path.join(golden_dir, "cnn1_bn_weights.safetensors"))
    assert "models.0.layers.0.1.running_mean" in sd
    y = OC.model_forward(golden["cnn1_x"], cfg, sd)
    assert (y - golden["cnn1_y"]).abs().max().item() <= 1e-13


def test_bf16_emulation_is_close_but_not_identical(golden, golden_dir):
    cfg = json.load(open(os.path.join(golden_dir, "MLACFD.json")))
    sd = st.load_file(os.path.join(golden_dir, "mlacfd_weights.safetensors"))
    y = OC.model_forward(golden["mlacfd_x"], cfg, sd)
    yb = OC.model_forward(golden["mlacfd_x"], cfg, sd, round_bf16=True)
    rel = ((yb - y).norm() / y.norm()).item()
    assert 1e-5 < rel < 1e-2


def test_correction_add_semantics(golden, golden_dir):
    cfg = json.load(open(os.path.join(golden_dir, "MLACFD.json")))
    sd = st.load_file(os.path.join(golden_dir, "mlacfd_weights.safetensors"))
    x = golden["mlacfd_x"]
    u2, v2 = OC.correction_add(x[:, 0], x[:, 1], cfg, sd)
    assert torch.equal(u2, x[:, 0] + golden["mlacfd_y"][:, 0])
    assert torch.equal(v2, x[:, 1] + golden["mlacfd_y"][:, 1])


def test_config_helpers_follow_reference_semantics():
    assert OC.param_to_list(3, 4) == [3, 3, 3, 3]
    assert OC.param_to_list([1, 2], 2) == [1, 2]
    with pytest.raises(ValueError):
        OC.param_to_list([1, 2], 3)
    with pytest.raises(TypeError):
        OC.param_to_list("3", 3)
    assert OC.structure_to_list({"in_channels": 2, "hidden_channels": [4, 5], "out_channels": 2}) == [2, 4, 5, 2]
