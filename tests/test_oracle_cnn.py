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


def test_resnet_oracle_matches_reference_golden(golden_dir):
    """ResNet basic-block chain (reference configs/fullmodels/basicResnet1.json, src/models/ResNet.py:26-50): oracle vs
    the reference's own forward (tests/golden/make_golden_resnet.py), fp64."""
    import json
    import os
    import safetensors.torch as st
    from oracle import cnn as OC
    cfg = json.load(open(os.path.join(golden_dir, "basicResnet1.json")))
    sd = st.load_file(os.path.join(golden_dir, "resnet_weights.safetensors"))
    g = st.load_file(os.path.join(golden_dir, "resnet_golden.safetensors"))
    y = OC.model_forward(g["x"], cfg, sd)
    assert (y - g["y"]).abs().max().item() <= 1e-14
    assert (g["y"] >= 0).all()                       # every block ends with act(x + y)
    yb = OC.model_forward(g["x"], cfg, sd, round_bf16=True)
    assert 1e-4 < ((yb - g["y"]).norm() / g["y"].norm()).item() < 3e-2       # the stated bf16 storage error


def test_resnet_module_tree_loads_reference_checkpoint(golden_dir, built_lib):
    """models.buildModel(basicResnet1.json) has the reference's parameter names and folds BN like the oracle."""
    import json
    import os
    import safetensors.torch as st
    import torch
    from ml_accelerated_simulation_b200 import models
    cfg = json.load(open(os.path.join(golden_dir, "basicResnet1.json")))
    sd = st.load_file(os.path.join(golden_dir, "resnet_weights.safetensors"))
    model = models.buildModel(cfg)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    layers, shortcuts, final_relu = model.flat_layers()
    assert len(layers) == 12 and final_relu
    assert sorted(shortcuts) == [1, 3, 5, 7, 9, 11]
    assert shortcuts[1][0] == -1 and shortcuts[1][1].shape == (64, 2, 1, 1)          # 1x1 conv of the network input
    assert shortcuts[3] == (1, None, None)                                             # identity
    assert shortcuts[11][0] == 9 and shortcuts[11][1].shape == (2, 64, 1, 1)         # 1x1 conv of the block input
    w, b = layers[0]
    bn = "models.0.layers.0.1."
    scale = sd[bn + "weight"].double() / torch.sqrt(sd[bn + "running_var"].double() + 1e-5)
    assert torch.allclose(w, sd["models.0.layers.0.0.weight"].double() * scale[:, None, None, None], atol=1e-15)


@pytest.mark.parametrize("name,tag", [("basicBottleneck1", "bottleneck"), ("basicResNeXt1", "resnext")])
def test_bottleneck_and_resnext_oracle_match_reference_golden(golden_dir, name, tag):
    """bottleneck (1x1 / 3x3 / 1x1, src/models/ResNet.py:26-50) and ResNeXt (groups = 2, conv1 with its
    own BatchNorm, :54-92) chains: oracle vs the reference's own fp64 forward (tests/golden/make_golden_resnet.py)."""
    import json
    import os
    import safetensors.torch as st
    from oracle import cnn as OC
    cfg = json.load(open(os.path.join(golden_dir, f"{name}.json")))
    sd = st.load_file(os.path.join(golden_dir, f"{tag}_weights.safetensors"))
    g = st.load_file(os.path.join(golden_dir, f"{tag}_golden.safetensors"))
    y = OC.model_forward(g["x"], cfg, sd)
    assert (y - g["y"]).abs().max().item() <= 1e-14
    yb = OC.model_forward(g["x"], cfg, sd, round_bf16=True)
    assert 1e-4 < ((yb - g["y"]).norm() / g["y"].norm()).item() < 3e-2


@pytest.mark.parametrize("name,tag", [("basicBottleneck1", "bottleneck"), ("basicResNeXt1", "resnext")])
def test_bottleneck_and_resnext_module_trees(golden_dir, built_lib, name, tag):
    """models.buildModel has the reference's parameter names, the reference's initial weights under the same seed
    (construction order), and its dense 3x3 embedding of the 1x1 / grouped layers is the same linear map."""
    import json
    import os
    import safetensors.torch as st
    import torch
    import torch.nn.functional as F
    from ml_accelerated_simulation_b200 import models
    cfg = json.load(open(os.path.join(golden_dir, f"{name}.json")))
    sd = st.load_file(os.path.join(golden_dir, f"{tag}_weights.safetensors"))
    g = st.load_file(os.path.join(golden_dir, f"{tag}_golden.safetensors"))
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(2025)                           # the seed of make_golden_resnet.py
        fresh = models.buildModel(cfg)
    finally:
        torch.set_default_dtype(prev)
    for k, t in fresh.state_dict().items():
        if k.endswith(".0.weight") or k.endswith(".0.bias") or k.endswith("conv1.weight"):     # convolutions: untouched by the script
            assert torch.equal(t.float(), sd[k]), k
    model = models.buildModel(cfg)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    layers, shortcuts, final_relu = model.flat_layers()
    assert len(layers) == 18 and final_relu and sorted(shortcuts) == [2, 5, 8, 11, 14, 17]
    assert all(tuple(w.shape[2:]) == (3, 3) for w, _ in layers)
    x, acts = g["x"], {}
    acts[-1] = x
    for i, (w, b) in enumerate(layers):
        y = F.conv2d(acts[i - 1], w, b, padding=1)
        if i in shortcuts:
            src, w1, b1 = shortcuts[i]
            y = y + (acts[src] if w1 is None else F.conv2d(acts[src], w1.reshape(w1.shape[0], -1, 1, 1), b1))
        acts[i] = torch.relu(y)
    assert (acts[17] - g["y"]).abs().max().item() <= 1e-13


def test_128_channel_net_is_refused_not_emulated(built_lib):
    import pytest as _pt
    from ml_accelerated_simulation_b200 import models
    with _pt.raises(NotImplementedError):
        models.buildModel([{"name": "CHANNELAWARE", "structures": {"vector_components": [2, 128, 128, 128],
                                                                     "combined_cnn": [128, 128, 128, 64, 2]}}])
