"""CPU oracle for the learned-correction CNN half of the rollout hot path.  TEST INFRASTRUCTURE ONLY.

Restates the forward pass of the reference's ``CNN`` block as built by ``buildModel``:
  * ``src/models/CNN.py:24-40``  -- stack of ``Conv2d(k, s, p, groups)`` (optionally followed by
    ``BatchNorm2d``), activation + dropout after every layer but the last;
  * ``src/models/tools.py:1-34,37-61`` -- ``paramToList`` / ``structureLoader`` / ``getAct``;
  * ``src/models/modelbuilder.py:11-31`` -- sequential application of blocks;
  * weights contract ``src/train.py:50`` (``st.save_model``): keys ``models.{m}.layers.{i}.weight|bias``
    (bn false) or ``models.{m}.layers.{i}.0.*`` (conv) + ``models.{m}.layers.{i}.1.*`` (BatchNorm).

PINNED: ``tests/golden/make_golden.py`` imports the reference's unmodified ``src/models`` in the build
container, runs it in fp64 and commits inputs/outputs/weights under ``tests/golden/``;
``tests/test_oracle_cnn.py`` checks this file against those vectors.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def param_to_list(param, n: int) -> list:
    """src/models/tools.py:1-19."""
    if isinstance(param, list):
        if len(param) != n:
            raise ValueError("Hyperparameter list length must match the number of layers minus one.")
        return param
    if isinstance(param, (int, float)):
        return [param] * n
    raise TypeError("Hyperparameter must be a float or a list of floats.")


def structure_to_list(structure: dict) -> List[int]:
    """src/models/tools.py:22-34: [in] + hidden + [out]."""
    hold = list(structure.values())
    return [hold[0]] + list(hold[1]) + [hold[2]] if len(hold) == 3 else hold


def _act(name: str):
    name = name.lower()
    if name == "relu":
        return F.relu
    if name == "selu":
        return F.selu
    if name == "gelu":
        return F.gelu
    if name == "lrelu":
        return F.leaky_relu
    raise ValueError(f"Activation function [{name}] not supported by the oracle")


def cnn_block_forward(x: torch.Tensor, config: dict, sd: Dict[str, torch.Tensor], prefix: str,
                      round_bf16: bool = False) -> torch.Tensor:
    """Eval-mode forward of one ``CNN`` block (src/models/CNN.py:36-40).

    ``round_bf16``: emulate the B200 path's storage precision -- weights and every inter-layer
    activation (and the network input) are rounded to bf16, accumulation and bias stay in the
    tensor's own dtype (use fp64/fp32 inputs).  The final layer's output is NOT rounded.
    """
    chans = structure_to_list(config["structures"])
    n = len(chans) - 1
    ks = param_to_list(config["kernel_sizes"], n)
    st = param_to_list(config["strides"], n)
    pd = param_to_list(config["paddings"], n)
    gr = param_to_list(config["group"], n)
    act = _act(config["activation_func"])
    bn = bool(config["bn"])

    def q(t):
        return t.to(torch.bfloat16).to(t.dtype) if round_bf16 else t

    x = q(x)
    for i in range(n):
        cp = f"{prefix}layers.{i}.0." if bn else f"{prefix}layers.{i}."
        w = sd[cp + "weight"].to(x.dtype)
        b = sd[cp + "bias"].to(x.dtype)
        if bn:
            bp = f"{prefix}layers.{i}.1."
            g, beta = sd[bp + "weight"].to(x.dtype), sd[bp + "bias"].to(x.dtype)
            mean, var = sd[bp + "running_mean"].to(x.dtype), sd[bp + "running_var"].to(x.dtype)
            scale = g / torch.sqrt(var + 1e-5)                       # BatchNorm2d eval, eps=1e-5
            w = w * scale[:, None, None, None]
            b = (b - mean) * scale + beta
        x = F.conv2d(x, q(w), b, stride=st[i], padding=pd[i], groups=gr[i])
        if i < n - 1:
            x = q(act(x))           # dropout is the identity in eval mode (training=False)
    return x


def resnet_block_forward(x: torch.Tensor, config: dict, sd: Dict[str, torch.Tensor], prefix: str,
                         round_bf16: bool = False, last_block: bool = False, resnext: bool = False) -> torch.Tensor:
    """Eval-mode forward of one ``ResNetBlock`` (src/models/ResNet.py:26-50): every layer is Conv2d + BatchNorm2d
    (:31) with the block's kernel sizes / paddings / groups (basic: 3x3; bottleneck: 1x1, 3x3, 1x1), activation after
    every layer but the last (:44-46), optional 1x1 ``conv1`` on the block input (:33-39, :48-49), output
    ``act(x + y)`` (:50).  ``resnext``: the ``ResNeXtBlock`` variant (:54-92) -- the same layer stack (:76), ``conv1``
    followed by its own BatchNorm (:78-80).

    ``round_bf16`` as in cnn_block_forward: conv weights (3x3 and 1x1) and every stored activation -- including each
    block's output, except the network's final output (``last_block``) -- are rounded to bf16; ``x`` must already be
    rounded by the caller (it is the previous block's stored output or the rounded network input)."""
    chans = structure_to_list(config["structures"])
    n = len(chans) - 1
    ks = param_to_list(config["kernel_sizes"], n)
    st = param_to_list(config["strides"], n)
    pd = param_to_list(config["paddings"], n)
    gr = param_to_list(config["group"], n)
    act = _act(config["activation_func"])

    def q(t):
        return t.to(torch.bfloat16).to(t.dtype) if round_bf16 else t

    y = x
    for i in range(n):
        cp, bp = f"{prefix}layers.{i}.0.", f"{prefix}layers.{i}.1."
        w, b = sd[cp + "weight"].to(x.dtype), sd[cp + "bias"].to(x.dtype)
        g, beta = sd[bp + "weight"].to(x.dtype), sd[bp + "bias"].to(x.dtype)
        mean, var = sd[bp + "running_mean"].to(x.dtype), sd[bp + "running_var"].to(x.dtype)
        scale = g / torch.sqrt(var + 1e-5)
        w = w * scale[:, None, None, None]
        b = (b - mean) * scale + beta
        y = F.conv2d(y, q(w), b, stride=st[i], padding=pd[i], groups=gr[i])
        if i < n - 1:
            y = q(act(y))
    if config.get("1x1_conv", False):
        if resnext:
            # ResNeXtBlock (src/models/ResNet.py:78-80): conv1 = Sequential(Conv2d 1x1, BatchNorm2d), folded like the layers
            w1, b1 = sd[prefix + "conv1.0.weight"].to(x.dtype), sd[prefix + "conv1.0.bias"].to(x.dtype)
            g, beta = sd[prefix + "conv1.1.weight"].to(x.dtype), sd[prefix + "conv1.1.bias"].to(x.dtype)
            mean, var = sd[prefix + "conv1.1.running_mean"].to(x.dtype), sd[prefix + "conv1.1.running_var"].to(x.dtype)
            scale = g / torch.sqrt(var + 1e-5)
            x = F.conv2d(x, q(w1 * scale[:, None, None, None]), (b1 - mean) * scale + beta)
        else:
            x = F.conv2d(x, q(sd[prefix + "conv1.weight"].to(x.dtype)), sd[prefix + "conv1.bias"].to(x.dtype))
    out = act(x + y)
    return out if last_block else q(out)


def model_forward(x: torch.Tensor, configs: Sequence[dict], sd: Dict[str, torch.Tensor],
                  round_bf16: bool = False) -> torch.Tensor:
    """buildModel.forward for a list of CNN / ResNet blocks (src/models/modelbuilder.py:21-31)."""
    for m, config in enumerate(configs):
        name = config["name"].upper()
        if name == "CNN":
            x = cnn_block_forward(x, config, sd, f"models.{m}.", round_bf16)
        elif name in ("RESNETBLOCK", "RESNETBASICBLOCK", "RESNETBOTTLENECKBLOCK"):        # src/models/tools.py:93-95
            if m == 0 and round_bf16:
                x = x.to(torch.bfloat16).to(x.dtype)
            x = resnet_block_forward(x, config, sd, f"models.{m}.", round_bf16, last_block=(m == len(configs) - 1))
        elif name == "RESNEXTBLOCK":                                                      # src/models/tools.py:96-98
            if m == 0 and round_bf16:
                x = x.to(torch.bfloat16).to(x.dtype)
            x = resnet_block_forward(x, config, sd, f"models.{m}.", round_bf16, last_block=(m == len(configs) - 1),
                                     resnext=True)
        else:
            raise ValueError(f"oracle covers only CNN, ResNet and ResNeXt blocks, got {config['name']}")
    return x


def correction_add(u: torch.Tensor, v: torch.Tensor, configs, sd, input_scale: float = 1.0,
                   round_bf16: bool = False):
    """``v += model(v)`` (src/inference.py:102) with the repairs of SURVEY.md section 3.2 D5:
    x = stack((u, v), 1) * input_scale;  u += y[:, 0];  v += y[:, 1]."""
    x = torch.stack((u, v), dim=1) * input_scale
    y = model_forward(x, configs, sd, round_bf16)
    return u + y[:, 0], v + y[:, 1]


def rollout(u, v, solver_cfg, nsteps: int, configs=None, sd=None, input_scale: float = 1.0,
            round_bf16: bool = False, dt=None):
    """The hot loop src/inference.py:100-102: solver step, then correction add."""
    from . import solver as S
    inv = S.inverse_eigenvalues(solver_cfg, u.dtype, u.device)
    p = None
    for _ in range(nsteps):
        u, v, p = S.rk4_step(u, v, solver_cfg, dt, inv)
        if configs is not None:
            u, v = correction_add(u, v, configs, sd, input_scale, round_bf16)
    return u, v, p
