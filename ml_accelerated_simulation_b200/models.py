"""Correction-network front end: the reference's ``buildModel`` / ``CNN`` module API and checkpoint layout
(reference src/models/modelbuilder.py:5-31, src/models/CNN.py:24-40, src/models/tools.py:1-34; weights
contract src/train.py:50), with the forward pass executed by the tcgen05 convolution kernels.

The module tree and parameter names are identical to the reference's (``models.{m}.layers.{i}[.0|.1].*``)
so ``load_state_dict`` / ``safetensors.torch.save_model`` / ``load_model`` interoperate with reference
checkpoints.  Parameters are created with the same ``nn.Conv2d`` / ``nn.BatchNorm2d`` constructors in the
same order, so a given ``torch.manual_seed`` yields the reference's initial weights.

On the hot path: the CNN block family (3x3 / stride 1 / padding 1 / groups 1, ReLU, dropout inactive (eval), optional
BatchNorm (folded), in_channels = 2, hidden <= 64, out_channels <= 16) and chains of ResNet basic / bottleneck and
ResNeXt blocks (1x1 and 3x3, grouped or dense, BatchNorm folded, shortcuts fused into the layer epilogues).
Anything else raises -- there is no PyTorch fallback.
"""
from __future__ import annotations

import json
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn

from .plan import Plan, PlanConfig


def _as_list(param, n: int, what: str) -> list:
    if isinstance(param, list):
        if len(param) != n:
            raise ValueError(f"{what}: list length must equal the number of layers ({n})")
        return list(param)
    if isinstance(param, (int, float)):
        return [param] * n
    raise TypeError(f"{what}: expected a number or a list")


def _channels(structure: dict) -> List[int]:
    vals = list(structure.values())
    if len(vals) != 3:
        raise TypeError("structures must have in_channels, hidden_channels, out_channels")
    return [int(vals[0])] + [int(c) for c in vals[1]] + [int(vals[2])]


class CNN(nn.Module):
    """Parameter container with the reference CNN block's layout; forward runs on the B200 kernels."""

    def __init__(self, config: dict):
        super().__init__()
        chans = _channels(config["structures"])
        n = len(chans) - 1
        ks = _as_list(config["kernel_sizes"], n, "kernel_sizes")
        st = _as_list(config["strides"], n, "strides")
        pd = _as_list(config["paddings"], n, "paddings")
        gr = _as_list(config["group"], n, "group")
        self.dropouts = _as_list(config["dropouts"], n, "dropouts")
        if str(config["activation_func"]).lower() != "relu":
            raise NotImplementedError("only ReLU CNN blocks are on the B200 hot path")
        if any(k != 3 for k in ks) or any(s != 1 for s in st) or any(p != 1 for p in pd) or any(g != 1 for g in gr):
            raise NotImplementedError("only 3x3 / stride 1 / padding 1 / groups 1 convolutions are on the B200 hot path")
        if n < 2 or chans[0] != 2 or max(chans[1:-1]) > 64 or chans[-1] > 16:
            raise NotImplementedError("supported shapes: 2 -> (<=64)* -> (<=16) channels, at least two layers")
        self.channels = chans
        self.bn = bool(config["bn"])
        if self.bn:
            self.layers = nn.ModuleList([nn.Sequential(nn.Conv2d(chans[i], chans[i + 1], 3, 1, 1),
                                                       nn.BatchNorm2d(chans[i + 1])) for i in range(n)])
        else:
            self.layers = nn.ModuleList([nn.Conv2d(chans[i], chans[i + 1], 3, 1, 1) for i in range(n)])

    def folded_layers(self) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """[(weight, bias)] in fp64 with eval-mode BatchNorm folded into the convolution."""
        out = []
        for layer in self.layers:
            if self.bn:
                conv, bn = layer[0], layer[1]
                w = conv.weight.detach().double().cpu()
                b = conv.bias.detach().double().cpu()
                scale = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
                w = w * scale[:, None, None, None]
                b = (b - bn.running_mean.detach().double().cpu()) * scale + bn.bias.detach().double().cpu()
            else:
                w = layer.weight.detach().double().cpu()
                b = layer.bias.detach().double().cpu()
            out.append((w.contiguous(), b.contiguous()))
        return out

    def param_version(self) -> tuple:
        return tuple(int(p._version) for p in list(self.parameters()) + list(self.buffers()))


_RESNET_NAMES = ("RESNETBLOCK", "RESNETBASICBLOCK", "RESNETBOTTLENECKBLOCK")     # reference src/models/tools.py:93-95
_RESNEXT_NAMES = ("RESNEXTBLOCK",)                                                 # reference src/models/tools.py:96-98


def _fold_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode BatchNorm folded into the preceding convolution, fp64."""
    w = conv.weight.detach().double().cpu()
    b = conv.bias.detach().double().cpu()
    scale = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + bn.eps)
    return w * scale[:, None, None, None], (b - bn.running_mean.detach().double().cpu()) * scale + bn.bias.detach().double().cpu()


def _dense3x3(w: torch.Tensor, groups: int) -> torch.Tensor:
    """A 1x1 or 3x3, grouped or dense convolution weight [cout, cin/groups, k, k] as the dense 3x3 weight
    [cout, cin, 3, 3] of the SAME linear map (zero padding of width (k-1)/2, stride 1): a 1x1 kernel sits in the
    centre tap, a grouped kernel on the block diagonal.  The tensor-core layer kernel is a dense 64-channel 3x3
    implicit GEMM, so the structural zeros ride along as zero weights -- exact, not an approximation."""
    cout, cin_g, k, _ = w.shape
    cin = cin_g * groups
    out = torch.zeros(cout, cin, 3, 3, dtype=w.dtype)
    og = cout // groups
    lo = (3 - k) // 2
    for g in range(groups):
        out[g * og:(g + 1) * og, g * cin_g:(g + 1) * cin_g, lo:lo + k, lo:lo + k] = w[g * og:(g + 1) * og]
    return out


class ResNetBlock(nn.Module):
    """Parameter container with the reference ResNetBlock's layout (src/models/ResNet.py:26-39): every layer is
    ``Sequential(Conv2d, BatchNorm2d)``, optional 1x1 ``conv1`` on the block input; forward = ``act(x + y)`` (:41-50).
    On the B200 path: basic blocks (3x3, 3x3) and bottleneck blocks (1x1, 3x3, 1x1), stride 1, "same" padding, any
    group count that divides the channels, ReLU, two or more layers, at most 64 channels."""

    def __init__(self, config: dict):
        super().__init__()
        chans = _channels(config["structures"])
        n = len(chans) - 1
        ks = _as_list(config["kernel_sizes"], n, "kernel_sizes")
        st = _as_list(config["strides"], n, "strides")
        pd = _as_list(config["paddings"], n, "paddings")
        gr = _as_list(config["group"], n, "group")
        self.dropouts = _as_list(config["dropouts"], n, "dropouts")
        if str(config["activation_func"]).lower() != "relu":
            raise NotImplementedError("only ReLU ResNet blocks are on the B200 hot path")
        if any(k not in (1, 3) for k in ks) or any(s != 1 for s in st) or any(2 * p != k - 1 for p, k in zip(pd, ks)):
            raise NotImplementedError("only 1x1 / 3x3, stride 1, same-padding convolutions are on the B200 hot path")
        if n < 2:
            raise NotImplementedError("ResNet blocks on the B200 path need at least two conv layers")
        self.channels = chans
        self.groups = [int(g) for g in gr]
        self.layers = nn.ModuleList([nn.Sequential(nn.Conv2d(chans[i], chans[i + 1], kernel_size=ks[i], stride=1,
                                                             padding=pd[i], groups=gr[i]),
                                                   nn.BatchNorm2d(chans[i + 1])) for i in range(n)])
        self.conv1 = nn.Conv2d(chans[0], chans[-1], kernel_size=1) if config.get("1x1_conv", False) else None
        if self.conv1 is None and chans[0] != chans[-1]:
            raise ValueError("identity shortcut needs in_channels == out_channels (the reference would fail to add)")

    def folded_layers(self) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        out = []
        for layer, g in zip(self.layers, self.groups):
            w, b = _fold_bn(layer[0], layer[1])
            out.append((_dense3x3(w, g).contiguous(), b.contiguous()))
        return out

    def shortcut_weights(self):
        """(w1x1 [cout, cin, 1, 1], bias) of ``conv1`` in fp64, or None for the identity shortcut."""
        if self.conv1 is None:
            return None
        return self.conv1.weight.detach().double().cpu(), self.conv1.bias.detach().double().cpu()


class ResNeXtBlock(ResNetBlock):
    """Reference ``ResNeXtBlock`` (src/models/ResNet.py:54-92): the layer stack of a ResNetBlock built from the same
    config (:76 -- the reference constructs a whole ResNetBlock, including a ``conv1`` it then discards, and keeps its
    ``layers``; the construction order is repeated here so a seed yields the reference's initial weights), and a
    shortcut ``conv1 = Sequential(Conv2d 1x1, BatchNorm2d)`` (:78-80)."""

    def __init__(self, config: dict):
        super().__init__(config)                        # layers (+ the discarded plain conv1, in the reference's RNG order)
        chans = self.channels
        if config["1x1_conv"]:                          # the reference indexes the key directly (KeyError if absent)
            self.conv1 = nn.Sequential(nn.Conv2d(chans[0], chans[-1], kernel_size=1), nn.BatchNorm2d(chans[-1]))
        else:
            self.conv1 = None

    def shortcut_weights(self):
        if self.conv1 is None:
            return None
        return _fold_bn(self.conv1[0], self.conv1[1])


class buildModel(nn.Module):
    """Drop-in for the reference's ``buildModel(configs)`` (src/models/modelbuilder.py:5-31) for the conv families on the
    B200 path: one CNN block (e.g. baselines/MLACFD.json) or a chain of ResNet basic blocks (e.g. basicResnet1.json)."""

    def __init__(self, configs: Sequence[dict]):
        super().__init__()
        names = [str(c["name"]).upper() for c in configs]
        if len(configs) == 1 and names[0] == "CNN":
            self.family = "cnn"
            self.models = nn.ModuleList([CNN(configs[0])])
        elif configs and all(nm in _RESNET_NAMES + _RESNEXT_NAMES for nm in names):
            self.family = "resnet"
            configs = [dict(c) for c in configs]
            for i in range(len(configs) - 1):             # the reference corrects mismatching channel counts (:12-15)
                configs[i + 1]["structures"] = dict(configs[i + 1]["structures"])
                configs[i + 1]["structures"]["in_channels"] = configs[i]["structures"]["out_channels"]
            self.models = nn.ModuleList([(ResNeXtBlock if nm in _RESNEXT_NAMES else ResNetBlock)(c)
                                         for nm, c in zip(names, configs)])
            ch = [b.channels for b in self.models]
            if ch[0][0] != 2 or ch[-1][-1] > 4 or any(c > 64 for b in ch for c in b[1:-1]) or any(b[-1] > 64 for b in ch[:-1]):
                raise NotImplementedError("supported ResNet shapes: 2 -> (<=64)* -> (<=4) channels")
        else:
            raise NotImplementedError("the B200 hot path covers buildModel([CNN block]) and chains of ResNet (basic / "
                                      "bottleneck) and ResNeXt blocks with at most 64 channels; the 128-channel "
                                      "CHANNELAWARE net (smartConv.json) needs a K=128 / N=128 layer kernel that is not built")
        self.models_name = [c["name"] for c in configs]
        self.input_scale = 1.0
        self._plans: Dict[tuple, Tuple[Plan, tuple]] = {}
        self.eval()

    @property
    def cnn(self):
        return self.models[0]

    def param_version(self) -> tuple:
        return tuple(int(p._version) for p in list(self.parameters()) + list(self.buffers()))

    def flat_layers(self):
        """(layers, shortcuts, final_relu) in the form Plan.set_cnn takes."""
        if self.family == "cnn":
            return self.cnn.folded_layers(), None, False
        layers, shortcuts = [], {}
        for blk in self.models:
            start = len(layers)
            layers += blk.folded_layers()
            end = len(layers) - 1
            sc = blk.shortcut_weights()
            shortcuts[end] = (start - 1, None, None) if sc is None else (start - 1, sc[0], sc[1])
        return layers, shortcuts, True

    def attach(self, plan) -> None:
        """Upload (BN-folded, bf16-packed) weights into ``plan`` so plan.step(..., apply_cnn=True) can run."""
        layers, shortcuts, final_relu = self.flat_layers()
        plan.set_cnn(layers, shortcuts, final_relu)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise RuntimeError("the B200 CNN path is inference-only; call .eval()")
        if not x.is_cuda:
            raise RuntimeError("the B200 CNN path runs on CUDA tensors only (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 2:
            raise ValueError("expected input [batch, 2, nx, ny]")
        b, _, nx, ny = x.shape
        key = (b, nx, ny, x.device.index)
        ver = self.param_version()
        entry = self._plans.get(key)
        if entry is None or entry[1] != ver:
            plan = entry[0] if entry is not None else Plan(PlanConfig(nx=nx, ny=ny, batch=b, dt=1.0,
                                                                     device=x.device.index or 0))
            self.attach(plan)
            self._plans[key] = (plan, ver)
        plan = self._plans[key][0]
        u = x[:, 0].to(torch.float32).contiguous()
        v = x[:, 1].to(torch.float32).contiguous()
        return plan.cnn_forward(u, v, 1.0).to(x.dtype)


def load_config(path: str) -> list:
    with open(path, "r") as f:
        return json.load(f)


def load_model(config_path: str, weights_path: str) -> buildModel:
    """Build from a reference JSON config and load a reference ``.safetensors`` checkpoint (src/train.py:50)."""
    import safetensors.torch as st
    model = buildModel(load_config(config_path))
    sd = st.load_file(weights_path)
    model.load_state_dict(sd)
    return model
