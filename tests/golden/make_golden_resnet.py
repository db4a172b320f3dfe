"""Golden vectors for the ResNet-family correction nets from the reference's OWN model code (build container only).

    python tests/golden/make_golden_resnet.py [basicResnet1 | basicBottleneck1 | basicResNeXt1 ...]

Builds ``buildModel(configs/fullmodels/<name>.json)`` from the unmodified ``/root/reference/src/models``
(ResNetBlock: src/models/ResNet.py:26-50, ResNeXtBlock: :54-92), gives every BatchNorm non-trivial running statistics,
rounds the parameters to fp32-representable values (so the committed weight file can be fp32 and half the size), runs
the reference forward in fp64 eval mode and saves ``<name>.json``, ``<tag>_weights.safetensors`` (state_dict keys of the
reference) and ``<tag>_golden.safetensors`` (input / output); tag = resnet | bottleneck | resnext.
"""
import json
import os
import sys

import safetensors.torch as st
import torch

sys.dont_write_bytecode = True
REF = "/root/reference/src"
sys.path.insert(0, REF)
from models import buildModel  # noqa: E402  (the reference's own code)

HERE = os.path.dirname(os.path.abspath(__file__))


TAGS = {"basicResnet1": "resnet", "basicBottleneck1": "bottleneck", "basicResNeXt1": "resnext"}


def main(name="basicResnet1"):
    tag = TAGS[name]
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(2025)
    with open(os.path.join(REF, f"models/configs/fullmodels/{name}.json")) as f:
        cfg = json.load(f)
    with open(os.path.join(HERE, f"{name}.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    model = buildModel(cfg)
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.5 + 0.75)
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) * 0.5 + 0.75)
                mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.1)
        for p in list(model.parameters()) + list(model.buffers()):
            if p.dtype.is_floating_point:
                p.copy_(p.float().double())
    model.eval()
    x = torch.randn(2, 2, 64, 64, generator=g)
    with torch.inference_mode():
        y = model(x)
    sd = {k: (v.float() if v.dtype.is_floating_point else v).contiguous() for k, v in model.state_dict().items()}
    st.save_file(sd, os.path.join(HERE, f"{tag}_weights.safetensors"))
    st.save_file({"x": x, "y": y.contiguous()}, os.path.join(HERE, f"{tag}_golden.safetensors"))
    print("params", sum(p.numel() for p in model.parameters()), "y", tuple(y.shape), float(y.abs().max()))


if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["basicResnet1"]):
        main(nm)
