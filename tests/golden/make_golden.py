"""Generate golden vectors from the reference's OWN model code (run in the build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference/src/models`` unmodified (``buildModel``, src/models/modelbuilder.py:5-31),
builds the hot-path correction CNN from the reference's JSON configs, runs it in fp64 exactly as the
reference does (``torch.set_default_dtype(torch.float64)``, src/inference.py:60) and saves

  * ``mlacfd_weights.safetensors``  -- written by ``safetensors.torch.save_model`` as src/train.py:50 does
    (this is the drop-in weight layout the product loader must read);
  * ``cnn1_bn_weights.safetensors`` -- the ``bn: true`` variant (src/models/configs/cnn1.json) with
    non-trivial BatchNorm running statistics;
  * ``cnn_golden.safetensors``      -- inputs and reference outputs.

/root/reference does not exist on the GPU box; only the committed outputs travel.
"""
import json
import os
import sys

import torch
import safetensors.torch as st

sys.dont_write_bytecode = True
REF = "/root/reference/src"
sys.path.insert(0, REF)
from models import buildModel  # noqa: E402  (the reference's own code)

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(2025)                      # the reference's seed (src/train.py:57, src/speed.py:96)
    out = {}

    with open(os.path.join(REF, "models/configs/fullmodels/baselines/MLACFD.json")) as f:
        cfg = json.load(f)
    with open(os.path.join(HERE, "MLACFD.json"), "w") as f:
        json.dump(cfg, f, indent=1)
    model = buildModel(cfg).eval()
    with torch.no_grad():
        last = model.models[0].layers[-1]
        last.weight.mul_(1e-2)                   # trained-like correction magnitude
        last.bias.mul_(1e-2)
    st.save_model(model, os.path.join(HERE, "mlacfd_weights.safetensors"))
    g = torch.Generator().manual_seed(42)
    x = torch.randn(2, 2, 16, 24, generator=g) * 3.0
    with torch.inference_mode():
        out["mlacfd_x"] = x
        out["mlacfd_y"] = model(x.clone())
        xs = x / 7.0                              # training contract: src/data/dataloader.py:78-79
        out["mlacfd_y_scaled"] = model(xs)

    with open(os.path.join(REF, "models/configs/cnn1.json")) as f:
        cfg1 = json.load(f)
    with open(os.path.join(HERE, "cnn1.json"), "w") as f:
        json.dump(cfg1, f, indent=1)
    m1 = buildModel(cfg1)
    m1.train()
    with torch.no_grad():                         # populate BN running stats with something non-trivial
        for _ in range(3):
            m1(torch.randn(4, 2, 8, 8, generator=g) * 2.0)
    m1.eval()
    st.save_model(m1, os.path.join(HERE, "cnn1_bn_weights.safetensors"))
    x1 = torch.randn(3, 2, 8, 16, generator=g)
    with torch.inference_mode():
        out["cnn1_x"] = x1
        out["cnn1_y"] = m1(x1.clone())

    st.save_file({k: v.contiguous() for k, v in out.items()}, os.path.join(HERE, "cnn_golden.safetensors"))
    for k, v in out.items():
        print(k, tuple(v.shape), v.dtype, float(v.abs().max()))


if __name__ == "__main__":
    main()
