"""Golden vectors for the data-generation helpers from the reference's OWN code (build container only).

    python tests/golden/make_golden_datagen.py

``/root/reference/src/data_generation.py`` cannot be imported (it imports ``torch_cfd`` and ``pydrive2`` and writes a
file at import time), so the function definitions it needs are lifted verbatim out of the unmodified source with
``ast`` and executed: ``block_reduce`` (:50-62), ``_normalize_axis`` (:65-70), ``slice_along_axis`` (:72-89),
``downsample_staggered_velocity_component`` (:91-94).  The per-component vmap of ``downsample_staggered_velocity``
(:106) is applied here exactly as the reference applies it.  Output: ``datagen_golden.safetensors``.
"""
import ast
import os
from functools import partial

import safetensors.torch as st
import torch
import torch.utils._pytree as pytree

SRC = "/root/reference/src/data_generation.py"
HERE = os.path.dirname(os.path.abspath(__file__))
WANTED = {"block_reduce", "_normalize_axis", "slice_along_axis", "downsample_staggered_velocity_component"}


def reference_functions():
    tree = ast.parse(open(SRC).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    assert {n.name for n in body} == WANTED
    ns = {"torch": torch, "pytree": pytree, "partial": partial}
    exec(compile(ast.Module(body=body, type_ignores=[]), SRC, "exec"), ns)
    return ns


def main():
    ns = reference_functions()
    comp = ns["downsample_staggered_velocity_component"]
    g = torch.Generator().manual_seed(1234)
    out = {}
    for name, (b, nx, ny, f) in {"a": (2, 64, 64, 16), "b": (3, 32, 48, 4), "c": (1, 128, 64, 8)}.items():
        u = torch.randn(b, nx, ny, generator=g, dtype=torch.float64)
        v = torch.randn(b, nx, ny, generator=g, dtype=torch.float64)
        cu = torch.vmap(partial(comp, direction=0, factor=f))(u)          # src/data_generation.py:106, j = 0
        cv = torch.vmap(partial(comp, direction=1, factor=f))(v)          # j = 1
        out[f"{name}_u"], out[f"{name}_v"], out[f"{name}_cu"], out[f"{name}_cv"] = u, v, cu.contiguous(), cv.contiguous()
        out[f"{name}_factor"] = torch.tensor([f])
    st.save_file(out, os.path.join(HERE, "datagen_golden.safetensors"))
    print({k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
