"""``torch_cfd.finite_differences`` subset: divergence (reference test/sim_test.py:42)."""
from __future__ import annotations

import torch

from .grids import GridVariable


def divergence(v) -> GridVariable:
    """Backward-difference divergence of a staggered velocity at cell centres (CUDA kernel)."""
    from .fvm import _plan_for
    u, w = v[0], v[1]
    plan = _plan_for(u.grid, u.data, dt=1.0, equation=None)
    uf = u.data.to(torch.float32).contiguous()
    vf = w.data.to(torch.float32).contiguous()
    d = plan.divergence(uf, vf).to(u.data.dtype)
    return GridVariable(d, u.grid.cell_center, u.grid, u.bc)
