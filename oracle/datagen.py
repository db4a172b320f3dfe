"""CPU oracle for the data-generation helpers next to the rollout path.  TEST INFRASTRUCTURE ONLY.

Restates, in plain torch indexing, the reference's
  * ``block_reduce``                                  src/data_generation.py:50-62
  * ``slice_along_axis``                              src/data_generation.py:72-89
  * ``downsample_staggered_velocity_component``       src/data_generation.py:91-94
  * ``downsample_staggered_velocity``                 src/data_generation.py:97-111
  * the pair loop of ``main``                         src/data_generation.py:203-213

PINNED: ``tests/golden/make_golden_datagen.py`` executes the reference's OWN function bodies (extracted from the
unmodified source file with ``ast``; the module itself cannot be imported because ``torch_cfd`` and ``pydrive2`` are
absent) and commits inputs/outputs as ``tests/golden/datagen_golden.safetensors``; ``tests/test_oracle_datagen.py``
checks this file against them bit for bit.
"""
from __future__ import annotations

from typing import Tuple

import torch


def downsample_component(u: torch.Tensor, direction: int, factor: int) -> torch.Tensor:
    """One un-batched component [nx, ny] (the reference vmaps this over the batch, :106)."""
    nx, ny = u.shape
    if nx % factor or ny % factor:
        raise ValueError("`block_size` must divide `array.shape`")
    if direction == 0:
        w = u[factor - 1::factor, :]                                     # [nx/f, ny]
        return w.reshape(nx // factor, ny // factor, factor).mean(dim=-1)
    w = u[:, factor - 1::factor]                                         # [nx, ny/f]
    return w.reshape(nx // factor, factor, ny // factor).mean(dim=1)


def downsample_staggered_velocity(u: torch.Tensor, v: torch.Tensor, factor: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched [b, nx, ny] fields: component j is sliced along direction j (u: x, v: y)."""
    cu = torch.stack([downsample_component(a, 0, factor) for a in u])
    cv = torch.stack([downsample_component(a, 1, factor) for a in v])
    return cu, cv


def datagen_pair(u, v, fine_cfg, coarse_cfg, n_between: int, n_inner: int):
    """src/data_generation.py:203-213: returns ((u_f, v_f), (u_c, v_c)) of one stored pair."""
    from . import solver as S
    inv_f = S.inverse_eigenvalues(fine_cfg, u.dtype)
    inv_c = S.inverse_eigenvalues(coarse_cfg, u.dtype)
    for _ in range(n_between):
        u, v, _ = S.rk4_step(u, v, fine_cfg, None, inv_f)
    cu, cv = downsample_staggered_velocity(u, v, fine_cfg.nx // coarse_cfg.nx)
    for _ in range(n_inner):
        u, v, _ = S.rk4_step(u, v, fine_cfg, None, inv_f)
    cu, cv, _ = S.rk4_step(cu, cv, coarse_cfg, None, inv_c)
    return (u, v), (cu, cv)
