"""Fused rollout driver: the hot loop of reference src/inference.py:100-102 on device-resident state."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .plan import Plan, PlanConfig


class Rollout:
    """``for t in range(n): v, p = step_fn.forward(v, dt, equation=ns2d); v += model(v)``.

    State is two contiguous fp32 tensors ``[batch, nx, ny]`` that stay in HBM for the whole rollout.
    ``input_scale``: 1.0 reproduces the literal reference loop; 1/7 is what trained weights expect
    (src/data/dataloader.py:78-79) -- see SURVEY.md section 3.2, defect D5.
    """

    def __init__(self, cfg: PlanConfig, model=None, input_scale: float = 1.0):
        self.plan = Plan(cfg)
        self.cfg = cfg
        self.dt = self.plan.dt
        self.input_scale = float(input_scale)
        self.has_model = model is not None
        if model is not None:
            model.attach(self.plan)

    def step(self, u: torch.Tensor, v: torch.Tensor, nsteps: int = 1, p: Optional[torch.Tensor] = None,
             apply_cnn: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        use = self.has_model if apply_cnn is None else apply_cnn
        self.plan.step(u, v, nsteps=nsteps, apply_cnn=use, input_scale=self.input_scale, p=p)
        return u, v

    def step_host(self, hu: torch.Tensor, hv: torch.Tensor, nsteps: int = 1, hp: Optional[torch.Tensor] = None):
        self.plan.rollout_host(hu, hv, nsteps, apply_cnn=self.has_model, input_scale=self.input_scale, hp=hp)
        return hu, hv

    def cells_per_step(self) -> int:
        return self.cfg.batch * self.cfg.nx * self.cfg.ny
