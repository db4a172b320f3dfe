"""``torch_cfd.fvm`` names on the rollout hot path: ``RKStepper`` and ``NavierStokes2DFVMProjection``
(reference call sites: src/inference.py:64,84-93,101; src/data_generation.py:129,167-187,206-212;
test/sim_test.py:54,57-67,81).

``RKStepper.forward(v, dt, equation=ns2d) -> (v_next, p)`` runs ONE classic-RK4 step with a spectral
pressure projection after every stage, entirely in the sm_100a kernels behind include/mlsim.h.  The
arithmetic is fp32 on the device whatever dtype the caller holds (the reference asks for float64; the
result is cast back to the caller's dtype).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .grids import Grid, GridVariable, GridVariableVector
from .plan import Plan, PlanConfig

_PLAN_CACHE: Dict[tuple, Plan] = {}


def _plan_for(grid: Grid, like: torch.Tensor, dt: float, equation) -> Plan:
    if not like.is_cuda:
        raise RuntimeError("the B200 rollout path runs on CUDA tensors only (no CPU fallback)")
    batch = like.shape[0] if like.dim() == 3 else 1
    if equation is not None:
        forcing = equation.forcing
        key_phys = (equation.viscosity, equation.density, equation.drag,
                    None if forcing is None else (forcing.wave_number, forcing.scale))
    else:
        key_phys = None
    key = (grid.shape, grid.domain, batch, float(dt), like.device.index, key_phys)
    plan = _PLAN_CACHE.get(key)
    if plan is None:
        (x0, x1), (y0, y1) = grid.domain
        cfg = PlanConfig(nx=grid.shape[0], ny=grid.shape[1], batch=batch, lx=x1 - x0, ly=y1 - y0, dt=float(dt),
                         device=like.device.index or 0)
        if equation is not None:
            cfg.viscosity = float(equation.viscosity)
            cfg.density = float(equation.density)
            cfg.drag = float(equation.drag)
            if equation.forcing is None:
                cfg.forcing_scale = 0.0
            else:
                cfg.forcing_wavenumber = equation.forcing.wave_number
                cfg.forcing_scale = equation.forcing.scale
        plan = Plan(cfg)
        _PLAN_CACHE[key] = plan
    return plan


class NavierStokes2DFVMProjection:
    """Parameter holder for the incompressible 2-D Navier-Stokes equation on a periodic MAC grid
    (van-Leer TVD advection, 5-point diffusion, forcing, linear drag, exact spectral projection)."""

    def __init__(self, viscosity: float, grid: Grid, bcs=None, density: float = 1.0, drag: float = 0.0,
                 forcing=None, solver=None, **_ignored):
        self.viscosity = float(viscosity)
        self.grid = grid
        self.bcs = bcs
        self.density = float(density)
        self.drag = float(drag)
        self.forcing = forcing
        self.solver = solver

    def to(self, device=None, *args, **kwargs):
        return self


class RKStepper:
    def __init__(self, method: str = "classic_rk4", requires_grad: bool = False, dtype=torch.float64):
        if method != "classic_rk4":
            raise NotImplementedError(f"only classic_rk4 is on the hot path (got {method})")
        if requires_grad:
            raise NotImplementedError("the CUDA rollout path is inference-only")
        self.method = method
        self.dtype = dtype

    @classmethod
    def from_method(cls, method: str = "classic_rk4", requires_grad: bool = False, dtype=torch.float64, **_):
        return cls(method, requires_grad, dtype)

    def forward(self, v: GridVariableVector, dt: float, equation: NavierStokes2DFVMProjection
                ) -> Tuple[GridVariableVector, GridVariable]:
        u_var, v_var = v[0], v[1]
        grid = u_var.grid
        data = u_var.data
        squeeze = data.dim() == 2
        ud = u_var.data[None] if squeeze else u_var.data
        vd = v_var.data[None] if squeeze else v_var.data
        # one plan per (equation, batch, dt, device), remembered on the equation object: no key building per step
        cache = equation.__dict__.setdefault("_mlsim_plans", {})
        ck = (ud.shape[0], float(dt), ud.device.index)
        plan = cache.get(ck)
        if plan is None:
            plan = cache[ck] = _plan_for(grid, ud, dt, equation)
        if data.dtype == torch.float64:
            # the reference's dtype: fused fp64 <-> fp32 conversion inside the C ABI (mlsim_step_f64)
            uo, vo, po = plan.step_f64(ud.contiguous(), vd.contiguous(), nsteps=1, want_p=True)
        else:
            uf = ud.to(torch.float32, copy=True).contiguous()
            vf = vd.to(torch.float32, copy=True).contiguous()
            p = plan.new_field()
            plan.step(uf, vf, nsteps=1, apply_cnn=False, p=p)
            uo, vo, po = uf.to(data.dtype), vf.to(data.dtype), p.to(data.dtype)
        if squeeze:
            uo, vo, po = uo[0], vo[0], po[0]
        vec = GridVariableVector((GridVariable(uo, u_var.offset, grid, u_var.bc),
                                  GridVariable(vo, v_var.offset, grid, v_var.bc)))
        return vec, GridVariable(po, grid.cell_center, grid, u_var.bc)

    __call__ = forward
