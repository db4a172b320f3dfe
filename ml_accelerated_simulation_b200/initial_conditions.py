"""``torch_cfd.initial_conditions.filtered_velocity_field`` (reference src/inference.py:76-78).

Setup code, run once per rollout (not on the per-step hot path): a band-limited random field is made
divergence free by the CUDA projection kernels and rescaled to ``max_velocity``.
The spectral filtering of the white noise runs on the projection's own transform kernels
(``mlsim_spectral_multiply``); no library FFT is involved.
"""
from __future__ import annotations

import math

import torch

from .boundaries import PeriodicBoundaryConditions
from .grids import Grid, GridVariable, GridVariableVector
from .plan import Plan, PlanConfig


def _spectral_filter(grid: Grid, peak_wavenumber: float) -> torch.Tensor:
    """Host fp64 table [nx, ny//2+1]: log-normal band around ``peak_wavenumber`` divided by |k| (0 at k = 0)."""
    nx, ny = grid.shape
    (x0, x1), (y0, y1) = grid.domain
    ix = torch.arange(nx, dtype=torch.float64)
    kx = torch.where(ix < (nx + 1) // 2, ix, ix - nx) * (2 * math.pi / (x1 - x0))         # fftfreq ordering
    ky = torch.arange(ny // 2 + 1, dtype=torch.float64) * (2 * math.pi / (y1 - y0))
    kk = torch.sqrt(kx[:, None] ** 2 + ky[None, :] ** 2)
    logvar = 0.25
    mu = math.log(peak_wavenumber) + logvar
    safe = torch.where(kk > 0, kk, torch.ones_like(kk))
    pdf = torch.exp(-(torch.log(safe) - mu) ** 2 / (2 * logvar)) / (safe * math.sqrt(2 * math.pi * logvar))
    return torch.where(kk > 0, pdf / safe, torch.zeros_like(kk))


def filtered_velocity_field(grid: Grid, max_velocity: float = 1.0, peak_wavenumber: float = 3.0,
                            iterations: int = 3, random_state: int = 0, device=None, batch_size: int = 1,
                            dtype=None) -> GridVariableVector:
    device = torch.device(device if device is not None else "cuda")
    if device.type != "cuda":
        raise RuntimeError("filtered_velocity_field of the B200 build needs a CUDA device")
    dtype = dtype or torch.get_default_dtype()
    nx, ny = grid.shape
    g = torch.Generator(device="cpu").manual_seed(int(random_state))
    filt = _spectral_filter(grid, peak_wavenumber)
    (x0, x1), (y0, y1) = grid.domain
    plan = Plan(PlanConfig(nx=nx, ny=ny, batch=batch_size, lx=x1 - x0, ly=y1 - y0, dt=1.0,
                           device=device.index or 0))
    comps = []
    for _ in range(2):
        noise = torch.randn(batch_size, nx, ny, generator=g, dtype=torch.float64).to(device, torch.float32).contiguous()
        comps.append(plan.spectral_multiply(noise, filt))
    # normalise first so fp32 projection works at O(1) magnitudes
    u, v = comps
    plan.normalize_speed(u, v, max_velocity)
    for _ in range(iterations):
        plan.project(u, v, want_p=False)
        plan.normalize_speed(u, v, max_velocity)
    plan.close()
    bc = PeriodicBoundaryConditions()
    return GridVariableVector((GridVariable(u.to(dtype), (1.0, 0.5), grid, bc),
                               GridVariable(v.to(dtype), (0.5, 1.0), grid, bc)))
