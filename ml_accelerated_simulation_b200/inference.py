"""Repaired, runnable version of the reference's rollout script (src/inference.py:52-111) on the B200 path.

Repairs: D1 model built with ``buildModel(config)``; D2 no ``__metadata__`` pop;
D3 explicit ``--weights`` path instead of a per-process-salted hash; D4 ``v = v0`` before the loop;
D5 ``x = stack((u, v), 1) * input_scale``; ``u += y[:, 0]``; ``v += y[:, 1]`` fused into the last conv.
"""
from __future__ import annotations

import argparse
import math

import torch

from . import boundaries, grids
from .equations import stable_time_step
from .forcings import KolmogorovForcing
from .initial_conditions import filtered_velocity_field
from .models import buildModel, load_config, load_model
from .plan import PlanConfig
from .rollout import Rollout


def main(model_config: str, weights: str = None, steps: int = None, n: int = 64, batch: int = 1,
         input_scale: float = 1.0, simulation_time: float = 30.0, seed: int = 42, quiet: bool = False):
    density, max_velocity, peak_wavenumber, cfl, viscosity = 1.0, 7.0, 4.0, 0.5, 1e-3
    device = torch.device("cuda:0")
    diam = 2 * math.pi
    grid = grids.Grid((n, n), domain=((0, diam), (0, diam)), device=device)
    dt = stable_time_step(dx=min(grid.step), max_velocity=max_velocity, max_courant_number=cfl, viscosity=viscosity)
    v0 = filtered_velocity_field(grid, max_velocity, peak_wavenumber, iterations=16, random_state=seed,
                                 device=device, batch_size=batch, dtype=torch.float32)
    boundaries.get_pressure_bc_from_velocity(v0)
    forcing = KolmogorovForcing(diam=diam, wave_number=int(peak_wavenumber), grid=grid,
                                offsets=(v0[0].offset, v0[1].offset))
    model = load_model(model_config, weights) if weights else buildModel(load_config(model_config))
    cfg = PlanConfig(nx=n, ny=n, batch=batch, lx=diam, ly=diam, dt=dt, viscosity=viscosity, density=density,
                     drag=0.1, forcing_wavenumber=forcing.wave_number, forcing_scale=forcing.scale)
    ro = Rollout(cfg, model, input_scale=input_scale)
    u, v = v0[0].data.contiguous(), v0[1].data.contiguous()
    nsteps = steps if steps is not None else round(simulation_time / dt)
    ro.step(u, v, nsteps)
    torch.cuda.synchronize()
    if not quiet:
        print(f"rollout: {nsteps} steps, grid {n}x{n}, batch {batch}, max|u|={u.abs().max().item():.4f}, "
              f"finite={bool(torch.isfinite(u).all() and torch.isfinite(v).all())}")
    return u, v


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model_config", required=True, help="reference JSON config (e.g. MLACFD.json)")
    ap.add_argument("--weights", default=None, help="reference .safetensors checkpoint (default: seed-initialised)")
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--input_scale", type=float, default=1.0)
    with torch.inference_mode():
        main(**vars(ap.parse_args()))
