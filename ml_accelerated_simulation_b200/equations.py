"""``torch_cfd.equations`` names used by the reference."""
from __future__ import annotations


def stable_time_step(dx: float, max_velocity: float, max_courant_number: float, viscosity: float,
                     ndim: int = 2) -> float:
    """CFL / diffusion limited time step, as called at reference src/inference.py:68-73
    (SURVEY.md App. A.1): min(C*dx/max_velocity, dx^2/(viscosity*2^ndim))."""
    dt_adv = max_courant_number * dx / max_velocity
    if viscosity is None or viscosity <= 0:
        return dt_adv
    return min(dt_adv, dx * dx / (viscosity * (2 ** ndim)))
