"""CPU oracle for the solver half of the rollout hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference's solver arithmetic lives in the third-party package
``torch_cfd`` (PyPI ``torch-cfd`` / GitHub ``scaomath/torch-cfd``, a PyTorch port of
``google/jax-cfd``).  It is imported by the reference at ``src/inference.py:4-12``,
``src/data_generation.py:2-9``, ``src/speed.py:2-9`` and ``test/sim_test.py:2-8`` but is
neither vendored nor version-pinned there (no requirements/lock file exists) and is not
installed in the build container.  None of the reference's tests assert a numeric value at
that boundary (``test/sim_test.py:82-85`` checks only for NaN).  This file therefore restates
the *published* finite-volume algorithm (jax-cfd ``advect_van_leer_using_limiters`` /
``navier_stokes_rk`` / ``solve_fast_diag``; formulas in SURVEY.md App. A) and is validated by
physics known-answer tests (``tests/test_oracle_solver.py``).  Every recalled choice is a
keyword so it can be flipped if torch_cfd ever becomes available to diff against.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this module; the product path (``ml_accelerated_simulation_b200``) never does.

Conventions: arrays ``c[b, i, j]``, ``i`` <-> x (axis -2), ``j`` <-> y
(axis -1, contiguous); periodic; MAC staggering: u at (i+1, j+1/2)h, v at (i+1/2, j+1)h,
pressure/divergence at cell centres.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional, Tuple

import torch


@dataclasses.dataclass
class SolverConfig:
    """Physics constants; defaults are the reference's (src/inference.py:54-93)."""
    nx: int = 64
    ny: int = 64
    lx: float = 2 * math.pi          # src/inference.py:61,66  (diam = 2*pi)
    ly: float = 2 * math.pi
    viscosity: float = 1e-3          # src/inference.py:58
    density: float = 1.0             # src/inference.py:54
    drag: float = 0.1                # src/inference.py:89
    forcing_wavenumber: int = 4      # src/inference.py:81 (int(peak_wavenumber))
    forcing_scale: float = 1.0       # KolmogorovForcing default scale
    max_velocity: float = 7.0        # src/inference.py:55
    cfl: float = 0.5                 # src/inference.py:57
    advection: str = "van_leer"      # van_leer | upwind | lax_wendroff
    lw_dt_scale: float = 1.0         # 1.0: F receives the full dt (App. A.2); 0.0: central

    @property
    def hx(self) -> float:
        return self.lx / self.nx

    @property
    def hy(self) -> float:
        return self.ly / self.ny

    @property
    def dt(self) -> float:
        return stable_time_step(min(self.hx, self.hy), self.max_velocity, self.cfl, self.viscosity)


def stable_time_step(dx: float, max_velocity: float, max_courant_number: float, viscosity: float,
                     ndim: int = 2) -> float:
    """torch_cfd.equations.stable_time_step as called at src/inference.py:68-73 (App. A.1)."""
    dt_adv = max_courant_number * dx / max_velocity
    if viscosity <= 0:
        return dt_adv
    dt_diff = dx * dx / (viscosity * (2 ** ndim))
    return min(dt_adv, dt_diff)


def _S(c: torch.Tensor, k: int, axis: int) -> torch.Tensor:
    """Value at index +k along axis (periodic) == roll by -k."""
    return torch.roll(c, -k, axis)


def _sdiv(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """jax-cfd ``safe_div``: a / where(b != 0, b, 1).  (Where b == 0 the caller's result is
    multiplied by (c_high - c_low) == 0, so the default numerator never reaches the output.)"""
    return a / torch.where(b != 0, b, torch.ones_like(b))


def _van_leer(r: torch.Tensor) -> torch.Tensor:
    return torch.where(r > 0, _sdiv(2 * r, 1 + r), torch.zeros_like(r))


def _face_value(c: torch.Tensor, w: torch.Tensor, axis: int, h: float, dt: float,
                scheme: str) -> torch.Tensor:
    """TVD face interpolation of c along ``axis`` at the +1/2 face, face velocity w (App. A.3.2)."""
    cr = _S(c, 1, axis)
    pos = w > 0
    low = torch.where(pos, c, cr)
    if scheme == "upwind":
        return low
    courant = dt * w / h
    high = torch.where(pos, c + 0.5 * (1 - courant) * (cr - c), cr - 0.5 * (1 + courant) * (cr - c))
    if scheme == "lax_wendroff":
        return high
    cl = _S(c, -1, axis)
    crr = _S(c, 2, axis)
    d = cr - c
    r = torch.where(pos, _sdiv(c - cl, d), _sdiv(crr - cr, d))
    phi = _van_leer(r)
    return low - (low - high) * phi


def explicit_terms(u: torch.Tensor, v: torch.Tensor, cfg: SolverConfig, dt: float
                   ) -> Tuple[torch.Tensor, torch.Tensor]:
    """F(u, v; dt) = advection + diffusion + forcing + drag (App. A.3).  ``dt`` is the FULL step."""
    hx, hy = cfg.hx, cfg.hy
    X, Y = -2, -1
    dtl = dt * cfg.lw_dt_scale
    out = []
    for comp, c in enumerate((u, v)):
        if comp == 0:   # control volume of u: faces shifted along x
            wx = 0.5 * (u + _S(u, 1, X))
            wy = 0.5 * (v + _S(v, 1, X))
        else:           # control volume of v: faces shifted along y
            wx = 0.5 * (u + _S(u, 1, Y))
            wy = 0.5 * (v + _S(v, 1, Y))
        fx = _face_value(c, wx, X, hx, dtl, cfg.advection) * wx
        fy = _face_value(c, wy, Y, hy, dtl, cfg.advection) * wy
        adv = -((fx - _S(fx, -1, X)) / hx + (fy - _S(fy, -1, Y)) / hy)
        lap = (_S(c, -1, X) + _S(c, 1, X) - 2 * c) / (hx * hx) + (_S(c, -1, Y) + _S(c, 1, Y) - 2 * c) / (hy * hy)
        rhs = adv + (cfg.viscosity / cfg.density) * lap - cfg.drag * c
        if comp == 0 and cfg.forcing_scale != 0.0:
            rhs = rhs + forcing_profile(cfg, c.dtype, c.device) / cfg.density
        out.append(rhs)
    return out[0], out[1]


def forcing_profile(cfg: SolverConfig, dtype=torch.float64, device="cpu") -> torch.Tensor:
    """KolmogorovForcing (src/inference.py:81-82): fx = scale*sin(k*y_j) on u, y_j = (j+1/2)*hy."""
    j = torch.arange(cfg.ny, dtype=torch.float64, device=device)
    y = (j + 0.5) * cfg.hy
    return (cfg.forcing_scale * torch.sin(cfg.forcing_wavenumber * y)).to(dtype)


def laplacian_eigenvalues(cfg: SolverConfig, dtype=torch.float64, device="cpu") -> torch.Tensor:
    """Eigenvalues of the periodic 5-point Laplacian on the rfft2 grid, shape [nx, ny//2+1]."""
    kx = torch.arange(cfg.nx, dtype=torch.float64, device=device)
    ky = torch.arange(cfg.ny // 2 + 1, dtype=torch.float64, device=device)
    lx = (2 * torch.cos(2 * math.pi * kx / cfg.nx) - 2) / (cfg.hx ** 2)
    ly = (2 * torch.cos(2 * math.pi * ky / cfg.ny) - 2) / (cfg.hy ** 2)
    return (lx[:, None] + ly[None, :]).to(dtype)


def inverse_eigenvalues(cfg: SolverConfig, dtype=torch.float64, device="cpu") -> torch.Tensor:
    """Pseudo-inverse spectrum: 1/lambda where |lambda| > 10*eps(dtype), else 0 (App. A.4)."""
    lam = laplacian_eigenvalues(cfg, torch.float64, device)
    cutoff = 10 * torch.finfo(dtype).eps
    inv = torch.where(lam.abs() > cutoff, 1.0 / torch.where(lam != 0, lam, torch.ones_like(lam)),
                      torch.zeros_like(lam))
    return inv.to(dtype)


def divergence(u: torch.Tensor, v: torch.Tensor, cfg: SolverConfig) -> torch.Tensor:
    """fdm.divergence (test/sim_test.py:42): backward differences to cell centres."""
    return (u - _S(u, -1, -2)) / cfg.hx + (v - _S(v, -1, -1)) / cfg.hy


def pressure_projection(u: torch.Tensor, v: torch.Tensor, cfg: SolverConfig,
                        inv_eig: Optional[torch.Tensor] = None
                        ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """P(u, v) -> (u', v', p): exact spectral pseudo-inverse of the 5-point Laplacian (App. A.4)."""
    if inv_eig is None:
        inv_eig = inverse_eigenvalues(cfg, u.dtype, u.device)
    rhs = divergence(u, v, cfg)
    p = torch.fft.irfft2(torch.fft.rfft2(rhs) * inv_eig, s=rhs.shape[-2:])
    u2 = u - (_S(p, 1, -2) - p) / cfg.hx
    v2 = v - (_S(p, 1, -1) - p) / cfg.hy
    return u2, v2, p


def rk4_step(u: torch.Tensor, v: torch.Tensor, cfg: SolverConfig, dt: Optional[float] = None,
             inv_eig: Optional[torch.Tensor] = None
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """RKStepper.from_method('classic_rk4').forward(v, dt, equation) (src/inference.py:64,101):
    classic RK4 with a projection after every stage (App. A.2)."""
    if dt is None:
        dt = cfg.dt
    if inv_eig is None:
        inv_eig = inverse_eigenvalues(cfg, u.dtype, u.device)
    a = (0.5, 0.5, 1.0)
    b = (1.0 / 6, 1.0 / 3, 1.0 / 3, 1.0 / 6)
    us, vs = u, v
    ks = []
    for s in range(4):
        ku, kv = explicit_terms(us, vs, cfg, dt)
        ks.append((ku, kv))
        if s < 3:
            us, vs, _ = pressure_projection(u + dt * a[s] * ku, v + dt * a[s] * kv, cfg, inv_eig)
    du = sum(b[s] * ks[s][0] for s in range(4))
    dv = sum(b[s] * ks[s][1] for s in range(4))
    return pressure_projection(u + dt * du, v + dt * dv, cfg, inv_eig)


def filtered_velocity_field(cfg: SolverConfig, batch: int, seed: int = 42, peak_wavenumber: float = 4.0,
                            iterations: int = 16, dtype=torch.float64
                            ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Band-limited random divergence-free IC scaled to ``cfg.max_velocity`` (App. A.5; the role of
    torch_cfd's filtered_velocity_field at src/inference.py:76-78).  Setup only, not per-step."""
    g = torch.Generator().manual_seed(seed)
    kx = torch.fft.fftfreq(cfg.nx, d=1.0 / cfg.nx, dtype=torch.float64) * (2 * math.pi / cfg.lx)
    ky = torch.fft.rfftfreq(cfg.ny, d=1.0 / cfg.ny, dtype=torch.float64) * (2 * math.pi / cfg.ly)
    kk = torch.sqrt(kx[:, None] ** 2 + ky[None, :] ** 2)
    logvar = 0.25
    mu = math.log(peak_wavenumber) + logvar     # log-normal with mode at peak_wavenumber
    safe = torch.where(kk > 0, kk, torch.ones_like(kk))
    filt = torch.exp(-(torch.log(safe) - mu) ** 2 / (2 * logvar)) / (safe * math.sqrt(2 * math.pi * logvar)) / safe
    filt = torch.where(kk > 0, filt, torch.zeros_like(filt))
    comps = []
    for _ in range(2):
        noise = torch.randn(batch, cfg.nx, cfg.ny, generator=g, dtype=torch.float64)
        comps.append(torch.fft.irfft2(torch.fft.rfft2(noise) * filt, s=(cfg.nx, cfg.ny)))
    u, v = comps
    inv_eig = inverse_eigenvalues(cfg, torch.float64)
    for _ in range(iterations):
        u, v, _ = pressure_projection(u, v, cfg, inv_eig)
        speed = torch.sqrt(u * u + v * v).amax(dim=(-2, -1), keepdim=True)
        u = u * (cfg.max_velocity / speed)
        v = v * (cfg.max_velocity / speed)
    return u.to(dtype), v.to(dtype)
