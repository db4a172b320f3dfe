"""Known-answer tests pinning the solver oracle.  The oracle's parity with
torch_cfd itself is UNPINNED (package absent, see oracle/solver.py header); these tests pin its physics."""
import math

import pytest
import torch

from oracle import solver as S


def test_stable_time_step_matches_reference_constants():
    # src/inference.py:68-73 -> 0.5*dx/7 for the 64^2 grid; src/data_generation.py:135-147 for 1024^2
    assert S.SolverConfig(nx=64, ny=64).dt == pytest.approx(0.5 * (2 * math.pi / 64) / 7.0, rel=1e-15)
    assert S.SolverConfig(nx=64, ny=64).dt == pytest.approx(7.0125e-3, rel=1e-4)
    assert S.SolverConfig(nx=1024, ny=1024).dt == pytest.approx(4.383e-4, rel=1e-3)
    # diffusion-limited branch
    assert S.stable_time_step(0.1, 1.0, 0.5, 10.0) == pytest.approx(0.1 * 0.1 / 40.0)


def test_projection_invariants():
    cfg = S.SolverConfig()
    g = torch.Generator().manual_seed(0)
    u = torch.randn(2, 64, 64, generator=g, dtype=torch.float64)
    v = torch.randn(2, 64, 64, generator=g, dtype=torch.float64)
    u2, v2, p = S.pressure_projection(u, v, cfg)
    assert S.divergence(u2, v2, cfg).abs().max() < 1e-12
    u3, v3, _ = S.pressure_projection(u2, v2, cfg)
    assert (u3 - u2).abs().max() < 1e-13 and (v3 - v2).abs().max() < 1e-13
    assert abs(p.mean().item()) < 1e-15


def test_kolmogorov_laminar_fixed_point_is_exact():
    cfg = S.SolverConfig()
    h, kf = cfg.hx, cfg.forcing_wavenumber
    A = 1.0 / (cfg.viscosity * (4 / h ** 2) * math.sin(kf * h / 2) ** 2 + cfg.drag)
    assert A == pytest.approx(8.63591885296836, rel=1e-13)
    u0 = (A * S.forcing_profile(cfg))[None, None, :].expand(1, 64, 64).contiguous()
    v0 = torch.zeros_like(u0)
    u1, v1, _ = S.rk4_step(u0, v0, cfg)
    assert (u1 - u0).abs().max().item() < 1e-13
    assert v1.abs().max().item() < 1e-13


def _taylor_green_error(n, advection, lw):
    cfg = S.SolverConfig(nx=n, ny=n, viscosity=0.01, drag=0.0, forcing_scale=0.0, advection=advection,
                         lw_dt_scale=lw)
    h = cfg.hx
    i = torch.arange(n, dtype=torch.float64)
    u = (torch.cos((i + 1) * h)[:, None] * torch.sin((i + 0.5) * h)[None, :])[None]
    v = (-torch.sin((i + 0.5) * h)[:, None] * torch.cos((i + 1) * h)[None, :])[None]
    dt = min(0.5 * h, h * h / (4 * 0.01))
    steps = math.ceil(1 / dt)
    dt = 1 / steps
    ue, ve = u * math.exp(-0.02), v * math.exp(-0.02)
    for _ in range(steps):
        u, v, _ = S.rk4_step(u, v, cfg, dt)
    return max((u - ue).abs().max().item(), (v - ve).abs().max().item())


@pytest.mark.parametrize("n,expected", [(32, 2.214e-2), (64, 1.236e-2)])
def test_taylor_green_van_leer(n, expected):
    assert _taylor_green_error(n, "van_leer", 1.0) == pytest.approx(expected, rel=2e-3)


def test_taylor_green_limiters():
    assert _taylor_green_error(64, "lax_wendroff", 1.0) == pytest.approx(1.223e-2, rel=2e-3)
    assert _taylor_green_error(64, "upwind", 1.0) == pytest.approx(3.544e-2, rel=2e-3)


def test_second_order_convergence_without_lw_time_term():
    e32 = _taylor_green_error(32, "lax_wendroff", 0.0)
    e64 = _taylor_green_error(64, "lax_wendroff", 0.0)
    assert e32 == pytest.approx(6.260e-5, rel=2e-3)
    assert e32 / e64 == pytest.approx(4.0, rel=0.02)


def test_sim_test_smoke_property_stays_finite():
    """reference test/sim_test.py:18-85 (128^2, vmax 3, iterations 3, seed 41): only asserts no NaN.
    200 of its 2000 steps keep the CPU suite short."""
    cfg = S.SolverConfig(nx=128, ny=128, max_velocity=3.0)
    u, v = S.filtered_velocity_field(cfg, 1, seed=41, iterations=3)
    assert S.divergence(u, v, cfg).norm().item() < 1e-9            # test/sim_test.py:42-45
    inv = S.inverse_eigenvalues(cfg)
    for _ in range(200):
        u, v, _ = S.rk4_step(u, v, cfg, None, inv)
    assert torch.isfinite(u).all() and torch.isfinite(v).all()
    assert u.abs().max().item() < 10.0


def test_initial_condition_contract():
    cfg = S.SolverConfig()
    u, v = S.filtered_velocity_field(cfg, 3, seed=42)
    speed = torch.sqrt(u * u + v * v).amax(dim=(-2, -1))
    assert torch.allclose(speed, torch.full_like(speed, 7.0), rtol=1e-12)
    assert S.divergence(u, v, cfg).abs().max().item() < 1e-11


def test_matches_torch_cfd():
    """Consumes tests/golden/torchcfd_golden.safetensors when tests/golden/make_golden_torchcfd.py could produce it
    (it needs the real torch_cfd package; see DESIGN.md section 2 for the dated outcome of the attempt).  Until then the
    solver oracle is PARITY UNPINNED and this test skips -- it must never silently pass."""
    import os
    import safetensors.torch as st
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "torchcfd_golden.safetensors")
    if not os.path.exists(path):
        pytest.skip("no torch_cfd fixture (package unobtainable: parity unpinned)")
    g = st.load_file(path)
    cfg = S.SolverConfig()
    u0, v0 = g["v0_u"], g["v0_v"]
    dt = float(g["dt"][0])
    assert dt == pytest.approx(cfg.dt, rel=1e-15)

    def close(a, b):
        return ((a - b).norm() / b.norm()).item() <= 1e-12
    if "F_u" in g:
        fu, fv = S.explicit_terms(u0, v0, cfg, dt)
        assert close(fu, g["F_u"]) and close(fv, g["F_v"])
    if "P_u" in g:
        pu, pv, pp = S.pressure_projection(u0, v0, cfg)
        assert close(pu, g["P_u"]) and close(pv, g["P_v"])
    u, v, p = S.rk4_step(u0, v0, cfg, dt)
    assert close(u, g["step1_u"]) and close(v, g["step1_v"])
    for _ in range(9):
        u, v, p = S.rk4_step(u, v, cfg, dt)
    assert close(u, g["step10_u"]) and close(v, g["step10_v"])
