"""The drop-in route on the GPU: the reference's loop written with the reference's import names (compat/ aliases).

``_sim_test_body`` is the simulation section of reference test/sim_test.py:18-90.  When the reference tree is
available (``MLSIM_REFERENCE_ROOT`` or /root/reference) the lines are read from THERE and executed unmodified;
on the GPU box it is not, so the same call sequence is written out below (the reference has no other content in that
section: constants, the torch_cfd calls of SURVEY.md 8b, and the NaN-guarded double loop)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


@pytest.fixture()
def compat_path(built_lib):
    saved = list(sys.path)
    sys.path[:0] = [os.path.join(ROOT, "compat"), ROOT]
    prev = torch.get_default_dtype()
    yield
    torch.set_default_dtype(prev)
    sys.path[:] = saved


def _run_sim_test_body():
    ref = os.path.join(os.environ.get("MLSIM_REFERENCE_ROOT", "/root/reference"), "test", "sim_test.py")
    ns = {"__name__": "sim_test_body"}
    if os.path.exists(ref):
        lines = open(ref).read().splitlines()
        head = [ln for ln in lines[:9] if "torch" in ln]                 # :1-8 torch / torch_cfd imports
        body = lines[17:90]                                              # :18-90, unmodified
        src = "\n".join(head + ["from tqdm import tqdm"] + body)
        cwd = os.getcwd()
        os.chdir("/tmp")                                                 # the script touches ./test.txt (:33)
        try:
            exec(compile(src, ref, "exec"), ns)
        finally:
            os.chdir(cwd)
        return ns, "reference file"
    from torch_cfd import grids, boundaries
    from torch_cfd.initial_conditions import filtered_velocity_field
    from torch_cfd.equations import stable_time_step
    from torch_cfd.fvm import RKStepper, NavierStokes2DFVMProjection
    from torch_cfd.forcings import KolmogorovForcing
    import torch_cfd.finite_differences as fdm
    n, batch_size, density, max_velocity, peak_wavenumber, cfl, viscosity = 128, 1, 1.0, 3.0, 4.0, 0.5, 1e-3
    device = torch.device(DEV)
    torch.set_default_dtype(torch.float64)
    inner_steps, outer_steps, diam = 20, 100, 2 * torch.pi
    grid = grids.Grid((n, n), domain=((0, diam), (0, diam)), device=device)
    v0 = filtered_velocity_field(grid, max_velocity, peak_wavenumber, iterations=3, random_state=41, device=device,
                                 batch_size=batch_size)
    v0div = fdm.divergence(v0)
    boundaries.get_pressure_bc_from_velocity(v0)
    div_norm = float(torch.linalg.norm(v0div).data)
    dt = stable_time_step(dx=min(grid.step), max_velocity=max_velocity, max_courant_number=cfl, viscosity=viscosity)
    step_fn = RKStepper.from_method(method="classic_rk4", requires_grad=False, dtype=torch.float64)
    forcing_fn = KolmogorovForcing(diam=diam, wave_number=int(peak_wavenumber), grid=grid,
                                   offsets=(v0[0].offset, v0[1].offset))
    ns2d = NavierStokes2DFVMProjection(viscosity=viscosity, grid=grid, bcs=(v0[0].bc, v0[1].bc), density=density,
                                       drag=0.1, forcing=forcing_fn, solver=step_fn).to(v0.device)
    v = v0
    trajectory = [[v0[0].data.detach().cpu().numpy()], [v0[1].data.detach().cpu().numpy()]]
    nan_count = 0
    with torch.no_grad():
        for i in range(outer_steps):
            for j in range(inner_steps):
                v, p = step_fn.forward(v, dt, equation=ns2d)
                if torch.isnan(v[0].data).any():
                    nan_count += 1
                    break
            if nan_count > 0:
                break
            trajectory[0].append(v[0].data.detach().cpu().numpy())
            trajectory[1].append(v[1].data.detach().cpu().numpy())
    ns.update(dict(v=v, v0=v0, nan_count=nan_count, trajectory=trajectory, dt=dt, n=n, v0div=v0div, div_norm=div_norm,
                   ns2d=ns2d, outer_steps=outer_steps, inner_steps=inner_steps))
    return ns, "transcription"


def test_sim_test_body_runs_through_the_aliases(compat_path):
    """reference test/sim_test.py:18-90: 2000 RK4 steps at 128 x 128 through RKStepper.forward, NaN guard as the
    reference has it, and the fused rollout (Plan.step, 2000 steps in one call) gives the same final state."""
    ns, how = _run_sim_test_body()
    v, v0 = ns["v"], ns["v0"]
    assert ns["nan_count"] == 0 and len(ns["trajectory"][0]) == 101, how
    assert v[0].data.dtype == torch.float64 and v[0].data.shape == (1, 128, 128)
    assert bool(torch.isfinite(v[0].data).all()) and bool(torch.isfinite(v[1].data).all())
    assert 0.5 < float(v[0].data.abs().max()) < 10.0                      # SURVEY App. B probe: max|u| = 2.40 after 2000 steps
    from ml_accelerated_simulation_b200 import Plan, PlanConfig
    plan = Plan(PlanConfig(nx=128, ny=128, batch=1, dt=float(ns["dt"]), max_velocity=3.0))
    u, w = v0[0].data.float().contiguous(), v0[1].data.float().contiguous()
    plan.step(u, w, nsteps=2000)
    assert torch.equal(u.double(), v[0].data) and torch.equal(w.double(), v[1].data)      # same kernels, same order
    plan.close()


def test_step_f64_equals_step(built_lib):
    from ml_accelerated_simulation_b200 import Plan, PlanConfig
    plan = Plan(PlanConfig(nx=64, ny=128, batch=3))
    g = torch.Generator(device=DEV).manual_seed(0)
    u = torch.randn(3, 64, 128, generator=g, device=DEV)
    v = torch.randn(3, 64, 128, generator=g, device=DEV)
    plan.project(u, v, want_p=False)
    u64, v64 = u.double(), v.double()
    p32 = plan.new_field()
    plan.step(u, v, nsteps=2, p=p32)
    uo, vo, po = plan.step_f64(u64, v64, nsteps=2)
    assert uo.dtype == torch.float64 and torch.equal(uo, u.double()) and torch.equal(vo, v.double()) and torch.equal(po, p32.double())
    plan.step_f64(u64, v64, nsteps=2, want_p=False, inplace=True)
    assert torch.equal(u64, uo) and torch.equal(v64, vo)
    with pytest.raises(ValueError):
        plan.step_f64(u, v)
    plan.close()


def test_two_plans_on_two_devices_in_one_process(built_lib):
    """One plan == one device, but one process may own several plans (mlsim_config.device): the per-device kernel
    attributes (dynamic shared memory) and the current-device switch of every entry point must follow the plan."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import json
    import safetensors.torch as st
    from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
    g = os.path.join(ROOT, "tests", "golden")
    cfgj = json.load(open(os.path.join(g, "MLACFD.json")))
    sd = st.load_file(os.path.join(g, "mlacfd_weights.safetensors"))
    model = models.buildModel(cfgj)
    model.load_state_dict(sd)
    gen = torch.Generator().manual_seed(3)
    u0 = torch.randn(2, 128, 128, generator=gen)
    v0 = torch.randn(2, 128, 128, generator=gen)
    outs = []
    plans = [Plan(PlanConfig(nx=128, ny=128, batch=2, device=d)) for d in (0, 1)]       # both alive at once
    for d, plan in enumerate(plans):
        model.attach(plan)
    torch.cuda.set_device(0)                                                             # current device != plan 1's device
    for d, plan in enumerate(plans):
        u, v = u0.to(f"cuda:{d}"), v0.to(f"cuda:{d}")
        plan.project(u, v, want_p=False)
        plan.step(u, v, nsteps=2, apply_cnn=True)
        hu, hv = u0.clone(), v0.clone()
        plan.rollout_host(hu, hv, 0)
        torch.cuda.synchronize(d)
        outs.append((u.cpu(), v.cpu()))
    assert torch.isfinite(outs[0][0]).all()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    with pytest.raises(ValueError):
        plans[1].step(u0.to("cuda:0"), v0.to("cuda:0"))                                  # a field on the wrong device is refused
    for plan in plans:
        plan.close()
