"""GPU parity of the long-grid / slab-decomposed paths, through the C ABI:

  * plain plans with nx or ny > 2048 (4-stage row transforms, split x-transform) against the oracle;
  * slab mode at world size 1 on one GPU (the same kernels and layouts as the multi-GPU run, exchanges are local
    copies) against the oracle: solver, projection, solver + CNN;
  * slab mode at world size 2 / 4 over NCCL (one process per GPU, torchrun) when the box has enough GPUs: the ranks'
    gathered result must equal the single-GPU plain plan on the same input (fp32 round-off) and the oracle.

Tolerances: the solver protocol of test_gpu_parity.py (relL2(K32,O64) <= 4*relL2(O32,O64) + 2e-7).
"""
import json
import os
import subprocess
import sys

import pytest
import safetensors.torch as st
import torch

from oracle import cnn as OC
from oracle import solver as S

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module")
def mlsim(built_lib):
    import ml_accelerated_simulation_b200 as m
    return m


@pytest.fixture(scope="module")
def mlacfd(golden_dir):
    cfg = json.load(open(os.path.join(golden_dir, "MLACFD.json")))
    sd = st.load_file(os.path.join(golden_dir, "mlacfd_weights.safetensors"))
    return cfg, sd


# ------------------------------------------------------------------------------------------ long plain grids
@pytest.mark.parametrize("nx,ny", [(4096, 64), (64, 4096), (8192, 64), (64, 8192), (4096, 128)])
def test_projection_long_grids(mlsim, nx, ny):
    cfg = S.SolverConfig(nx=nx, ny=ny)
    g = torch.Generator().manual_seed(nx + ny)
    u64 = torch.randn(1, nx, ny, generator=g, dtype=torch.float64)
    v64 = torch.randn(1, nx, ny, generator=g, dtype=torch.float64)
    plan = mlsim.Plan(mlsim.PlanConfig(nx=nx, ny=ny, batch=1))
    u, v = u64.float().to(DEV), v64.float().to(DEV)
    p = plan.project(u, v)
    ou, ov, op = S.pressure_projection(u64, v64, cfg)
    qu, qv, qp = S.pressure_projection(u64.float(), v64.float(), cfg)
    assert rel(u, ou) <= 4 * rel(qu, ou) + 2e-7
    assert rel(v, ov) <= 4 * rel(qv, ov) + 2e-7
    assert rel(p, op) <= 4 * rel(qp, op) + 2e-6
    plan.close()


def test_step_long_grid(mlsim):
    cfg = S.SolverConfig(nx=4096, ny=64)
    u64, v64 = S.filtered_velocity_field(cfg, 1, seed=3, iterations=4)
    plan = mlsim.Plan(mlsim.PlanConfig(nx=4096, ny=64, batch=1))
    u, v = u64.float().to(DEV), v64.float().to(DEV)
    plan.step(u, v, nsteps=2)
    ou, ov, _ = OC.rollout(u64, v64, cfg, 2)
    qu, qv, _ = OC.rollout(u64.float(), v64.float(), cfg, 2)
    assert rel(u, ou) <= 4 * rel(qu, ou) + 2e-7 and rel(v, ov) <= 4 * rel(qv, ov) + 2e-7
    plan.close()


def test_projection_8192_squared_properties(mlsim):
    """BASELINE config 5 size on one GPU: size-independent properties (divergence-free, idempotent)."""
    nx = ny = 8192
    plan = mlsim.Plan(mlsim.PlanConfig(nx=nx, ny=ny, batch=1))
    g = torch.Generator(device=DEV).manual_seed(1)
    u = torch.randn(1, nx, ny, generator=g, device=DEV)
    v = torch.randn(1, nx, ny, generator=g, device=DEV)
    plan.project(u, v, want_p=False)
    d = plan.divergence(u, v)
    h = 2 * torch.pi / nx
    vmax = max(u.abs().max().item(), v.abs().max().item())
    assert d.abs().max().item() * h / vmax <= 2e-5
    u2, v2 = u.clone(), v.clone()
    plan.project(u2, v2, want_p=False)
    assert rel(u2, u) < 1e-5 and rel(v2, v) < 1e-5
    plan.close()


# ------------------------------------------------------------------------------------------ slab mode, one GPU
@pytest.mark.parametrize("nx,ny,batch", [(64, 64, 2), (256, 128, 1), (4096, 64, 1)])
def test_slab_world1_matches_oracle(mlsim, mlacfd, nx, ny, batch):
    from ml_accelerated_simulation_b200.slab import CudaSlabBackend, SlabGeometry, SlabStepper
    cfgj, sd = mlacfd
    cfg = S.SolverConfig(nx=nx, ny=ny)
    u64, v64 = S.filtered_velocity_field(cfg, batch, seed=11, iterations=4)
    be = CudaSlabBackend(SlabGeometry(nx=nx, ny=ny, batch=batch, rank=0, world=1))
    from ml_accelerated_simulation_b200 import models
    model = models.buildModel(cfgj)
    model.load_state_dict(sd)
    be.set_cnn(model.cnn.folded_layers())
    stp = SlabStepper(be)
    # projection alone
    g = torch.Generator().manual_seed(5)
    ru = torch.randn(batch, nx, ny, generator=g, dtype=torch.float64)
    rv = torch.randn(batch, nx, ny, generator=g, dtype=torch.float64)
    stp.scatter(ru.float().to(DEV), rv.float().to(DEV))
    p = stp.project(want_p=True)
    ou, ov, op = S.pressure_projection(ru, rv, cfg)
    qu, qv, qp = S.pressure_projection(ru.float(), rv.float(), cfg)
    u, v = stp.gather()
    assert rel(u, ou) <= 4 * rel(qu, ou) + 2e-7 and rel(v, ov) <= 4 * rel(qv, ov) + 2e-7
    assert rel(p, op) <= 4 * rel(qp, op) + 2e-6
    # solver steps
    stp.scatter(u64.float().to(DEV), v64.float().to(DEV))
    stp.step(3)
    u, v = stp.gather()
    ou, ov, _ = OC.rollout(u64, v64, cfg, 3)
    qu, qv, _ = OC.rollout(u64.float(), v64.float(), cfg, 3)
    assert rel(u, ou) <= 4 * rel(qu, ou) + 2e-7 and rel(v, ov) <= 4 * rel(qv, ov) + 2e-7
    # solver + CNN (bf16 storage error dominates: compare with the bf16-emulating oracle)
    if nx * ny <= 256 * 128:
        stp.scatter(u64.float().to(DEV), v64.float().to(DEV))
        stp.step(3, apply_cnn=True)
        u, v = stp.gather()
        eu, ev, _ = OC.rollout(u64, v64, cfg, 3, cfgj, sd, 1.0, round_bf16=True)
        assert rel(u, eu) <= 4 * rel(qu, ou) + 1e-6 and rel(v, ev) <= 4 * rel(qv, ov) + 1e-6
    be.close()


# ------------------------------------------------------------------------------------------ slab mode over NCCL
@pytest.mark.parametrize("world", [2, 4])
def test_slab_nccl_matches_single_gpu(built_lib, world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = str(tmp_path / "slab.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "slab_worker.py"),
           "--nx", "256", "--ny", "128", "--steps", "3", "--out", out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.load(open(out))
    assert res["solver_vs_oracle_u"] <= 4 * res["o32_vs_oracle_u"] + 2e-7, res
    assert res["solver_vs_oracle_v"] <= 4 * res["o32_vs_oracle_v"] + 2e-7, res
    assert res["cnn_vs_bf16_oracle_u"] <= 4 * res["o32_vs_oracle_u"] + 1e-6, res
    assert res["cnn_vs_bf16_oracle_v"] <= 4 * res["o32_vs_oracle_v"] + 1e-6, res
    assert res["vs_plain_plan_u"] <= 1e-6 and res["vs_plain_plan_v"] <= 1e-6, res
