"""Solver-only step time at BASELINE config 2 (or argv nx ny batch), 300 steps, for tuning sweeps over env hooks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ml_accelerated_simulation_b200 import Plan, PlanConfig
nx, ny, batch = (int(x) for x in (sys.argv[1:4] + ["256", "256", "64"][len(sys.argv) - 1:]))
plan = Plan(PlanConfig(nx=nx, ny=ny, batch=batch))
g = torch.Generator(device="cuda").manual_seed(0)
u = torch.randn(batch, nx, ny, device="cuda", generator=g); v = torch.randn(batch, nx, ny, device="cuda", generator=g)
plan.project(u, v, want_p=False)
u *= 3; v *= 3
plan.step(u, v, nsteps=20)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); plan.step(u, v, nsteps=100); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 100)
cells = nx * ny * batch
print(f"{nx}x{ny} b{batch} env={ {k: v for k, v in os.environ.items() if k.startswith('MLSIM_')} }: {best:.4f} ms/step -> frac {312 * cells / (best * 1e-3) / 6554.9e9:.3f}")
