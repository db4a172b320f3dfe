"""Small driver for ncu captures: a few steps of the hot path at BASELINE config 2 (or argv: nx ny batch cnn)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models

nx, ny, batch, cnn, blocks = (int(x) for x in (sys.argv[1:6] + ["256", "256", "64", "0", "0"][len(sys.argv) - 1:]))
plan = Plan(PlanConfig(nx=nx, ny=ny, batch=batch, solver_blocks=blocks))
if cnn:
    cfgj = json.load(open(os.path.join(ROOT, "tests/golden/MLACFD.json")))
    model = models.buildModel(cfgj); model.attach(plan)
g = torch.Generator(device="cuda").manual_seed(0)
u = torch.randn(batch, nx, ny, device="cuda", generator=g); v = torch.randn(batch, nx, ny, device="cuda", generator=g)
plan.project(u, v, want_p=False)
u *= 3; v *= 3
plan.step(u, v, nsteps=3, apply_cnn=bool(cnn))
torch.cuda.synchronize()
print("ok", float(u.abs().max()))
