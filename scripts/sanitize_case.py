"""Tiny end-to-end case for compute-sanitizer runs (64x128 grid, batch 2, solver + CNN, 2 steps)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import safetensors.torch as st
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfgj = json.load(open(os.path.join(ROOT, "tests/golden/MLACFD.json")))
sd = st.load_file(os.path.join(ROOT, "tests/golden/mlacfd_weights.safetensors"))
model = models.buildModel(cfgj); model.load_state_dict(sd)
plan = Plan(PlanConfig(nx=64, ny=128, batch=2)); model.attach(plan)
g = torch.Generator().manual_seed(0)
u = torch.randn(2, 64, 128, generator=g).cuda(); v = torch.randn(2, 64, 128, generator=g).cuda()
plan.project(u, v, want_p=False)
p = plan.new_field()
plan.step(u, v, nsteps=2, apply_cnn=True, p=p)
y = plan.cnn_forward(u, v, 1.0)
hu, hv = u.cpu(), v.cpu()
plan.rollout_host(hu, hv, 1, apply_cnn=True)
torch.cuda.synchronize()
print("ok", bool(torch.isfinite(u).all()), float(y.abs().max()))
