"""Per-class conv kernel times in the step for several builds: python scripts/ab_cnn.py <nx> <batch> <lib.so> ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from ml_accelerated_simulation_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
nx, batch, steps = int(sys.argv[2]), int(sys.argv[3]), 20
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
plan = Plan(PlanConfig(nx=nx, ny=nx, batch=batch)); model.attach(plan)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
u = torch.randn(batch, nx, nx, generator=g, device=dev); v = torch.randn(batch, nx, nx, generator=g, device=dev)
plan.project(u, v, want_p=False); plan.normalize_speed(u, v, 7.0)
plan.step(u, v, nsteps=5, apply_cnn=True); torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.step(u, v, nsteps=steps, apply_cnn=True); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / steps)
cls = {}
for c, nm in ((8, "fused"), (5, "mid"), (6, "last")):
    plan.profile_enable(c, 8 * steps); plan.step(u, v, nsteps=steps, apply_cnn=True); ms, n = plan.profile_read()
    cls[nm] = round(ms / max(n, 1), 4)
print(f"{os.path.basename(sys.argv[1]):28s} {nx}^2 x {batch}: step {best:.4f} ms; per launch: {cls}; checksum {float(u.double().abs().sum()):.6e}", flush=True)
''' % ROOT
nx, batch = sys.argv[1], sys.argv[2]
for rep in range(2):
    for lib in sys.argv[3:]:
        subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(lib), nx, batch], check=False)
