"""A/B of two builds of the library on the same box: python scripts/ab_step.py <lib_a.so> <lib_b.so> [nx batch steps]
Alternates the two builds (A B A B) and prints the step time and the mid-conv launch time of each."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from ml_accelerated_simulation_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
nx, batch, steps = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
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
plan.profile_enable(5, 8 * steps); plan.step(u, v, nsteps=steps, apply_cnn=True); ms, n = plan.profile_read()
k5 = plan.time_kernel(5, 3, 20)
print(f"{os.path.basename(sys.argv[1])}: step {best:.4f} ms (best of 3 x {steps}), mid conv in-step {ms/max(n,1):.4f} ms, isolated {k5:.4f} ms", flush=True)
''' % ROOT
a, b = sys.argv[1], sys.argv[2]
nx, batch, steps = (sys.argv[3:6] + ["256", "64", "20"][len(sys.argv[3:6]):])
for lib in (a, b, a, b):
    subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(lib), nx, batch, steps], check=False)
