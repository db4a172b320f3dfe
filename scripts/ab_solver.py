"""Solver-only A/B of library builds on the same box:
    python scripts/ab_solver.py <nx> <batch> <lib.so> [<lib.so> ...]
prints the solver step time (best of 3 x 20 steps) and the isolated K1 / K2 / K3 / K4 times of every build, twice."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, torch
sys.path.insert(0, %r)
from ml_accelerated_simulation_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
from ml_accelerated_simulation_b200 import Plan, PlanConfig
nx, batch = int(sys.argv[2]), int(sys.argv[3])
plan = Plan(PlanConfig(nx=nx, ny=nx, batch=batch))
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
u = torch.randn(batch, nx, nx, generator=g, device=dev); v = torch.randn(batch, nx, nx, generator=g, device=dev)
plan.project(u, v, want_p=False); plan.normalize_speed(u, v, 7.0)
ref = (u.clone(), v.clone())
plan.step(u, v, nsteps=5); torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.step(u, v, nsteps=20); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
ks = [plan.time_kernel(k, 3, 20) for k in (0, 1, 2, 3)]
cells = nx * nx * batch
chk = float(u.double().abs().sum())
print(f"{os.path.basename(sys.argv[1]):34s} {nx}^2 x {batch}: step {best:.4f} ms (frac {296 * cells / (best * 1e-3) / 6554.9e9:.3f}); K1 {ks[0]*1e3:.1f} K2 {ks[1]*1e3:.1f} K3 {ks[2]*1e3:.1f} K4 {ks[3]*1e3:.1f} us; checksum {chk:.6e}", flush=True)
''' % ROOT
nx, batch = sys.argv[1], sys.argv[2]
libs = sys.argv[3:]
for rep in range(2):
    for lib in libs:
        subprocess.run([sys.executable, "-c", CHILD, os.path.abspath(lib), nx, batch], check=False)
