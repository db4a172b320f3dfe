"""Solver-only step time vs the number of stream blocks: python scripts/blocks_sweep.py nx batch"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_accelerated_simulation_b200 import Plan, PlanConfig
nx, b = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
out = []
for blocks in (1, 2, 4, 8):
    if blocks > b: continue
    plan = Plan(PlanConfig(nx=nx, ny=nx, batch=b, solver_blocks=blocks))
    g = torch.Generator(device=dev).manual_seed(0)
    u = torch.randn(b, nx, nx, generator=g, device=dev); v = torch.randn(b, nx, nx, generator=g, device=dev)
    plan.project(u, v, want_p=False); plan.normalize_speed(u, v, 7.0)
    plan.step(u, v, nsteps=5); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.step(u, v, nsteps=20); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    out.append(f"blocks={blocks}: {best:.4f} ms ({296 * nx * nx * b / (best * 1e-3) / 6554.9e9:.3f})")
    plan.close()
print(f"{nx}^2 x {b}: " + " | ".join(out), flush=True)
