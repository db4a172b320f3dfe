"""Solver-only step time at small batches (the blocks of the host pipeline) for the execution options."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_accelerated_simulation_b200 import Plan, PlanConfig
dev = torch.device("cuda:0")
for b in (9, 18, 37, 64):
    out = []
    for blocks in (1, 2, 4):
        for cl in (False, True):
            plan = Plan(PlanConfig(nx=256, ny=256, batch=b, solver_blocks=blocks, cluster_projection=cl))
            g = torch.Generator(device=dev).manual_seed(0)
            u = torch.randn(b, 256, 256, generator=g, device=dev); v = torch.randn(b, 256, 256, generator=g, device=dev)
            plan.project(u, v, want_p=False); plan.normalize_speed(u, v, 7.0)
            plan.step(u, v, nsteps=5); torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); plan.step(u, v, nsteps=20); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 20)
            out.append(f"blocks={blocks}{' cluster' if cl else ''}: {best:.4f}")
            plan.close()
    print(f"batch {b:3d}: " + " | ".join(out), flush=True)
