"""Why does bench.py's solver-only pass read slower than scripts/ab_solver.py?  Repeats the pass in the bench's context
(plan with the network attached, right after CNN-heavy stepping) and prints device time and host enqueue time per step.
    python scripts/solver_pass_probe.py [steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models

K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for blocks in (0, 1, 2, 4):
    plan = Plan(PlanConfig(nx=256, ny=256, batch=64, solver_blocks=blocks))
    model = models.buildModel(json.load(open(os.path.join(ROOT, "tests/golden/MLACFD.json"))))
    model.attach(plan)
    u = torch.randn(64, 256, 256, generator=g, device=dev); v = torch.randn(64, 256, 256, generator=g, device=dev)
    plan.project(u, v, want_p=False); plan.normalize_speed(u, v, 7.0)
    for heavy in (False, True):
        if heavy:
            plan.step(u, v, nsteps=60, apply_cnn=True)
        plan.step(u, v, nsteps=3, apply_cnn=False)
        out = []
        for rep in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter(); e0.record(); plan.step(u, v, nsteps=K, apply_cnn=False); e1.record()
            t1 = time.perf_counter(); torch.cuda.synchronize()
            out.append((round(e0.elapsed_time(e1) / K, 4), round((t1 - t0) * 1e3 / K, 4)))
        print(f"solver_blocks={blocks} after_cnn={heavy} launches/step={plan.launches_per_step(False)}: (device ms, host enqueue ms) per step:", out, flush=True)
    plan.close()
