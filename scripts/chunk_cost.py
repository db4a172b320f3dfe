"""Device time of one solver+CNN step as a function of the batch (what the host pipeline's blocks cost)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
dev = torch.device("cuda:0")
for b in (4, 9, 18, 27, 28, 32, 37, 46, 55, 64):
    plan = Plan(PlanConfig(nx=256, ny=256, batch=b)); model.attach(plan)
    g = torch.Generator(device=dev).manual_seed(0)
    u = torch.randn(b, 256, 256, generator=g, device=dev); v = torch.randn(b, 256, 256, generator=g, device=dev)
    plan.project(u, v, want_p=False); plan.normalize_speed(u, v, 7.0)
    plan.step(u, v, nsteps=5, apply_cnn=True); torch.cuda.synchronize()
    res = []
    for cnn in (True, False):
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.step(u, v, nsteps=20, apply_cnn=cnn); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 20)
        res.append(best)
    print(f"batch {b:3d}: step {res[0]:.4f} ms ({res[0]/b*1e3:.1f} us/traj), solver only {res[1]:.4f} ms", flush=True)
    plan.close()
