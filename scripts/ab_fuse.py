"""Fused first layer on / off on the same box: python scripts/ab_fuse.py [nx batch steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
nx, batch, steps = (int(x) for x in (sys.argv[1:4] + ["256", "64", "20"][len(sys.argv[1:4]):]))
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
u0 = torch.randn(batch, nx, nx, generator=g, device=dev); v0 = torch.randn(batch, nx, nx, generator=g, device=dev)
for rep in range(2):
    for fuse in (False, True):
        plan = Plan(PlanConfig(nx=nx, ny=nx, batch=batch, fuse_first_layer=fuse)); model.attach(plan)
        u, v = u0.clone(), v0.clone()
        plan.project(u, v, want_p=False); plan.normalize_speed(u, v, 7.0)
        plan.step(u, v, nsteps=5, apply_cnn=True); torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.step(u, v, nsteps=steps, apply_cnn=True); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps)
        cls = {}
        for c, nm in ((4, "first"), (5, "mid"), (6, "last")):
            plan.profile_enable(c, 8 * steps); plan.step(u, v, nsteps=steps, apply_cnn=True); ms, n = plan.profile_read()
            cls[nm] = (round(ms / steps, 4), n // steps)
        plan.profile_enable(-1, 0)
        print(f"{nx}^2 x {batch} fuse_first_layer={int(fuse)}: step {best:.4f} ms; conv ms/step (launches): {cls}", flush=True)
        plan.close()
