"""Config 1 (64 x 64, batch 1) and other latency-bound shapes: three-kernel projection vs the single cluster kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
dev = torch.device("cuda:0")
for n, b in ((64, 1), (64, 4), (64, 16), (128, 1), (128, 4), (256, 1), (256, 2)):
    out = []
    for cl in (False, True):
        plan = Plan(PlanConfig(nx=n, ny=n, batch=b, cluster_projection=cl)); model.attach(plan)
        g = torch.Generator(device=dev).manual_seed(0)
        u = torch.randn(b, n, n, generator=g, device=dev); v = torch.randn(b, n, n, generator=g, device=dev)
        plan.project(u, v, want_p=False); plan.normalize_speed(u, v, 7.0)
        plan.step(u, v, nsteps=20, apply_cnn=True); torch.cuda.synchronize()
        r = []
        for cnn in (True, False):
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); plan.step(u, v, nsteps=100, apply_cnn=cnn); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 100)
            r.append(best)
        out.append(f"{'cluster' if cl else '3-kernel'}: step {r[0]:.4f} solver {r[1]:.4f}")
        plan.close()
    print(f"{n}^2 x {b}: " + " | ".join(out), flush=True)
