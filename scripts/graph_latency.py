"""Step latency of the launch-bound shapes with and without the captured step graph (mlsim_config.step_graph).
    python scripts/graph_latency.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models

dev = torch.device("cuda:0")
cfgj = json.load(open(os.path.join(ROOT, "tests/golden/MLACFD.json")))
for nx, batch in ((64, 1), (64, 16), (128, 4), (256, 4), (256, 64)):
    for cnn in (True, False):
        row = []
        for sg in (False, True):
            plan = Plan(PlanConfig(nx=nx, ny=nx, batch=batch, step_graph=sg))
            torch.manual_seed(1); model = models.buildModel(cfgj); model.attach(plan)
            g = torch.Generator(device=dev).manual_seed(0)
            u = torch.randn(batch, nx, nx, generator=g, device=dev); v = torch.randn(batch, nx, nx, generator=g, device=dev)
            plan.project(u, v, want_p=False); plan.normalize_speed(u, v, 7.0)
            for _ in range(3):
                plan.step(u, v, nsteps=16, apply_cnn=cnn)
            best = 1e9
            for rep in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(); plan.step(u, v, nsteps=96, apply_cnn=cnn); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 96)
            row.append((best, plan.step_graph_launches(), float(u.double().abs().sum())))
            plan.close()
        print(f"{nx}^2 x {batch} cnn={cnn}: eager {row[0][0]*1e3:.1f} us/step, graph {row[1][0]*1e3:.1f} us/step "
              f"({row[1][1]} graph launches), same result: {row[0][2] == row[1][2]}", flush=True)
