import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
plan = Plan(PlanConfig(nx=256, ny=256, batch=64)); model.attach(plan)
for which in (5, 6):
    ms = plan.time_kernel(which, 3, 10)
    n, d = plan.conv_debug()
    rows = d[4]
    print(f"kernel {which}: {ms:.4f} ms, {n} CTAs, rows/CTA {rows:.1f}")
    print(f"  MMA warp: total {d[0]:.0f} cyc = {d[0]/rows:.0f}/row | wait tempty {d[1]/rows:.0f}/row | wait full {d[2]/rows:.0f}/row | issue {d[3]/rows:.0f}/row | other {(d[0]-d[1]-d[2]-d[3])/rows:.0f}/row")
    print(f"  epilogue warp: total {d[5]:.0f} cyc | waiting tfull {d[6]:.0f} ({100*d[6]/max(d[5],1):.0f}%)")
