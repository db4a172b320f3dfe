"""Sweep of the host-pipeline block sizes of mlsim_rollout_host at config 2: python scripts/e2e_blocks.py 9,46,9 9,9,37,9 ..."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
plan = Plan(PlanConfig(nx=256, ny=256, batch=64)); model.attach(plan)
u0, v0 = bench._device_state(64, 42, torch.device("cuda:0"))
hu, hv = u0.cpu().pin_memory(), v0.cpu().pin_memory()
for rep in range(2):
    for ch in ["auto"] + sys.argv[1:]:
        plan.set_host_blocks(None if ch == "auto" else [int(x) for x in ch.split(",")])
        hu.copy_(u0); hv.copy_(v0)
        plan.rollout_host(hu, hv, 1, apply_cnn=True); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(20): plan.rollout_host(hu, hv, 1, apply_cnn=True)
        dt = (time.perf_counter() - t) / 20
        print(f"blocks={ch}: {dt*1e3:.3f} ms/step -> {64*256*256/dt/1e9:.3f} Gcell-steps/s", flush=True)
