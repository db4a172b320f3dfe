import os, sys, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
plan = Plan(PlanConfig(nx=256, ny=256, batch=64)); model.attach(plan)
u0, v0 = bench._device_state(64, 42, torch.device("cuda:0"))
hu, hv = u0.cpu().pin_memory(), v0.cpu().pin_memory()
for ch in sys.argv[1:] or ["default"]:
    os.environ.pop("MLSIM_HOST_CHUNK_SIZES", None)
    if ch != "default":
        os.environ["MLSIM_HOST_CHUNK_SIZES"] = ch
    plan.rollout_host(hu, hv, 1, apply_cnn=True); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10): plan.rollout_host(hu, hv, 1, apply_cnn=True)
    dt = (time.perf_counter() - t) / 10
    print(f"chunks={ch}: {dt*1e3:.3f} ms/step -> {64*256*256/dt/1e9:.3f} Gcell-steps/s")
