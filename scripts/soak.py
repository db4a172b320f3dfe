"""Soak / determinism check of the rollout path: the same long rollout twice must give bit-identical states
(a race in the mbarrier / TMEM protocols of the conv kernels or in the stream forks of the solver would show up
as a difference, a protocol hang as a trapped launch)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
nx, batch, steps = (int(x) for x in (sys.argv[1:4] + ["256", "64", "4000"][len(sys.argv[1:4]):]))
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
dev = torch.device("cuda:0")
plan = Plan(PlanConfig(nx=nx, ny=nx, batch=batch)); model.attach(plan)
g = torch.Generator(device=dev).manual_seed(0)
u0 = torch.randn(batch, nx, nx, generator=g, device=dev); v0 = torch.randn(batch, nx, nx, generator=g, device=dev)
plan.project(u0, v0, want_p=False); plan.normalize_speed(u0, v0, 7.0)
outs = []
for rep in range(2):
    u, v = u0.clone(), v0.clone()
    plan.finite_check_enable(500)
    t = time.time()
    for _ in range(steps // 500):
        plan.step(u, v, nsteps=500, apply_cnn=True)
    torch.cuda.synchronize()
    bad, done = plan.finite_check_read()
    print(f"rep {rep}: {done} steps in {time.time() - t:.1f} s, first non-finite step {bad}, max|u| {float(u.abs().max()):.3f}", flush=True)
    outs.append((u, v))
print("bit-identical:", bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])))
hu, hv = u0.cpu().pin_memory(), v0.cpu().pin_memory()
plan.rollout_host(hu, hv, 200, apply_cnn=True)
u, v = u0.clone(), v0.clone(); plan.step(u, v, nsteps=200, apply_cnn=True)
print("host rollout vs device rollout relL2:", float((hu.cuda() - u).norm() / u.norm()))
