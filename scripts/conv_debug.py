"""Cycle counters of the mid conv layer (mlsim_conv_debug) at a given shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 256
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
plan = Plan(PlanConfig(nx=nx, ny=nx, batch=batch)); model.attach(plan)
for which in (5, 8, 6):
    ms = plan.time_kernel(which, 3, 10)
    n, d = plan.conv_debug()
    names = ["mma_total", "wait_a0", "wait_input", "issue", "rows", "epi_total", "epi_wait", "-"]
    print(f"kernel class {which}: {ms:.4f} ms, {n} CTAs:", {k: round(x) for k, x in zip(names, d)})
    if d[4] > 0:
        print(f"   per input row: total {d[0]/d[4]:.0f} clk, waiting for input {d[2]/d[4]:.0f}, issuing {d[3]/d[4]:.0f}; epilogue waiting {d[6]/max(d[5],1):.2%} of its time")
    if which == 8:
        n, d = plan.conv_debug_first()
        r = max(d[5], 1)
        print(f"   first-layer group per row: total {d[0]/r:.0f} clk = stage wait {d[1]/r:.0f} + accumulator wait {d[2]/r:.0f} + drain {d[3]/r:.0f} + advance/build {d[4]/r:.0f}")
plan.close()
