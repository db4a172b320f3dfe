"""What limits the end-to-end (host-buffer) path when all GPUs of the box step at once?  

    python scripts/host_link_probe.py                      # 1 process, GPU 0
    torchrun --nproc-per-node 8 scripts/host_link_probe.py # 8 processes, one per GPU

Per rank, between barriers, with CUDA events / wall clock:
  (i)   bare pinned-memory copies of this workload's size (2 x 16.8 MB each way per step): H2D alone, D2H alone, both at once;
  (ii)  the same with the copying thread pinned to a core set of its own (os.sched_setaffinity, ranks spread over the cores);
  (iii) the CPU time to ENQUEUE one step (launch-thread cost; torchrun sets OMP_NUM_THREADS=1 -- irrelevant here, the
        step is launched from one thread) against the step's device time;
  (iv)  rollout_host itself (H2D | step | D2H pipeline).
Rank 0 prints one JSON line with per-rank means and the aggregate GB/s."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)

def gather(x):
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    if world == 1:
        return [float(x)]
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]

n = 64 * 256 * 256
h_in = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
h_out = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
d = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(2)]
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
MB = 2 * n * 4 / 1e6

def copies(h2d, d2h, reps=20):
    barrier()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s_in):
                for a, b in zip(d, h_in): a.copy_(b, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s_out):
                for a, b in zip(h_out, d): a.copy_(b, non_blocking=True)
        s_in.synchronize(); s_out.synchronize()
    dt = (time.perf_counter() - t) / reps
    barrier()
    return dt

res = {"world": world, "bytes_each_way_MB": MB}
for tag in ("default_affinity", "spread_affinity"):
    if tag == "spread_affinity":
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(world, 1))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        res["cores_total"] = len(cores)
    for name, (a, b) in (("h2d", (True, False)), ("d2h", (False, True)), ("both", (True, True))):
        dt = copies(a, b)
        per_rank = gather(MB * ((1 if a else 0) + (1 if b else 0)) / dt / 1e3)        # GB/s moved by this rank
        if rank == 0:
            res[f"{tag}.{name}"] = {"ms": dt * 1e3, "per_rank_GBs_mean": sum(per_rank) / world, "aggregate_GBs": sum(per_rank),
                                    "per_rank_GBs_min": min(per_rank)}

import bench
from ml_accelerated_simulation_b200 import Plan, PlanConfig, models
cfgj, sd = bench._model_state()
model = models.buildModel(cfgj); model.load_state_dict(sd)
plan = Plan(PlanConfig(nx=256, ny=256, batch=64, device=local)); model.attach(plan)
u0, v0 = bench._device_state(64, 42 + rank, dev)
u, v = u0.clone(), v0.clone()
plan.step(u, v, nsteps=5, apply_cnn=True); barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t = time.perf_counter(); e0.record(); plan.step(u, v, nsteps=20, apply_cnn=True); e1.record(); t_enq = (time.perf_counter() - t) / 20
barrier()
dev_ms = gather(e0.elapsed_time(e1) / 20); enq_ms = gather(t_enq * 1e3)
hu, hv = u0.cpu().pin_memory(), v0.cpu().pin_memory()
plan.rollout_host(hu, hv, 1, apply_cnn=True); barrier()
t = time.perf_counter()
for _ in range(20): plan.rollout_host(hu, hv, 1, apply_cnn=True)
e2e = gather((time.perf_counter() - t) / 20 * 1e3)
barrier()
if rank == 0:
    res["step_device_ms"] = {"mean": sum(dev_ms) / world, "max": max(dev_ms)}
    res["step_enqueue_cpu_ms"] = {"mean": sum(enq_ms) / world, "max": max(enq_ms)}
    res["rollout_host_ms"] = {"mean": sum(e2e) / world, "max": max(e2e), "per_rank": e2e}
    res["e2e_bytes_per_step_all_ranks_MB"] = 2 * MB * world
    res["e2e_aggregate_GBs"] = 2 * MB * world / (max(e2e))
    print(json.dumps(res), flush=True)
plan.close()
if world > 1:
    dist.destroy_process_group()
