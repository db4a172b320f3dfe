"""torchrun worker of tests/test_gpu_slab.py::test_slab_nccl_matches_single_gpu (one process per GPU, NCCL):
runs the slab-decomposed rollout and compares the gathered result with the oracle and with the plain single-GPU plan."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import safetensors.torch as st
import torch
import torch.distributed as dist


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm()).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--ny", type=int, default=128)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--p2p", type=int, default=0)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ml_accelerated_simulation_b200 as m
    from ml_accelerated_simulation_b200 import models
    from ml_accelerated_simulation_b200.slab import CudaSlabBackend, SlabComm, SlabGeometry, SlabStepper
    from oracle import cnn as OC
    from oracle import solver as S

    g = os.path.join(ROOT, "tests", "golden")
    cfgj = json.load(open(os.path.join(g, "MLACFD.json")))
    sd = st.load_file(os.path.join(g, "mlacfd_weights.safetensors"))
    cfg = S.SolverConfig(nx=a.nx, ny=a.ny)
    u64, v64 = S.filtered_velocity_field(cfg, a.batch, seed=11, iterations=4)      # same seed -> same field on every rank
    dev = torch.device("cuda", local)
    geom = SlabGeometry(nx=a.nx, ny=a.ny, batch=a.batch, rank=rank, world=world)
    be = CudaSlabBackend(geom, device=local)
    model = models.buildModel(cfgj)
    model.load_state_dict(sd)
    be.set_cnn(model.cnn.folded_layers())
    stp = SlabStepper(be, SlabComm(geom, p2p=bool(a.p2p)))

    res = {}
    stp.scatter(u64.float().to(dev), v64.float().to(dev))
    stp.step(a.steps)
    u, v = stp.gather()
    stp.scatter(u64.float().to(dev), v64.float().to(dev))
    stp.step(a.steps, apply_cnn=True)
    uc, vc = stp.gather()
    torch.cuda.synchronize()
    if rank == 0:
        ou, ov, _ = OC.rollout(u64, v64, cfg, a.steps)
        qu, qv, _ = OC.rollout(u64.float(), v64.float(), cfg, a.steps)
        eu, ev, _ = OC.rollout(u64, v64, cfg, a.steps, cfgj, sd, 1.0, round_bf16=True)
        plan = m.Plan(m.PlanConfig(nx=a.nx, ny=a.ny, batch=a.batch, device=local))
        model.attach(plan)
        pu, pv = u64.float().to(dev), v64.float().to(dev)
        plan.step(pu, pv, nsteps=a.steps, apply_cnn=True)
        res = {"world": world, "p2p": a.p2p, "solver_vs_oracle_u": rel(u, ou), "solver_vs_oracle_v": rel(v, ov),
               "o32_vs_oracle_u": rel(qu, ou), "o32_vs_oracle_v": rel(qv, ov),
               "cnn_vs_bf16_oracle_u": rel(uc, eu), "cnn_vs_bf16_oracle_v": rel(vc, ev),
               "vs_plain_plan_u": rel(uc, pu), "vs_plain_plan_v": rel(vc, pv), "exchanges": stp.exchanges}
        print(json.dumps(res), flush=True)
        if a.out:
            json.dump(res, open(a.out, "w"))
        plan.close()
    dist.barrier()
    be.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
