"""Slab decomposition (BASELINE config 5) on CPU: the product's sequencing + exchanges (slab.SlabStepper / SlabComm)
driven through a layout-exact model of the kernels (tests/slab_model.py), world sizes 1 and 2 (gloo, 127.0.0.1),
compared with the undecomposed fp64 oracle."""
import json
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ml_accelerated_simulation_b200.slab import SlabComm, SlabGeometry, SlabStepper
from oracle import cnn as OC
from oracle import solver as S

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    import safetensors.torch as st
    cfgj = json.load(open(os.path.join(HERE, "golden", "MLACFD.json")))
    sd = st.load_file(os.path.join(HERE, "golden", "mlacfd_weights.safetensors"))
    return cfgj, sd


def _run_rank(rank, world, nx, ny, batch, r0, nsteps, with_cnn):
    from slab_model import ModelSlabBackend
    cfg = S.SolverConfig(nx=nx, ny=ny)
    u0, v0 = S.filtered_velocity_field(cfg, batch, seed=7)
    geom = SlabGeometry(nx=nx, ny=ny, batch=batch, rank=rank, world=world)
    be = ModelSlabBackend(geom, cfg, r0=r0, cnn=_golden() if with_cnn else None)
    stp = SlabStepper(be, SlabComm(geom))
    stp.scatter(u0, v0)
    # projection alone first (idempotent on a divergence-free field up to round-off), then the rollout
    p = stp.project(want_p=True)
    stp.step(nsteps, apply_cnn=with_cnn, want_p=True)
    u, v = stp.gather()
    return cfg, u0, v0, u, v, stp


def _check(cfg, u0, v0, u, v, nsteps, with_cnn):
    cfgj, sd = _golden()
    ou, ov, _ = S.pressure_projection(u0, v0, cfg)
    ou, ov, op = OC.rollout(ou, ov, cfg, nsteps, cfgj if with_cnn else None, sd, 1.0)
    eu = ((u - ou).norm() / ou.norm()).item()
    ev = ((v - ov).norm() / ov.norm()).item()
    assert eu < 1e-11 and ev < 1e-11, (eu, ev)
    return op


@pytest.mark.parametrize("r0", [1, 2, 4])
def test_single_rank_slab_matches_oracle(r0):
    cfg, u0, v0, u, v, stp = _run_rank(0, 1, 64, 64, 2, r0, 2, True)
    op = _check(cfg, u0, v0, u, v, 2, True)
    assert ((stp.pressure - op).norm() / op.norm()).item() < 1e-9
    assert stp.exchanges == 3 + 2 * (4 * 3 + 1)           # project: halo + 2 a2a; per step: 4 x (halo + 2 a2a) + cnn halo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, nx, ny, batch, r0, with_cnn):
    import sys
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg, u0, v0, u, v, stp = _run_rank(rank, world, nx, ny, batch, r0, 2, with_cnn)
    if rank == 0:
        torch.save({"u": u, "v": v, "u0": u0, "v0": v0}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny,r0,with_cnn", [(2, 64, 64, 1, True), (2, 128, 64, 2, False), (4, 128, 64, 1, True)])
def test_multi_rank_gloo_slab_matches_oracle(tmp_path, world, nx, ny, r0, with_cnn):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(world, _free_port(), out, nx, ny, 1, r0, with_cnn), nprocs=world, join=True)
    res = torch.load(out)
    cfg = S.SolverConfig(nx=nx, ny=ny)
    _check(cfg, res["u0"], res["v0"], res["u"], res["v"], 2, with_cnn)


def test_geometry_config5():
    g = SlabGeometry(nx=8192, ny=8192, batch=1, rank=7, world=8)
    assert (g.nxl, g.ext_rows, g.nyc, g.kyw, g.kywp, g.ky0, g.kyv, g.r0) == (1024, 1040, 4097, 513, 516, 3591, 506, 8)
    assert sum(SlabGeometry(8192, 8192, 1, r, 8).kyv for r in range(8)) == 4097
    with pytest.raises(ValueError):
        SlabGeometry(nx=64, ny=64, batch=1, rank=0, world=4)         # 16 rows per rank
