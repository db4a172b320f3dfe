"""Data generation on the B200 path: the reference's ``src/data_generation.py`` loop with every per-step operation in
the sm_100a kernels.

Mirrors, with the same names and argument meaning:
  * ``downsample_staggered_velocity(source_grid, destination_grid, velocity)``  (reference :97-111)
  * the pair loop of ``main`` (:193-213): ``time_between_samples/dt`` fine steps, downsample, ``inner_step`` fine
    steps and ONE coarse step of the downsampled field -- ``PairGenerator.next_pair`` runs it as one stream-ordered
    device pipeline (``mlsim_datagen_pair``), nothing returns to the host in between;
  * the sample file (:215-221): safetensors keys ``{idx}_f`` (fine, ``[2, N, N]``) and ``{idx}_c`` (coarse,
    ``[2, n, n]``), component 0 = u, 1 = v, axis 0 = x, dtype float64 -- readable by the reference's loader
    (src/data/dataloader.py:66-80).

Not mirrored: the Google-Drive upload (:223-226) and the import-time file creation (:47-48) -- control plane, out of
scope.
"""
from __future__ import annotations

import math
from typing import Dict, Iterator, Optional, Tuple

import torch

from . import _lib
from .equations import stable_time_step
from .grids import Grid, GridVariable, GridVariableVector
from .initial_conditions import filtered_velocity_field
from .plan import Plan, PlanConfig


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).contiguous()


def downsample_fields(u: torch.Tensor, v: torch.Tensor, factor: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """[batch, nx, ny] float32 CUDA tensors -> [batch, nx/factor, ny/factor] (mlsim_downsample_staggered)."""
    if not (u.is_cuda and v.is_cuda):
        raise RuntimeError("the B200 data-generation path runs on CUDA tensors only (no CPU fallback)")
    u, v = _f32c(u), _f32c(v)
    b, nx, ny = u.shape
    if nx % factor or ny % factor:
        raise ValueError(f"`block_size` must divide `array.shape`; got {factor}, {(nx, ny)}.")
    cu = torch.empty((b, nx // factor, ny // factor), dtype=torch.float32, device=u.device)
    cv = torch.empty_like(cu)
    lib = _lib.load()
    with torch.cuda.device(u.device):
        _lib.check(lib.mlsim_downsample_staggered(u.data_ptr(), v.data_ptr(), cu.data_ptr(), cv.data_ptr(), b, nx, ny,
                                                  int(factor), torch.cuda.current_stream(u.device).cuda_stream))
    return cu, cv


def downsample_staggered_velocity(source_grid: Grid, destination_grid: Grid, velocity) -> GridVariableVector:
    """Reference src/data_generation.py:97-111 (u sliced along x / averaged over y, v the other way round)."""
    factor = round(destination_grid.step[0] / source_grid.step[0])
    u_var, v_var = velocity[0], velocity[1]
    squeeze = u_var.data.dim() == 2
    ud = u_var.data[None] if squeeze else u_var.data
    vd = v_var.data[None] if squeeze else v_var.data
    cu, cv = downsample_fields(ud, vd, factor)
    cu, cv = cu.to(u_var.data.dtype), cv.to(v_var.data.dtype)
    if squeeze:
        cu, cv = cu[0], cv[0]
    return GridVariableVector((GridVariable(cu, u_var.offset, destination_grid, u_var.bc),
                               GridVariable(cv, v_var.offset, destination_grid, v_var.bc)))


class PairGenerator:
    """The (fine, coarse) pair pipeline of reference ``main`` (:113-213) for one batch of trajectories."""

    def __init__(self, high_res: int = 1024, low_res: int = 64, batch_size: int = 64, max_velocity: float = 7.0,
                 peak_wavenumber: float = 4.0, cfl_safety_factor: float = 0.5, viscosity: float = 1e-3,
                 density: float = 1.0, drag: float = 0.1, time_between_samples: float = 1.0, device: int = 0):
        self.high_res, self.low_res, self.batch = high_res, low_res, batch_size
        self.max_velocity, self.peak_wavenumber = max_velocity, peak_wavenumber
        diam = 2 * math.pi
        self.device = torch.device("cuda", device)
        self.full_grid = Grid((high_res, high_res), domain=((0, diam), (0, diam)), device=self.device)
        self.coarse_grid = Grid((low_res, low_res), domain=((0, diam), (0, diam)), device=self.device)
        self.dt = stable_time_step(dx=min(self.full_grid.step), max_velocity=max_velocity,
                                   max_courant_number=cfl_safety_factor, viscosity=viscosity)          # :135-140
        self.delta_t = stable_time_step(dx=min(self.coarse_grid.step), max_velocity=max_velocity,
                                        max_courant_number=cfl_safety_factor, viscosity=viscosity)     # :142-147
        self.inner_step = round(self.delta_t / self.dt)                                                 # :150
        self.time_between_samples = float(time_between_samples)
        self.steps_between = round(time_between_samples / self.dt)                                      # :205
        common = dict(batch=batch_size, viscosity=viscosity, density=density, drag=drag,
                      forcing_wavenumber=int(peak_wavenumber), device=device)
        self.fine = Plan(PlanConfig(nx=high_res, ny=high_res, dt=self.dt, **common))
        self.coarse = Plan(PlanConfig(nx=low_res, ny=low_res, dt=self.delta_t, **common))
        self.u: Optional[torch.Tensor] = None
        self.v: Optional[torch.Tensor] = None

    def reset(self, random_state: int, iterations: int = 50) -> None:
        """New initial condition (:194-196)."""
        v0 = filtered_velocity_field(self.full_grid, self.max_velocity, self.peak_wavenumber, iterations=iterations,
                                     random_state=random_state, device=self.device, batch_size=self.batch,
                                     dtype=torch.float32)
        self.u, self.v = v0[0].data.contiguous(), v0[1].data.contiguous()

    def next_pair(self, steps_between: Optional[int] = None, inner_step: Optional[int] = None):
        """Advance the fine state and return ((u_f, v_f), (u_c, v_c)) of the next stored pair (:203-213)."""
        if self.u is None:
            raise RuntimeError("call reset(random_state) first")
        nb = self.steps_between if steps_between is None else steps_between
        ni = self.inner_step if inner_step is None else inner_step
        n = self.low_res
        cu = torch.empty((self.batch, n, n), dtype=torch.float32, device=self.device)
        cv = torch.empty_like(cu)
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.mlsim_datagen_pair(self.fine._h, self.coarse._h, self.u.data_ptr(), self.v.data_ptr(),
                                              cu.data_ptr(), cv.data_ptr(), int(nb), int(ni),
                                              torch.cuda.current_stream().cuda_stream))
        return (self.u, self.v), (cu, cv)

    def close(self):
        self.fine.close()
        self.coarse.close()


def pairs_to_dataset(pairs, first_index: int = 0, fine_as_coarse: bool = False) -> Dict[str, torch.Tensor]:
    """The reference's sample layout (:215-220): ``{idx}_f`` / ``{idx}_c`` = stack((u, v))[:, b] as float64 on the host.

    ``fine_as_coarse``: store the fine field downsampled to the coarse grid.  The reference writer stores ``_f`` at
    full resolution while its loader computes ``c_full - coarse`` (src/data/dataloader.py:75, the downsampling there
    is commented out, :69-73) -- the two halves of the reference disagree; this switch produces the file the loader
    can actually consume, using the same device kernel."""
    dataset: Dict[str, torch.Tensor] = {}
    for pi, ((uf, vf), (uc, vc)) in enumerate(pairs):
        batch = uf.shape[0]
        if fine_as_coarse and uf.shape[-1] != uc.shape[-1]:
            uf, vf = downsample_fields(uf, vf, uf.shape[-1] // uc.shape[-1])
        fine = torch.stack((uf, vf)).to(torch.float64).cpu()
        coarse = torch.stack((uc, vc)).to(torch.float64).cpu()
        for b in range(batch):
            dataset[f"{first_index + pi * batch + b}_f"] = fine[:, b].contiguous()
            dataset[f"{first_index + pi * batch + b}_c"] = coarse[:, b].contiguous()
    return dataset


def generate(path: str, sample_size: int = 1, pairs_per_sample: Optional[int] = None, seed: int = 0,
             simulation_time: float = 30.0, **kw) -> Iterator[Dict[str, torch.Tensor]]:
    """The outer loops of reference ``main`` (:192-228) without the upload: yields (and saves to ``path``) one dataset
    dict per stored pair, exactly as the reference overwrites its save file after every pair."""
    from safetensors.torch import save_file
    gen = PairGenerator(**kw)
    try:
        for i in range(sample_size):
            gen.reset(seed + i)
            n_pairs = pairs_per_sample
            if n_pairs is None:
                # reference :203 -- round(((simulation_time/time_between_samples)/dt)//inner_step): floor-divide the
                # float first, then round
                n_pairs = round(((simulation_time / gen.time_between_samples) / gen.dt) // gen.inner_step)
            for _ in range(n_pairs):
                (uf, vf), (uc, vc) = gen.next_pair()
                dataset = pairs_to_dataset([((uf.clone(), vf.clone()), (uc, vc))])
                save_file(dataset, path)
                yield dataset
    finally:
        gen.close()
