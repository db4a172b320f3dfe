"""Python owner of an ``mlsim_plan`` (include/mlsim.h): one grid/batch/physics configuration on one B200.

PyTorch is used for device memory and streams only; every numerical operation on the hot path is a
hand-written sm_100a kernel reached through the C ABI.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _lib

ADVECTION = {"van_leer": 0, "upwind": 1, "lax_wendroff": 2}


@dataclasses.dataclass
class PlanConfig:
    """Physics defaults are the reference's (src/inference.py:54-93)."""
    nx: int = 64
    ny: int = 64
    batch: int = 1
    lx: float = 2 * math.pi
    ly: float = 2 * math.pi
    dt: Optional[float] = None           # default: stable_time_step(min(h), max_velocity, cfl, viscosity)
    viscosity: float = 1e-3
    density: float = 1.0
    drag: float = 0.1
    forcing_wavenumber: int = 4
    forcing_scale: float = 1.0
    advection: str = "van_leer"
    lw_dt_scale: float = 1.0
    max_velocity: float = 7.0
    cfl: float = 0.5
    device: int = 0
    slab_rank: int = 0                   # slab decomposition along x (see slab.py); slab_world == 0: plain plan
    slab_world: int = 0
    solver_blocks: int = 0               # execution tuning, 0 = measured default (see include/mlsim.h)
    host_blocks: int = 0
    fuse_first_layer: bool = True        # first conv layer computed inside the first hidden layer's kernel (when the net allows)
    cluster_projection: bool = False
    step_graph: bool = False             # replay a captured CUDA graph of the step (saves host enqueue time only)

    def resolved_dt(self) -> float:
        if self.dt is not None:
            return float(self.dt)
        from .equations import stable_time_step
        return stable_time_step(dx=min(self.lx / self.nx, self.ly / self.ny), max_velocity=self.max_velocity,
                                max_courant_number=self.cfl, viscosity=self.viscosity)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Plan:
    def __init__(self, cfg: PlanConfig):
        if not torch.cuda.is_available():
            raise RuntimeError("mlsim needs a CUDA device (sm_100a); there is no CPU fallback")
        self._lib = _lib.load()
        self.cfg = cfg
        self.dt = cfg.resolved_dt()
        c = _lib.MlsimConfig(cfg.nx, cfg.ny, cfg.batch, cfg.lx, cfg.ly, self.dt, cfg.viscosity, cfg.density,
                             cfg.drag, cfg.forcing_wavenumber, cfg.forcing_scale, ADVECTION[cfg.advection],
                             cfg.lw_dt_scale, cfg.device, cfg.slab_rank, cfg.slab_world, cfg.solver_blocks,
                             cfg.host_blocks, int(not cfg.fuse_first_layer), int(cfg.cluster_projection),
                             int(cfg.step_graph))
        h = C.c_void_p()
        _lib.check(self._lib.mlsim_plan_create(C.byref(c), C.byref(h)))
        self._h = h
        self.device = torch.device("cuda", cfg.device)
        self.shape = (cfg.batch, cfg.nx, cfg.ny)
        self.cnn_out_channels = 0
        self.has_cnn = False

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.mlsim_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _field(self, t: torch.Tensor, name: str) -> torch.Tensor:
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                and tuple(t.shape) == self.shape and t.device.index == self.cfg.device):
            raise ValueError(f"{name} must be a contiguous float32 CUDA tensor of shape {self.shape} on cuda:{self.cfg.device}")
        return t

    def new_field(self) -> torch.Tensor:
        return torch.empty(self.shape, dtype=torch.float32, device=self.device)

    def _stream(self) -> int:
        # the current stream of the PLAN's device (not of whatever device happens to be current)
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ correction network
    def set_cnn(self, layers: Sequence[Tuple[torch.Tensor, torch.Tensor]], shortcuts=None, final_relu: bool = False):
        """layers: [(weight[cout,cin,3,3], bias[cout]), ...] on the host, any float dtype (BN already folded).
        shortcuts: {layer: (src_layer, w1x1 | None, b1x1 | None)} -- ResNet ``act(x + y)`` (mlsim_cnn_set_shortcut);
        final_relu: activation after the last layer too (ResNet blocks end with one, CNN blocks do not)."""
        n = len(layers)
        _lib.check(self._lib.mlsim_cnn_begin(self._h, n))
        for i, (w, b) in enumerate(layers):
            w64 = w.detach().to("cpu", torch.float64).contiguous()
            b64 = b.detach().to("cpu", torch.float64).contiguous()
            cout, cin, kh, kw = w64.shape
            if kh != 3 or kw != 3:
                raise ValueError("only 3x3 convolutions are on the hot path")
            _lib.check(self._lib.mlsim_cnn_set_layer(self._h, i, cin, cout, 3, w64.data_ptr(), b64.data_ptr(), 1,
                                                      1 if (i < n - 1 or final_relu) else 0))
        for layer, (src, w1, b1) in sorted((shortcuts or {}).items()):
            if w1 is None:
                _lib.check(self._lib.mlsim_cnn_set_shortcut(self._h, layer, src, None, None, 1))
            else:
                w64 = w1.detach().to("cpu", torch.float64).reshape(w1.shape[0], -1).contiguous()
                b64 = b1.detach().to("cpu", torch.float64).contiguous()
                _lib.check(self._lib.mlsim_cnn_set_shortcut(self._h, layer, src, w64.data_ptr(), b64.data_ptr(), 1))
        _lib.check(self._lib.mlsim_cnn_finalize(self._h))
        self.cnn_out_channels = int(layers[-1][0].shape[0])
        self.has_cnn = True

    # ------------------------------------------------------------------ hot path
    def step(self, u: torch.Tensor, v: torch.Tensor, nsteps: int = 1, apply_cnn: bool = False,
             input_scale: float = 1.0, p: Optional[torch.Tensor] = None):
        """In place: ``for _ in range(nsteps): v, p = step_fn.forward(v, dt, ns2d); v += model(v)``
        (reference src/inference.py:100-102)."""
        self._field(u, "u"); self._field(v, "v")
        if p is not None:
            self._field(p, "p")
        if apply_cnn and self.has_cnn and self.cnn_out_channels != 2:
            raise ValueError(f"v += model(v) needs a 2-channel network output, this one has {self.cnn_out_channels}")
        _lib.check(self._lib.mlsim_step(self._h, u.data_ptr(), v.data_ptr(), _ptr(p), nsteps, int(apply_cnn),
                                        float(input_scale), self._stream()))

    def step_f64(self, u: torch.Tensor, v: torch.Tensor, nsteps: int = 1, apply_cnn: bool = False,
                 input_scale: float = 1.0, want_p: bool = True, inplace: bool = False):
        """The same loop for float64 state (what the reference scripts hold): returns (u, v, p) as NEW float64 tensors
        (or updates u, v in place with ``inplace``); fp32 arithmetic, one fused conversion pass each way."""
        for t, name in ((u, "u"), (v, "v")):
            if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == self.shape
                    and t.device.index == self.cfg.device):
                raise ValueError(f"{name} must be a contiguous float64 CUDA tensor of shape {self.shape} on cuda:{self.cfg.device}")
        if apply_cnn and self.has_cnn and self.cnn_out_channels != 2:
            raise ValueError(f"v += model(v) needs a 2-channel network output, this one has {self.cnn_out_channels}")
        uo, vo = (u, v) if inplace else (torch.empty_like(u), torch.empty_like(v))
        p = torch.empty_like(u) if want_p else None
        _lib.check(self._lib.mlsim_step_f64(self._h, u.data_ptr(), v.data_ptr(), uo.data_ptr(), vo.data_ptr(), _ptr(p),
                                            nsteps, int(apply_cnn), float(input_scale), self._stream()))
        return uo, vo, p

    def rollout_host(self, hu: torch.Tensor, hv: torch.Tensor, nsteps: int, apply_cnn: bool = False,
                     input_scale: float = 1.0, hp: Optional[torch.Tensor] = None):
        """Same loop through host tensors (H2D, nsteps, D2H inside); in place on hu, hv."""
        for t in (hu, hv) + ((hp,) if hp is not None else ()):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != self.shape:
                raise ValueError(f"host fields must be contiguous float32 CPU tensors of shape {self.shape}")
        _lib.check(self._lib.mlsim_rollout_host(self._h, hu.data_ptr(), hv.data_ptr(), _ptr(hp), nsteps,
                                                int(apply_cnn), float(input_scale)))

    def set_host_blocks(self, sizes=None) -> None:
        """Explicit block sizes of rollout_host's H2D | step | D2H pipeline (None: automatic)."""
        if not sizes:
            _lib.check(self._lib.mlsim_set_host_blocks(self._h, None, 0))
        else:
            arr = (C.c_int32 * len(sizes))(*[int(x) for x in sizes])
            _lib.check(self._lib.mlsim_set_host_blocks(self._h, arr, len(sizes)))

    # ------------------------------------------------------------------ pieces (parity hooks)
    def explicit_terms(self, u, v):
        self._field(u, "u"); self._field(v, "v")
        du, dv = self.new_field(), self.new_field()
        _lib.check(self._lib.mlsim_explicit_terms(self._h, u.data_ptr(), v.data_ptr(), du.data_ptr(), dv.data_ptr(),
                                                  self._stream()))
        return du, dv

    def project(self, u, v, want_p: bool = True):
        """In place on u, v; returns p (or None)."""
        self._field(u, "u"); self._field(v, "v")
        p = self.new_field() if want_p else None
        _lib.check(self._lib.mlsim_project(self._h, u.data_ptr(), v.data_ptr(), _ptr(p), self._stream()))
        return p

    def spectral_multiply(self, field: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
        """irfft2(rfft2(field) * table) on the plan's own transform kernels; table: host fp64 [nx, ny//2+1]."""
        self._field(field, "field")
        t = table.detach().to("cpu", torch.float64).contiguous()
        if tuple(t.shape) != (self.cfg.nx, self.cfg.ny // 2 + 1):
            raise ValueError(f"table must have shape {(self.cfg.nx, self.cfg.ny // 2 + 1)}")
        out = self.new_field()
        _lib.check(self._lib.mlsim_spectral_multiply(self._h, t.data_ptr(), field.data_ptr(), out.data_ptr(), self._stream()))
        return out

    def normalize_speed(self, u, v, max_velocity: float) -> None:
        """In place, per trajectory: u, v *= max_velocity / max sqrt(u^2 + v^2)."""
        self._field(u, "u"); self._field(v, "v")
        _lib.check(self._lib.mlsim_normalize_speed(self._h, u.data_ptr(), v.data_ptr(), float(max_velocity), self._stream()))

    def divergence(self, u, v):
        self._field(u, "u"); self._field(v, "v")
        d = self.new_field()
        _lib.check(self._lib.mlsim_divergence(self._h, u.data_ptr(), v.data_ptr(), d.data_ptr(), self._stream()))
        return d

    def cnn_forward(self, u, v, input_scale: float = 1.0):
        """y = model(stack((u, v), 1) * input_scale) -> [batch, cout, nx, ny] float32."""
        self._field(u, "u"); self._field(v, "v")
        y = torch.empty((self.cfg.batch, self.cnn_out_channels, self.cfg.nx, self.cfg.ny), dtype=torch.float32,
                        device=self.device)
        _lib.check(self._lib.mlsim_cnn_forward(self._h, u.data_ptr(), v.data_ptr(), y.data_ptr(), float(input_scale),
                                               self._stream()))
        return y

    def cnn_correct(self, u, v, input_scale: float = 1.0):
        self._field(u, "u"); self._field(v, "v")
        _lib.check(self._lib.mlsim_cnn_correct(self._h, u.data_ptr(), v.data_ptr(), float(input_scale), self._stream()))

    # ------------------------------------------------------------------ divergence guard
    def finite_check_enable(self, every: int) -> None:
        """Scan (u, v) for non-finite values after every `every`-th step on the device (0: off); no sync on the step path."""
        _lib.check(self._lib.mlsim_finite_check_enable(self._h, int(every)))

    def finite_check_read(self):
        """(first_bad_step or -1, steps_issued) as published by the device so far; non-blocking."""
        bad, done = C.c_int64(-1), C.c_int64(0)
        _lib.check(self._lib.mlsim_finite_check_read(self._h, C.byref(bad), C.byref(done)))
        return int(bad.value), int(done.value)

    # ------------------------------------------------------------------ introspection
    def launches_per_step(self, apply_cnn: bool) -> int:
        return int(self._lib.mlsim_launches_per_step(self._h, int(apply_cnn)))

    def step_graph_launches(self) -> int:
        """cudaGraphLaunch calls issued by step / step_f64 so far (0 while the plan steps with eager launches)."""
        return int(self._lib.mlsim_step_graph_launches(self._h))

    def workspace_bytes(self) -> int:
        return int(self._lib.mlsim_workspace_bytes(self._h))

    def profile_enable(self, kernel_class: int, max_samples: int = 4096) -> None:
        _lib.check(self._lib.mlsim_profile_enable(self._h, kernel_class, max_samples))

    def profile_read(self):
        """(total device ms, launches) of the profiled kernel class since the last read."""
        ms, n = C.c_float(0.0), C.c_int32(0)
        _lib.check(self._lib.mlsim_profile_read(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def conv_debug(self):
        out = (C.c_double * 8)()
        n = self._lib.mlsim_conv_debug(self._h, out)
        return n, list(out)

    def conv_debug_first(self):
        out = (C.c_double * 8)()
        n = self._lib.mlsim_conv_debug_first(self._h, out)
        return n, list(out)

    def time_kernel(self, which: int, warmup: int = 3, iters: int = 10) -> float:
        ms = C.c_float(0.0)
        _lib.check(self._lib.mlsim_time_kernel(self._h, which, warmup, iters, C.byref(ms), self._stream()))
        return float(ms.value)
