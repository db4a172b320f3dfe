"""Grid containers with the names the reference scripts use from ``torch_cfd.grids``
(call sites: reference src/inference.py:66,69; src/data_generation.py:131-143,104-111; test/sim_test.py:36-73).

State lives in plain torch tensors ``[batch, nx, ny]`` (axis -2 = x, axis -1 = y contiguous); the MAC
staggering is carried as metadata: u at offset (1, 1/2), v at (1/2, 1), pressure at (1/2, 1/2).
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence, Tuple

import torch


class Grid:
    def __init__(self, shape: Sequence[int], domain: Optional[Sequence[Tuple[float, float]]] = None,
                 step=None, device=None):
        self.shape = tuple(int(s) for s in shape)
        if domain is None:
            if step is None:
                step = 1.0
            steps = (step,) * len(self.shape) if isinstance(step, (int, float)) else tuple(step)
            domain = tuple((0.0, s * n) for s, n in zip(steps, self.shape))
        self.domain = tuple((float(a), float(b)) for a, b in domain)
        self.device = device
        self.ndim = len(self.shape)

    @property
    def step(self) -> Tuple[float, ...]:
        return tuple((hi - lo) / n for (lo, hi), n in zip(self.domain, self.shape))

    @property
    def cell_center(self) -> Tuple[float, ...]:
        return (0.5,) * self.ndim

    def __repr__(self):
        return f"Grid(shape={self.shape}, domain={self.domain})"


class GridVariable:
    """A field sample with its staggering offset and (periodic) boundary condition."""

    def __init__(self, array: torch.Tensor, offset: Tuple[float, ...], grid: Grid, bc=None):
        self.data = array
        self.offset = tuple(offset)
        self.grid = grid
        self.bc = bc

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def device(self):
        return self.data.device

    def clone(self):
        return GridVariable(self.data.clone(), self.offset, self.grid, self.bc)

    def to(self, *args, **kwargs):
        return GridVariable(self.data.to(*args, **kwargs), self.offset, self.grid, self.bc)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        """torch functions applied to a GridVariable act on its array (``torch.linalg.norm(fdm.divergence(v0))``,
        reference test/sim_test.py:45); the result is a plain tensor."""
        def unwrap(x):
            if isinstance(x, GridVariable):
                return x.data
            if isinstance(x, (list, tuple)) and not isinstance(x, GridVariableVector):
                return type(x)(unwrap(y) for y in x)
            return x
        return func(*[unwrap(a) for a in args], **{k: unwrap(v) for k, v in (kwargs or {}).items()})


class GridVariableVector(tuple):
    """Tuple of GridVariables (the velocity vector)."""

    def __new__(cls, items: Iterable[GridVariable]):
        return super().__new__(cls, tuple(items))

    @property
    def data(self):
        return tuple(v.data for v in self)

    @property
    def shape(self):
        return self[0].data.shape

    @property
    def dtype(self):
        return self[0].data.dtype

    @property
    def device(self):
        return self[0].data.device

    def clone(self):
        return GridVariableVector(v.clone() for v in self)

    def to(self, *args, **kwargs):
        return GridVariableVector(v.to(*args, **kwargs) for v in self)
