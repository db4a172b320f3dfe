"""``torch_cfd.forcings.KolmogorovForcing`` as constructed at reference src/inference.py:81-82."""
from __future__ import annotations


class KolmogorovForcing:
    """f = (scale * sin(k * y), 0) evaluated at the u sample points; the plan tabulates it per column."""

    def __init__(self, diam: float = None, wave_number: int = 1, grid=None, offsets=None, scale: float = 1.0,
                 swap_xy: bool = False, **_ignored):
        if swap_xy:
            raise NotImplementedError("swap_xy forcing is not on the hot path")
        self.diam = diam
        self.wave_number = int(wave_number)
        self.grid = grid
        self.offsets = offsets
        self.scale = float(scale)
