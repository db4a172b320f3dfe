"""``torch_cfd.forcings.KolmogorovForcing`` as constructed at reference src/inference.py:81-82."""
from __future__ import annotations


class KolmogorovForcing:
    """f = (scale * sin(k * y), 0) evaluated at the u sample points; the plan tabulates it per column."""

    def __init__(self, diam: float = None, wave_number: int = 1, grid=None, offsets=None, scale: float = 1.0,
                 swap_xy: bool = False, **_ignored):
        if swap_xy:
            raise NotImplementedError("swap_xy forcing is not on the hot path")
        # The reference builds it with diam = 2*pi = the domain length (src/inference.py:61,81-82); only there do the
        # candidate readings of the third-party formula ("sin(k y)" / "sin(2 pi k y / diam)") coincide, so any other
        # diam is refused instead of guessed (the C ABI refuses ly != 2*pi with forcing for the same reason).
        import math
        if diam is not None and abs(float(diam) - 2 * math.pi) > 1e-9:
            raise NotImplementedError("KolmogorovForcing is on the hot path for diam = 2*pi only")
        self.diam = diam
        self.wave_number = int(wave_number)
        self.grid = grid
        self.offsets = offsets
        self.scale = float(scale)
