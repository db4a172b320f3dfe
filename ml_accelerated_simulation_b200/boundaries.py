"""Boundary-condition tags (``torch_cfd.boundaries`` names).  Only periodic domains are on the hot path."""
from __future__ import annotations


class PeriodicBoundaryConditions:
    types = (("periodic", "periodic"), ("periodic", "periodic"))

    def __repr__(self):
        return "PeriodicBoundaryConditions()"


def periodic_boundary_conditions(ndim: int = 2):
    return PeriodicBoundaryConditions()


def get_pressure_bc_from_velocity(v):
    """Reference call sites ignore the result (src/inference.py:79, src/data_generation.py:156)."""
    return PeriodicBoundaryConditions()
