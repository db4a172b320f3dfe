"""``torch_cfd`` import name -> the B200 rollout path (see compat/README.md).  Alias package: no code of its own."""
import importlib
import sys

import ml_accelerated_simulation_b200 as _impl

__all__ = ["grids", "boundaries", "initial_conditions", "equations", "fvm", "forcings", "finite_differences",
           "tensor_utils"]

for _name in __all__:
    _mod = importlib.import_module(f"ml_accelerated_simulation_b200.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod        # `import torch_cfd.finite_differences as fdm` resolves to the same module object
    globals()[_name] = _mod

__version__ = "b200-" + getattr(_impl, "__version__", "0.2")
