"""``models`` import name of the reference (src/models/__init__.py:1: ``from .modelbuilder import buildModel``) ->
the B200 correction-network front end (see compat/README.md).  Alias package: no code of its own."""
from ml_accelerated_simulation_b200.models import CNN, ResNetBlock, ResNeXtBlock, buildModel, load_config, load_model  # noqa: F401


def getModel(config):
    """reference src/models/tools.py:80-124 for the block families on the B200 path."""
    name = str(config["name"]).upper()
    if name == "CNN":
        return CNN(config)
    if name in ("RESNETBLOCK", "RESNETBASICBLOCK", "RESNETBOTTLENECKBLOCK"):
        return ResNetBlock(config)
    if name == "RESNEXTBLOCK":
        return ResNeXtBlock(config)
    raise ValueError(f"Model Name [{name}] is not on the B200 hot path")
