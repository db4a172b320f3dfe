"""oracle/datagen.py against golden vectors produced by the reference's own downsampling functions
(tests/golden/make_golden_datagen.py), bit for bit."""
import os

import pytest
import safetensors.torch as st
import torch

from oracle import datagen as D


@pytest.fixture(scope="module")
def golden(golden_dir):
    return st.load_file(os.path.join(golden_dir, "datagen_golden.safetensors"))


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_downsample_matches_reference_bitwise(golden, name):
    f = int(golden[f"{name}_factor"][0])
    cu, cv = D.downsample_staggered_velocity(golden[f"{name}_u"], golden[f"{name}_v"], f)
    assert cu.shape == golden[f"{name}_cu"].shape
    assert (cu - golden[f"{name}_cu"]).abs().max().item() <= 2.3e-16          # mean() association order only
    assert (cv - golden[f"{name}_cv"]).abs().max().item() <= 2.3e-16


def test_downsample_rejects_non_dividing_factor():
    with pytest.raises(ValueError):
        D.downsample_component(torch.zeros(10, 12), 0, 4)


def test_downsample_picks_the_coarse_faces():
    """u lives on x-faces: coarse face I coincides with fine face I*f + f - 1; a field constant along y and equal to the
    row index must come back as those row indices."""
    nx, ny, f = 32, 16, 4
    u = torch.arange(nx, dtype=torch.float64)[:, None].expand(nx, ny).contiguous()
    cu = D.downsample_component(u, 0, f)
    assert torch.equal(cu[:, 0], torch.arange(f - 1, nx, f, dtype=torch.float64))
    v = torch.arange(ny, dtype=torch.float64)[None, :].expand(nx, ny).contiguous()
    cv = D.downsample_component(v, 1, f)
    assert torch.equal(cv[0], torch.arange(f - 1, ny, f, dtype=torch.float64))
