"""the sample file written by ``pairs_to_dataset`` must stay readable by the reference's loader.

The reference's own ``KolmogrovFlowData`` (src/data/dataloader.py:56-80; torch + safetensors only) is imported from
/root/reference -- present in the build container, absent on the GPU box, hence a CPU test that skips without it --
and reads back a file produced by the package's writer.  The device half of the writer (the staggered downsampler
behind ``fine_as_coarse``) has its own bit-level GPU test against the reference's function bodies
(tests/test_gpu_slab.py::test_downsample_kernel_matches_reference_golden)."""
import importlib.util
import os
import sys

import pytest
import torch

REF = "/root/reference/src/data/dataloader.py"


def _reference_loader():
    if not os.path.exists(REF):
        pytest.skip("reference tree not present (GPU box)")
    sys.dont_write_bytecode = True                      # the reference tree is read-only
    spec = importlib.util.spec_from_file_location("ref_dataloader", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_written_samples_read_back_through_the_reference_loader(tmp_path):
    from safetensors.torch import save_file
    from ml_accelerated_simulation_b200.data_generation import pairs_to_dataset
    ref = _reference_loader()
    g = torch.Generator().manual_seed(0)
    batch, n = 64, 64                                   # the loader hard-codes 64 samples per file (:61,67)
    # one stored pair of a 64-trajectory batch, already on the coarse grid (what fine_as_coarse=True stores)
    uf, vf, uc, vc = (torch.randn(batch, n, n, generator=g) for _ in range(4))
    ds = pairs_to_dataset([((uf, vf), (uc, vc))], fine_as_coarse=True)
    assert len(ds) == 2 * batch
    d = tmp_path / "data"
    d.mkdir()
    save_file(ds, str(d / "sample0.safetensors"))
    data = ref.KolmogrovFlowData(str(d))
    assert len(data) == 64
    for idx in (0, 17, 63):
        coarse, dif = data[idx]
        assert coarse.dtype == torch.float64 and coarse.shape == (2, n, n) and dif.shape == (2, n, n)
        # loader: dif = c_full - coarse; coarse /= 7  (src/data/dataloader.py:75-77); component 0 = u, 1 = v
        want_c = torch.stack((uc[idx], vc[idx])).double()
        want_f = torch.stack((uf[idx], vf[idx])).double()
        assert torch.equal(coarse, want_c / 7)
        assert torch.equal(dif, want_f - want_c)


def test_reference_loader_and_writer_agree_on_the_downsampler_contract():
    """The loader's (commented-out) downsampling is the same function the writer's kernel is pinned to."""
    ref = _reference_loader()
    from oracle import datagen as D
    g = torch.Generator().manual_seed(1)
    u = torch.randn(64, 64, generator=g, dtype=torch.float64)
    v = torch.randn(64, 64, generator=g, dtype=torch.float64)
    ru = ref.downsample_staggered_velocity_component(u, 0, 16)
    rv = ref.downsample_staggered_velocity_component(v, 1, 16)
    ou, ov = D.downsample_staggered_velocity(u[None], v[None], 16)
    assert (ou[0] - ru).abs().max() <= 1e-15 and (ov[0] - rv).abs().max() <= 1e-15
