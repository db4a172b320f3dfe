"""CPU check of the convolution kernels' index arithmetic: the numpy kernel model (tests/umma_model.py)
consumes the REAL operand images packed by the C ABI (mlsim_cnn_pack_host) and must reproduce a
bf16-rounded torch convolution exactly (same products, fp64 accumulate in both)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import umma_model as M


def _pack(lib, kind, w, b):
    w = np.ascontiguousarray(w, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    n = lib.mlsim_cnn_pack_host(kind, w.shape[1], w.shape[0], None, None, None, 0)
    out = np.zeros(n, dtype=np.uint8)
    assert lib.mlsim_cnn_pack_host(kind, w.shape[1], w.shape[0], w.ctypes.data, b.ctypes.data, out.ctypes.data, n) == n
    return out


def _q(t):
    return t.to(torch.bfloat16).double()


@pytest.mark.parametrize("nx,ny,rs", [(12, 40, 4), (11, 24, 5), (9, 136, 16)])
def test_three_layer_network_matches_torch(built_lib, nx, ny, rs):
    from ml_accelerated_simulation_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(nx * 1000 + ny)
    B = 2 if ny < 100 else 1
    w0, b0 = rng.standard_normal((5, 2, 3, 3)) * 0.3, rng.standard_normal(5) * 0.1
    w1, b1 = rng.standard_normal((7, 5, 3, 3)) * 0.3, rng.standard_normal(7) * 0.1
    w2, b2 = rng.standard_normal((2, 7, 3, 3)) * 0.3, rng.standard_normal(2) * 0.1
    u = rng.standard_normal((B, nx, ny)).astype(np.float32)
    v = rng.standard_normal((B, nx, ny)).astype(np.float32)
    scale = 0.5
    with np.errstate(invalid="ignore"):
        act0 = M.conv_first_model(u, v, _pack(lib, 0, w0, b0), scale)
        act1 = M.conv_mid_model(act0, _pack(lib, 1, w1, b1), B, nx, ny, rs, 64, True, grid=2)
        y = M.conv_mid_model(act1, _pack(lib, 2, w2, b2), B, nx, ny, rs, 16, False, grid=3)
    x = torch.tensor(np.stack((u, v), 1)).double() * scale
    fb = lambda b: torch.tensor(b).float().double()
    r0 = _q(torch.relu(F.conv2d(_q(x), _q(torch.tensor(w0)), fb(b0), padding=1)))
    r1 = _q(torch.relu(F.conv2d(r0, _q(torch.tensor(w1)), fb(b1), padding=1)))
    r2 = F.conv2d(r1, _q(torch.tensor(w2)), fb(b2), padding=1)
    a0 = M.unpack_activation(act0, B, nx, ny)
    a1 = M.unpack_activation(act1, B, nx, ny)
    # first layer: the bias travels through the GEMM as bf16 hi + lo parts (2^-17 relative): bf16 ulp flips only
    assert np.abs(a0[:, :5] - r0.numpy()).max() <= 2 ** -7 * np.abs(r0.numpy()).max() and not a0[:, 5:].any()
    assert np.abs(a0[:, :5] - r0.numpy()).mean() < 1e-6
    assert np.abs(a1[:, :7] - r1.numpy()).max() <= 2 ** -7 * np.abs(r1.numpy()).max()    # bf16 ulp flips only
    assert np.abs(a1[:, :7] - r1.numpy()).mean() < 1e-5 and not a1[:, 7:].any()
    assert np.abs(y[:, :2] - r2.numpy()).max() < 1e-2 * np.abs(r2.numpy()).max()
    assert not y[:, 2:].any()


def test_descriptor_decoding_roundtrip():
    smem = np.zeros(8192, dtype=np.uint8)
    vals = np.arange(128 * 16, dtype=np.float32).reshape(128, 16) / 8.0
    lbo, sbo, start = 2080, 128, 32
    for r in range(128):
        for k in range(16):
            addr = start + (r // 8) * sbo + (r % 8) * 16 + (k // 8) * lbo + (k % 8) * 2
            smem[addr:addr + 2] = M.bf16_bits(vals[r, k:k + 1]).view(np.uint8)
    got = M.gather_operand(smem, M.make_desc(start, lbo, sbo), 128)
    assert np.array_equal(got, M.bf16_round(vals))
