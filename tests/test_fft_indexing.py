"""Numpy model of the Stockham stage formulas, the two-real-rows-per-complex-FFT packing and the radix
plans used by csrc/fft.cuh / csrc/solver_kernels.cu."""
import numpy as np
import pytest

PLANS = {64: (8, 8), 128: (8, 16), 256: (16, 16), 512: (8, 8, 8), 1024: (8, 8, 16), 2048: (8, 16, 16)}


def dft_regs(v, sign):
    r = len(v)
    bits = r.bit_length() - 1
    a = [None] * r
    for i in range(r):
        a[int(format(i, f"0{bits}b")[::-1], 2)] = v[i]
    ln = 2
    while ln <= r:
        for b in range(0, r, ln):
            for k in range(ln // 2):
                t = np.exp(sign * 2j * np.pi * k / ln) * a[b + k + ln // 2]
                a[b + k + ln // 2] = a[b + k] - t
                a[b + k] = a[b + k] + t
        ln *= 2
    return a


def stage(x, n, r, ns, sign, tw):
    out = np.zeros(n, complex)
    t = n // r
    for j in range(t):
        k = j % ns
        v = [x[j + q * t] for q in range(r)]
        for q in range(1, r):
            w = tw[q * k * (n // (ns * r))]
            v[q] = v[q] * (w if sign < 0 else np.conj(w))
        big = dft_regs(v, sign)
        j0 = (j // ns) * ns * r + k
        for q in range(r):
            out[j0 + q * ns] = big[q]
    return out


def fft(x, radices, sign):
    n = len(x)
    tw = np.exp(-2j * np.pi * np.arange(n) / n)
    ns = 1
    for r in radices:
        x = stage(x, n, r, ns, sign, tw)
        ns *= r
    return x


@pytest.mark.parametrize("n", sorted(PLANS))
def test_forward_and_reversed_inverse_plans(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    rad = PLANS[n]
    assert np.prod(rad) == n
    assert np.abs(fft(x, rad, -1) - np.fft.fft(x)).max() < 1e-11
    assert np.abs(fft(x, rad[::-1], +1) - np.fft.ifft(x) * n).max() < 1e-11      # K4 / K3 inverse order


def test_last_forward_stage_outputs_feed_first_inverse_stage():
    """K3 keeps the last forward stage's registers for the first inverse stage: thread j owns
    points j + q*T in both (T = N / R_last)."""
    n, r = 256, 16
    t = n // r
    for j in range(t):
        ns = n // r
        j0 = (j // ns) * ns * r + (j % ns)
        assert [j0 + q * ns for q in range(r)] == [j + q * t for q in range(r)]


def test_two_real_rows_in_one_complex_fft():
    n = 64
    rng = np.random.default_rng(3)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    z = np.fft.fft(a + 1j * b)
    k = np.arange(n // 2 + 1)
    zm = z[(n - k) % n]
    A = 0.5 * (z.real[k] + zm.real) + 0.5j * (z.imag[k] - zm.imag)          # K2 untangle
    B = 0.5 * (z.imag[k] + zm.imag) - 0.5j * (z.real[k] - zm.real)
    assert np.abs(A - np.fft.rfft(a)).max() < 1e-12 and np.abs(B - np.fft.rfft(b)).max() < 1e-12
    zf = np.zeros(n, complex)                                                 # K4 rebuild
    for kk in range(n // 2 + 1):
        if kk in (0, n // 2):
            zf[kk] = A[kk].real + 1j * B[kk].real
        else:
            zf[kk] = (A[kk].real - B[kk].imag) + 1j * (A[kk].imag + B[kk].real)
            zf[n - kk] = (A[kk].real + B[kk].imag) + 1j * (B[kk].real - A[kk].imag)
    back = np.fft.ifft(zf)
    assert np.abs(back.real - a).max() < 1e-12 and np.abs(back.imag - b).max() < 1e-12


def test_split_length_is_mirrored_in_slab_py():
    """slab.SlabGeometry.r0 must follow the rule csrc/plan.cu uses for the outer radix of the split x-transform
    (mlsim_slab_geometry is compared with it at run time on the GPU; this is the same check without one)."""
    import os
    import re
    from ml_accelerated_simulation_b200 import slab
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "ml_accelerated_simulation_b200", "csrc", "plan.cu")).read()
    n2 = int(re.search(r"#define MLSIM_SPLIT_N2 (\d+)", src).group(1))
    assert "cfg->nx > 2048 ? cfg->nx / MLSIM_SPLIT_N2 : 1" in src
    assert (slab.SPLIT_N2, slab.MAX_COL_FFT) == (n2, 2048)
    for nx, want in ((256, 1), (2048, 1), (4096, 4096 // n2), (8192, 8192 // n2)):
        assert slab.SlabGeometry(nx=nx, ny=64, batch=1, rank=0, world=1).r0 == want
