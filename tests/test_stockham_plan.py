"""The index arithmetic of the two-stage Stockham plans of fpm_update_general_kernel (csrc/fpm_general_fused.cuh,
plan_stage), restated in numpy and checked against numpy.fft for every compiled (R1, R2): stage 1 = radix R1 without
twiddles, outputs contiguous (j*R1 + k); stage 2 = radix R2, inputs twiddled by W_N^(r*j), outputs at j + k*R1.
Also the "known-zero samples are not read" rule of the pruned stages (a zero-padded line transforms to the same
values whether the zeros are loaded or substituted)."""
import numpy as np
import pytest

PLANS = [(10, 9), (10, 10), (10, 8), (9, 8), (16, 6), (10, 6)]        # general_fused_plan() in csrc/fpmb200.cu


def two_stage(x, R1, R2, inverse=False, keep=None):
    N = R1 * R2
    sgn = 1.0 if inverse else -1.0
    src = x.astype(np.complex128).copy()
    if keep is not None:                                              # ZIN: samples outside `keep` substituted by zero
        src = np.where(keep, src, 0)
    mid = np.zeros(N, np.complex128)
    T = R2                                                            # stage 1: T = N / R1 work items
    for j in range(T):
        v = src[j + np.arange(R1) * T]
        V = np.array([np.sum(v * np.exp(sgn * 2j * np.pi * np.arange(R1) * k / R1)) for k in range(R1)])
        mid[j * R1 + np.arange(R1)] = V
    out = np.zeros(N, np.complex128)
    T = R1                                                            # stage 2: T = N / R2 work items
    for j in range(T):
        r = np.arange(R2)
        v = mid[j + r * T] * np.exp(sgn * 2j * np.pi * r * j / N)
        V = np.array([np.sum(v * np.exp(sgn * 2j * np.pi * r * k / R2)) for k in range(R2)])
        out[j + np.arange(R2) * T] = V
    return out


@pytest.mark.parametrize("R1,R2", PLANS)
def test_two_stage_plan_is_the_dft(R1, R2):
    rng = np.random.default_rng(R1 * 100 + R2)
    N = R1 * R2
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    assert np.allclose(two_stage(x, R1, R2), np.fft.fft(x), rtol=0, atol=1e-10)
    assert np.allclose(two_stage(x, R1, R2, inverse=True), np.fft.ifft(x) * N, rtol=0, atol=1e-10)


@pytest.mark.parametrize("R1,R2", PLANS[:2])
def test_known_zero_samples_need_not_be_read(R1, R2):
    """Wrapped box [lo, lo+n): stale values outside it do not matter when the stage substitutes zeros."""
    rng = np.random.default_rng(7)
    N = R1 * R2
    idx = np.arange(N)
    wrapped = np.where(idx < N // 2, idx, idx - N)
    for lo, n in ((-30, 61), (5, 20), (-N // 2, N)):
        keep = (wrapped >= lo) & (wrapped < lo + n)
        clean = np.where(keep, rng.standard_normal(N) + 1j * rng.standard_normal(N), 0)
        stale = np.where(keep, clean, 1e30)
        assert np.allclose(two_stage(stale, R1, R2, inverse=True, keep=keep), np.fft.ifft(clean) * N, rtol=0, atol=1e-9)
