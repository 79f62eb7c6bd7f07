"""Index arithmetic of the round-2 kernels restated in numpy and checked against numpy.fft (CPU suite):

* plan_fft_kernel (csrc/fpm_fft2d.cuh): three in-place decimation-in-frequency stages of an R0*R1*R2 line with one pad
  element per R0-block; position k0*(M0+1) + k1*M1 + k2 ends up holding X[k0 + R0*k1 + R0*R1*k2]; the row pass reads its
  source through the fftShift;
* phase C of fpm_update_phased_kernel (csrc/fpm_update_phased.cuh): the last forward column stage as a direct R1-term
  sum of the stage-B' outputs with the half-period sign folded."""
import numpy as np
import pytest

PLANS = [(16, 16, 1), (10, 6, 6), (8, 8, 6), (8, 8, 8), (10, 10, 6), (16, 8, 8), (16, 16, 6), (8, 8, 1), (16, 8, 1)]   # fpmb200.cu


def dft(v, sgn):
    n = len(v)
    k = np.arange(n)
    return np.array([np.sum(v * np.exp(sgn * 2j * np.pi * k * q / n)) for q in range(n)])


def planned_line(x, R0, R1, R2, inverse):
    L, M0 = R0 * R1 * R2, R1 * R2
    M1 = M0 // R1
    sgn = 1.0 if inverse else -1.0
    w = lambda e: np.exp(sgn * 2j * np.pi * e / L)
    buf = np.zeros(L + R0, np.complex128)
    n = np.arange(L)
    buf[n + n // M0] = x                                               # element n of the line sits at n + n / M0
    for m in range(M0):                                                # stage 0
        pos = m + np.arange(R0) * (M0 + 1)
        V = dft(buf[pos], sgn)
        buf[pos] = V * w(m * np.arange(R0))
    for b in range(R0):                                                # stage 1
        for m in range(M1):
            pos = b * (M0 + 1) + m + np.arange(R1) * M1
            V = dft(buf[pos], sgn)
            buf[pos] = V * (w(R0 * m * np.arange(R1)) if R2 > 1 else 1.0)
    if R2 > 1:                                                         # stage 2
        for b in range(R0):
            for k1 in range(R1):
                pos = b * (M0 + 1) + k1 * M1 + np.arange(R2)
                buf[pos] = dft(buf[pos], sgn)
    k = np.arange(L)
    return buf[(k % R0) * (M0 + 1) + ((k // R0) % R1) * M1 + k // (R0 * R1)]


@pytest.mark.parametrize("R0,R1,R2", PLANS)
def test_planned_line_is_the_dft(R0, R1, R2):
    L = R0 * R1 * R2
    rng = np.random.default_rng(L)
    x = rng.standard_normal(L) + 1j * rng.standard_normal(L)
    assert np.allclose(planned_line(x, R0, R1, R2, False), np.fft.fft(x), rtol=0, atol=1e-9)
    assert np.allclose(planned_line(x, R0, R1, R2, True), np.fft.ifft(x) * L, rtol=0, atol=1e-9)


def test_row_pass_reads_through_the_fftshift():
    """dst[r][c] = src[(r + h) % L][(c + h) % L] is the centred -> DC-at-corner layout change of fpmMain.cpp:358."""
    L = 12
    h = L // 2
    src = np.arange(L * L).reshape(L, L)
    r, c = np.mgrid[0:L, 0:L]
    assert np.array_equal(src[(r + h) % L, (c + h) % L], np.fft.ifftshift(src))


@pytest.mark.parametrize("R1", [8, 16])
def test_direct_last_column_stage(R1):
    """Forward transform of length N = R1 * 8 split as in the update kernels: stage B' (8-point transforms of the
    scrambled input, twiddled) then stage A' over k1 -- output i = 8 r + q equals sum_k1 B[8 k1 + q] W_R1^(r k1), and the
    direct sum with the half-period sign folded equals the butterfly."""
    R2, N = 8, R1 * 8
    rng = np.random.default_rng(R1)
    B = rng.standard_normal(N) + 1j * rng.standard_normal(N)          # stage-B' outputs (twiddles applied), index 8*k1 + q
    W = np.exp(-2j * np.pi / R1)
    full = np.zeros(N, np.complex128)
    for q in range(R2):
        v = B[R2 * np.arange(R1) + q]
        full[R2 * np.arange(R1) + q] = dft(v, -1.0)                     # X[8 r + q] for r = 0 .. R1-1
    TK = R1 // 2
    for i in rng.integers(0, N, 40):
        q, r = i % R2, i // R2
        sg = -1.0 if r & 1 else 1.0
        acc = sum((B[R2 * k + q] + sg * B[R2 * (k + TK) + q]) * W ** ((r * k) % R1) for k in range(TK))
        assert abs(acc - full[i]) < 1e-10
