/* oracle/fpm_oracle.c -- plain-C float64 restatement of the reference's runFPM() loop.
 * TEST INFRASTRUCTURE ONLY: built by oracle/Makefile into oracle/_build/liboracle_c.so and called
 * from tests/ and bench.py (CPU baseline "best case") -- never by the product.
 *
 * Follows fpmMain.cpp:301-482 statement by statement in the windowed form of SURVEY.md appendix A
 * (centred spectrum, fftShift folded into index arithmetic), cvComplex semantics per SURVEY 8c
 * R1-R5.  Parity status: unpinned against the reference binary (cannot be built: cvComplex is not
 * vendored, fpmMain.cpp:15); pinned against oracle/cv2_mirror.py (OpenCV's own cv::dft) by
 * tests/test_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cplx;

/* iterative radix-2 FFT, n power of two; sign=-1 forward (cv::dft), +1 inverse (unscaled) */
static void fft1d(cplx* a, int n, int stride, int sign, cplx* tmp) {
  for (int i = 0; i < n; ++i) tmp[i] = a[(size_t)i * stride];
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { cplx t = tmp[i]; tmp[i] = tmp[j]; tmp[j] = t; }
  }
  for (int len = 2; len <= n; len <<= 1) {
    double ang = sign * 2.0 * M_PI / len;
    for (int i = 0; i < n; i += len)
      for (int k = 0; k < len / 2; ++k) {
        double wr = cos(ang * k), wi = sin(ang * k);
        cplx u = tmp[i + k], v = tmp[i + k + len / 2];
        double xr = v.re * wr - v.im * wi, xi = v.re * wi + v.im * wr;
        tmp[i + k].re = u.re + xr; tmp[i + k].im = u.im + xi;
        tmp[i + k + len / 2].re = u.re - xr; tmp[i + k + len / 2].im = u.im - xi;
      }
  }
  for (int i = 0; i < n; ++i) a[(size_t)i * stride] = tmp[i];
}

static void fft2d(cplx* a, int n, int sign, cplx* tmp) {
  for (int r = 0; r < n; ++r) fft1d(a + (size_t)r * n, n, 1, sign, tmp);
  for (int c = 0; c < n; ++c) fft1d(a + c, n, n, sign, tmp);
}

/* One sub-aperture update (fpmMain.cpp:358-475).  objFc: [L][L] centred; P, S: [N][N] DC-at-corner. */
void fpm_oracle_update(cplx* objFc, cplx* P, const double* S, const uint16_t* I, int N, int L, int xs, int ys,
                       double delta1, double delta2, double eps, int kappa, cplx* work /* 4*N*N + N */) {
  const int H = N / 2, NN = N * N;
  cplx *O = work, *Phi = work + NN, *psi = work + 2 * NN, *tmp = work + 3 * NN;
  double pmax = 0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      cplx o = objFc[(size_t)(ys + (i + H) % N) * L + xs + (j + H) % N];                 /* :358-362 */
      cplx p = P[i * N + j];
      O[i * N + j] = o;
      Phi[i * N + j].re = o.re * p.re - o.im * p.im;                                      /* :364 */
      Phi[i * N + j].im = o.re * p.im + o.im * p.re;
      double pa = hypot(p.re, p.im);
      if (pa > pmax) pmax = pa;                                                           /* :415 */
    }
  memcpy(psi, Phi, sizeof(cplx) * NN);
  fft2d(psi, N, +1, tmp);                                                                 /* :365 */
  for (int k = 0; k < NN; ++k) {
    double pr = psi[k].re / NN, pi_ = psi[k].im / NN;
    double mag = hypot(pr + eps, pi_ + kappa * eps);                                      /* :390-391 */
    double a = sqrt((double)I[k]);                                                        /* :378-387 */
    psi[k].re = a * pr / mag; psi[k].im = a * pi_ / mag;                                  /* :392-393 */
  }
  fft2d(psi, N, -1, tmp);                                                                 /* :394 */
  /* object update :406-447 */
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      int k = i * N + j;
      cplx d = {psi[k].re - Phi[k].re, psi[k].im - Phi[k].im};                            /* :409 */
      cplx p = P[k];
      double pa = hypot(p.re, p.im);
      cplx num = {(d.re * p.re + d.im * p.im) * pa, (d.im * p.re - d.re * p.im) * pa};    /* d*|P|*conj(P) */
      double A = pmax * (pa * pa + delta2), B = pmax * kappa * delta2, den = A * A + B * B;
      cplx dO = {(num.re * A + num.im * B) / den, (num.im * A - num.re * B) / den};
      cplx* dst = &objFc[(size_t)(ys + (i + H) % N) * L + xs + (j + H) % N];
      dst->re += dO.re; dst->im += dO.im;
      psi[k] = d;                                  /* keep dPhi */
    }
  double omax = 0;                                                                        /* :460,467 */
  for (size_t k = 0; k < (size_t)L * L; ++k) { double a = hypot(objFc[k].re, objFc[k].im); if (a > omax) omax = a; }
  for (int k = 0; k < NN; ++k) {                                                          /* :459-475 */
    cplx o = O[k], d = psi[k];
    double oa = hypot(o.re, o.im);
    cplx num = {(d.re * o.re + d.im * o.im) * oa, (d.im * o.re - d.re * o.im) * oa};
    double A = omax * (oa * oa + delta1), B = omax * kappa * delta1, den = A * A + B * B;
    P[k].re += (num.re * A + num.im * B) / den * S[k];
    P[k].im += (num.im * A - num.re * B) / den * S[k];
  }
}

/* `n_updates` consecutive updates starting at slot 0 (wrapping), on caller-initialised state. */
void fpm_oracle_run(cplx* objFc, cplx* P, const double* S, const uint16_t* stack, const int16_t* cx, const int16_t* cy,
                    int N, int L, int n_leds, int n_updates, double delta1, double delta2, double eps, int kappa) {
  cplx* work = (cplx*)malloc(sizeof(cplx) * ((size_t)4 * N * N + N));
  for (int u = 0; u < n_updates; ++u) {
    int k = u % n_leds;
    fpm_oracle_update(objFc, P, S, stack + (size_t)k * N * N, N, L, cx[k], cy[k], delta1, delta2, eps, kappa, work);
  }
  free(work);
}
