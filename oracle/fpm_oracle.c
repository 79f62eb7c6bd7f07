/* oracle/fpm_oracle.c -- plain-C float64 restatement of the reference's runFPM() loop.
 * TEST INFRASTRUCTURE ONLY: built by oracle/Makefile into oracle/_build/liboracle_c.so and called
 * from tests/ (oracle/c_oracle.py) -- never by the product.  It exists because the numpy restatement
 * (oracle/fpm_oracle.py) needs minutes for BASELINE configs[2] / [4] at their full size (241 LEDs x 10
 * iterations at Np 256 / Nlarge 1536; 193 LEDs x 50 iterations): the full-size GPU parity tests use this one.
 *
 * Follows fpmMain.cpp:301-482 statement by statement in the windowed form of SURVEY.md appendix A
 * (centred spectrum, fftShift folded into index arithmetic), cvComplex semantics per SURVEY 8c
 * R1-R5.  Parity status: unpinned against the reference binary (cannot be built: cvComplex is not
 * vendored, fpmMain.cpp:15); pinned against oracle/fpm_oracle.py (and through it against
 * oracle/cv2_mirror.py, OpenCV's own cv::dft) by tests/test_oracle.py to 1e-13.
 *
 * Deliberate deviations from the letter of the numpy restatement, all below 1e-15 relative: |z| is
 * sqrt(re^2 + im^2) instead of hypot(); max|objF| (fpmMain.cpp:460,467: the FULL spectrum, every update) is
 * sqrt(max(re^2 + im^2)) taken over a cache of row maxima of which only the rows the window touched are rescanned
 * -- the same maximum as a full scan, without 2.4 M hypot calls per update at Nlarge 1536.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cplx;

static inline cplx cmul(cplx a, cplx b) { cplx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }

/* Stockham autosort FFT of length n = product of radices in {4, 2, 3, 5}; w[k] = exp(-2 pi i k / n).
 * sign = -1: forward (cv::dft), +1: inverse, unscaled.  x: n elements at `stride`; a, b: scratch [n]. */
static void fft1d(cplx* x, int n, int stride, const cplx* w, int sign, cplx* a, cplx* b) {
  for (int i = 0; i < n; ++i) a[i] = x[(size_t)i * stride];
  cplx *src = a, *dst = b;
  int len = n, s = 1;
  while (len > 1) {
    const int p = (len % 4 == 0) ? 4 : (len % 2 == 0) ? 2 : (len % 3 == 0) ? 3 : 5;
    const int m = len / p, tw = n / len;          /* exp(-2 pi i k / len) = w[k * tw]; r * k * tw < n always */
    for (int k = 0; k < m; ++k) {
      cplx t[5];
      for (int r = 1; r < p; ++r) { t[r] = w[(size_t)r * k * tw]; if (sign > 0) t[r].im = -t[r].im; }
      for (int q = 0; q < s; ++q) {
        cplx in[5], out[5];
        for (int j = 0; j < p; ++j) in[j] = src[q + s * (k + j * m)];
        if (p == 2) {
          out[0].re = in[0].re + in[1].re; out[0].im = in[0].im + in[1].im;
          out[1].re = in[0].re - in[1].re; out[1].im = in[0].im - in[1].im;
        } else if (p == 4) {                      /* W4 = -i forward, +i inverse */
          cplx s02 = {in[0].re + in[2].re, in[0].im + in[2].im}, d02 = {in[0].re - in[2].re, in[0].im - in[2].im};
          cplx s13 = {in[1].re + in[3].re, in[1].im + in[3].im}, d13 = {in[1].re - in[3].re, in[1].im - in[3].im};
          cplx jd = sign < 0 ? (cplx){d13.im, -d13.re} : (cplx){-d13.im, d13.re};      /* (-+i) * d13 */
          out[0].re = s02.re + s13.re; out[0].im = s02.im + s13.im;
          out[1].re = d02.re + jd.re;  out[1].im = d02.im + jd.im;
          out[2].re = s02.re - s13.re; out[2].im = s02.im - s13.im;
          out[3].re = d02.re - jd.re;  out[3].im = d02.im - jd.im;
        } else {
          for (int r = 0; r < p; ++r) {           /* p-point DFT, roots from the same table */
            cplx acc = in[0];
            for (int j = 1; j < p; ++j) {
              cplx root = w[(size_t)((j * r) % p) * (n / p)];
              if (sign > 0) root.im = -root.im;
              cplx u = cmul(in[j], root);
              acc.re += u.re; acc.im += u.im;
            }
            out[r] = acc;
          }
        }
        dst[q + s * (p * k)] = out[0];
        for (int r = 1; r < p; ++r) dst[q + s * (p * k + r)] = cmul(out[r], t[r]);
      }
    }
    cplx* tt = src; src = dst; dst = tt;
    len = m; s *= p;
  }
  for (int i = 0; i < n; ++i) x[(size_t)i * stride] = src[i];
}

static void fft2d(cplx* a, int n, const cplx* w, int sign, cplx* tmp /* 2n */) {
  for (int r = 0; r < n; ++r) fft1d(a + (size_t)r * n, n, 1, w, sign, tmp, tmp + n);
  for (int c = 0; c < n; ++c) fft1d(a + c, n, n, w, sign, tmp, tmp + n);
}

static void make_twiddles(cplx* w, int n) {
  for (int k = 0; k < n; ++k) { w[k].re = cos(-2.0 * M_PI * k / n); w[k].im = sin(-2.0 * M_PI * k / n); }
}

/* One sub-aperture update (fpmMain.cpp:358-475).  objFc: [L][L] centred; P, S: [N][N] DC-at-corner.
 * work: 3*N*N + 3*N elements; w: [N] twiddles. */
static void update(cplx* objFc, cplx* P, const double* S, const uint16_t* I, int N, int L, int xs, int ys,
                   double delta1, double delta2, double eps, int kappa, cplx* work, const cplx* w, double* rowmax2) {
  const int H = N / 2, NN = N * N;
  cplx *O = work, *Phi = work + NN, *psi = work + 2 * NN, *tmp = work + 3 * NN;
  double pmax2 = 0;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      cplx o = objFc[(size_t)(ys + (i + H) % N) * L + xs + (j + H) % N];                 /* :358-362 */
      cplx p = P[i * N + j];
      O[i * N + j] = o;
      Phi[i * N + j] = cmul(o, p);                                                        /* :364 */
      double pa2 = p.re * p.re + p.im * p.im;
      if (pa2 > pmax2) pmax2 = pa2;                                                       /* :415 */
    }
  const double pmax = sqrt(pmax2);
  memcpy(psi, Phi, sizeof(cplx) * NN);
  fft2d(psi, N, w, +1, tmp);                                                              /* :365 */
  for (int k = 0; k < NN; ++k) {
    double pr = psi[k].re / NN, pi_ = psi[k].im / NN;
    double tr = pr + eps, ti = pi_ + kappa * eps;                                         /* :390 */
    double mag = sqrt(tr * tr + ti * ti);                                                 /* :391 */
    double a = sqrt((double)I[k]);                                                        /* :378-387 */
    psi[k].re = a * pr / mag; psi[k].im = a * pi_ / mag;                                  /* :392-393 */
  }
  fft2d(psi, N, w, -1, tmp);                                                              /* :394 */
  /* object update :406-447 */
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      int k = i * N + j;
      cplx d = {psi[k].re - Phi[k].re, psi[k].im - Phi[k].im};                            /* :409 */
      cplx p = P[k];
      double pa = sqrt(p.re * p.re + p.im * p.im);
      cplx num = {(d.re * p.re + d.im * p.im) * pa, (d.im * p.re - d.re * p.im) * pa};    /* d*|P|*conj(P) */
      double A = pmax * (pa * pa + delta2), B = pmax * kappa * delta2, den = A * A + B * B;
      cplx dO = {(num.re * A + num.im * B) / den, (num.im * A - num.re * B) / den};
      cplx* dst = &objFc[(size_t)(ys + (i + H) % N) * L + xs + (j + H) % N];
      dst->re += dO.re; dst->im += dO.im;
      psi[k] = d;                                  /* keep dPhi */
    }
  /* :460,467 max|objF| over the FULL spectrum.  Only rows ys..ys+N-1 changed: their row maxima are recomputed, the
   * other rows' maxima are cached (exactly the same maximum as a full scan). */
  for (int r = ys; r < ys + N; ++r) {
    double m2 = 0;
    const cplx* row = objFc + (size_t)r * L;
    for (int k = 0; k < L; ++k) { double a2 = row[k].re * row[k].re + row[k].im * row[k].im; m2 = a2 > m2 ? a2 : m2; }
    rowmax2[r] = m2;
  }
  double omax2 = 0;
  for (int r = 0; r < L; ++r) omax2 = rowmax2[r] > omax2 ? rowmax2[r] : omax2;
  const double omax = sqrt(omax2);
  for (int k = 0; k < NN; ++k) {                                                          /* :459-475 */
    cplx o = O[k], d = psi[k];
    double oa = sqrt(o.re * o.re + o.im * o.im);
    cplx num = {(d.re * o.re + d.im * o.im) * oa, (d.im * o.re - d.re * o.im) * oa};
    double A = omax * (oa * oa + delta1), B = omax * kappa * delta1, den = A * A + B * B;
    P[k].re += (num.re * A + num.im * B) / den * S[k];
    P[k].im += (num.im * A - num.re * B) / den * S[k];
  }
}

/* `n_updates` consecutive updates starting at slot `slot_begin` (wrapping), on caller-initialised state.
 * Returns 0, or -1 when N has a prime factor other than 2, 3, 5. */
int fpm_oracle_run(cplx* objFc, cplx* P, const double* S, const uint16_t* stack, const int16_t* cx, const int16_t* cy,
                   int N, int L, int n_leds, int slot_begin, int n_updates, double delta1, double delta2, double eps,
                   int kappa) {
  int v = N;
  for (int f = 2; f <= 5; ++f) while (v % f == 0) v /= f;
  if (v != 1 || N < 2) return -1;
  cplx* work = (cplx*)malloc(sizeof(cplx) * ((size_t)3 * N * N + 3 * N));
  cplx* w = work + (size_t)3 * N * N + 2 * N;
  make_twiddles(w, N);
  double* rowmax2 = (double*)malloc(sizeof(double) * L);
  for (int r = 0; r < L; ++r) {
    double m2 = 0;
    for (int k = 0; k < L; ++k) { cplx z = objFc[(size_t)r * L + k]; double a2 = z.re * z.re + z.im * z.im; m2 = a2 > m2 ? a2 : m2; }
    rowmax2[r] = m2;
  }
  for (int u = 0; u < n_updates; ++u) {
    int k = (slot_begin + u) % n_leds;
    update(objFc, P, S, stack + (size_t)k * N * N, N, L, cx[k], cy[k], delta1, delta2, eps, kappa, work, w, rowmax2);
  }
  free(rowmax2);
  free(work);
  return 0;
}

/* 1-D transform exported for the unit test of the FFT itself (tests/test_oracle.py). */
int fpm_oracle_fft1d(cplx* x, int n, int sign) {
  cplx* t = (cplx*)malloc(sizeof(cplx) * 3 * (size_t)n);
  make_twiddles(t + 2 * n, n);
  fft1d(x, n, 1, t + 2 * n, sign, t, t + n);
  free(t);
  return 0;
}
