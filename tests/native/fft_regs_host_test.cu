// Host-side check of the in-register butterflies of csrc/fft_regs.cuh against a naive float64 DFT.
// Built and run by tests/test_fft_regs.py with `nvcc -x cu` (the functions are __host__ __device__).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "fft_regs.cuh"
using namespace fpm;

template <int R, bool INV> double check() {
  float2 v[R];
  double xr[R], xi[R];
  srand(7 * R + INV);
  for (int i = 0; i < R; ++i) {
    xr[i] = rand() / (double)RAND_MAX - 0.5; xi[i] = rand() / (double)RAND_MAX - 0.5;
    v[i] = make_float2((float)xr[i], (float)xi[i]);
  }
  if constexpr (R == 3) fft3<INV>(v[0], v[1], v[2]);
  else if constexpr (R == 5) fft5<INV>(v[0], v[1], v[2], v[3], v[4]);
  else if constexpr (R == 6 || R == 9 || R == 10 || R == 12 || R == 15 || R == 20 || R == 25) {
    float2 w[R];
    fft_reg<R, INV>(v);                                  // permuted output: slot i holds X[radix_out<R>(i)]
    for (int i = 0; i < R; ++i) w[radix_out<R>(i)] = v[i];
    for (int i = 0; i < R; ++i) v[i] = w[i];
  } else fftR<R, INV>(v);
  double err = 0, nrm = 0;
  for (int k = 0; k < R; ++k) {
    double sr = 0, si = 0;
    for (int n = 0; n < R; ++n) {
      double a = (INV ? 2.0 : -2.0) * M_PI * n * k / R;
      sr += xr[n] * cos(a) - xi[n] * sin(a); si += xr[n] * sin(a) + xi[n] * cos(a);
    }
    err += (v[k].x - sr) * (v[k].x - sr) + (v[k].y - si) * (v[k].y - si); nrm += sr * sr + si * si;
  }
  return sqrt(err / nrm);
}

// fft16_in6 against fft16 on the zero-padded input (must agree exactly) and twmul4 against cmul / cmulc
template <bool INV> int check_in6() {
  srand(99 + INV);
  int bad = 0;
  for (int rep = 0; rep < 100; ++rep) {
    float2 w[6], full[16], v[16];
    for (int i = 0; i < 16; ++i) full[i] = make_float2(0.f, 0.f);
    const int idx[6] = {0, 1, 2, 13, 14, 15};
    for (int i = 0; i < 6; ++i) {
      w[i] = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
      full[idx[i]] = w[i];
    }
    fft16<INV>(full);
    fft16_in6<INV>(w, v);
    for (int i = 0; i < 16; ++i) bad += !(v[i].x == full[i].x && v[i].y == full[i].y);
    const float2 a = w[0], t = w[1];
    const float2 f = twmul4(a, make_float4(t.x, t.y, -t.y, t.x)), g = cmul(a, t);
    // the packed device cmulc rounds a.y*b.x and fuses a.x*(-b.y) (the scalar host fallback fuses the other product)
    const float2 fc = twmul4(a, make_float4(t.x, -t.y, t.y, t.x));
    const float2 gc = make_float2(fmaf(a.x, t.x, a.y * t.y), fmaf(a.x, -t.y, a.y * t.x));
    bad += !(f.x == g.x && f.y == g.y && fc.x == gc.x && fc.y == gc.y);
  }
  return bad;
}

// fft_reg_in6<20> against fft_reg<20> of the zero-padded input, fft_reg_out6<20> against the six wanted outputs of
// fft_reg<20> (pruned butterflies of the Np = 200 plan): relative error at float rounding level
template <bool INV> int check_pruned20() {
  srand(5 + INV);
  int bad = 0;
  for (int rep = 0; rep < 100; ++rep) {
    float2 w[6], full[20], v[20];
    for (int i = 0; i < 20; ++i) full[i] = make_float2(0.f, 0.f);
    const int idx[6] = {0, 1, 2, 17, 18, 19};
    for (int i = 0; i < 6; ++i) {
      w[i] = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
      full[idx[i]] = w[i];
    }
    fft_reg<20, INV>(full);
    fft_reg_in6<20, INV>(w, v);
    for (int i = 0; i < 20; ++i) bad += !(fabsf(v[i].x - full[i].x) < 2e-6f && fabsf(v[i].y - full[i].y) < 2e-6f);
    float2 x[20], y[20], o[6];
    for (int i = 0; i < 20; ++i) x[i] = y[i] = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
    fft_reg<20, INV>(x);
    fft_reg_out6<20, INV>(y, o);
    for (int i = 0; i < 20; ++i)
      for (int k = 0; k < 6; ++k)
        if (radix_out<20>(i) == idx[k]) bad += !(fabsf(o[k].x - x[i].x) < 4e-6f && fabsf(o[k].y - x[i].y) < 4e-6f);
  }
  return bad;
}

int main() {
  double e[] = {check<2, false>(), check<2, true>(), check<3, false>(), check<3, true>(), check<4, false>(), check<4, true>(),
                check<5, false>(), check<5, true>(), check<8, false>(), check<8, true>(), check<16, false>(), check<16, true>(),
                check<32, false>(), check<32, true>(), check<6, false>(), check<6, true>(), check<9, false>(), check<9, true>(),
                check<10, false>(), check<10, true>(), check<12, false>(), check<12, true>(), check<15, false>(), check<15, true>(),
                check<20, false>(), check<20, true>(), check<25, false>(), check<25, true>()};
  const char* names[] = {"2f", "2i", "3f", "3i", "4f", "4i", "5f", "5i", "8f", "8i", "16f", "16i", "32f", "32i", "6f", "6i", "9f", "9i", "10f", "10i",
                         "12f", "12i", "15f", "15i", "20f", "20i", "25f", "25i"};
  int bad = check_in6<false>() + check_in6<true>();
  printf("in6/twmul4 mismatches %d\n", bad);
  const int bad20 = check_pruned20<false>() + check_pruned20<true>();
  printf("pruned radix-20 mismatches %d\n", bad20);
  bad += bad20;
  for (int i = 0; i < 28; ++i) { printf("%s %.3e\n", names[i], e[i]); bad += !(e[i] < 5e-7); }
  return bad;
}
