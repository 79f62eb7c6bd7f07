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
  else fftR<R, INV>(v);
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

int main() {
  double e[] = {check<2, false>(), check<2, true>(), check<3, false>(), check<3, true>(), check<4, false>(), check<4, true>(),
                check<5, false>(), check<5, true>(), check<8, false>(), check<8, true>(), check<16, false>(), check<16, true>(),
                check<32, false>(), check<32, true>()};
  const char* names[] = {"2f", "2i", "3f", "3i", "4f", "4i", "5f", "5i", "8f", "8i", "16f", "16i", "32f", "32i"};
  int bad = 0;
  for (int i = 0; i < 14; ++i) { printf("%s %.3e\n", names[i], e[i]); bad += !(e[i] < 5e-7); }
  return bad;
}
