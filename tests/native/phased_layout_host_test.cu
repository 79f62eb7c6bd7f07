// Host-side check of PhasedShape<N>::layout (csrc/fpm_update_phased.cuh): the shared-memory carve-up of
// fpm_update_phased_kernel for the shipped geometries -- alignment of the TMA destinations, no overlaps, totals within
// the 227 KB a B200 CTA can opt into.
#include <cstdio>
#include "fpm_update_phased.cuh"
using namespace fpm;
template <int N> static int check(const char* name, int r, int L) {
  using PS = PhasedShape<N>;
  const int NR = 2 * r + 1, NC = 2 * r + 1, ocp = (NC + 2) & ~1;
  int off[PS::NOFF];
  PS::layout(NR, NC, ocp, L, off);
  int bad = 0;
  if (!PS::box_ok(-r, r, -r, r)) { printf("%s: box rejected\n", name); bad = 1; }
  if (off[PS::WIN0] % 128 || off[PS::WIN1] % 128) { printf("%s: window not 128-byte aligned\n", name); bad = 1; }
  const size_t win = sizeof(float2) * (size_t)NR * ocp;
  if ((size_t)(off[PS::WIN1] - off[PS::WIN0]) < win || (size_t)(off[PS::PC] - off[PS::WIN1]) < win) { printf("%s: windows overlap\n", name); bad = 1; }
  if ((size_t)(off[PS::QC] - off[PS::PC]) < win || (size_t)(off[PS::SC] - off[PS::QC]) < win) { printf("%s: P/Q overlap\n", name); bad = 1; }
  if ((size_t)(off[PS::WPIX] - off[PS::SC]) < win / 2) { printf("%s: S overlaps W\n", name); bad = 1; }
  const int tmc = (NC >> 4) + 2;
  int wsh = 0; while ((1 << wsh) < (tmc << 4)) ++wsh;
  if ((size_t)(off[PS::UCELL] - off[PS::WPIX]) < sizeof(float) * ((size_t)NR << wsh)) { printf("%s: W overlaps U\n", name); bad = 1; }
  if ((size_t)(off[PS::RMAX] - off[PS::UCELL]) < sizeof(float) * (size_t)L * (L >> 4)) { printf("%s: U overlaps Rm\n", name); bad = 1; }
  if ((size_t)(off[PS::TOTAL] - off[PS::RMAX]) < sizeof(float) * (size_t)L) { printf("%s: Rm cut\n", name); bad = 1; }
  for (int k = 0; k < PS::TOTAL; ++k) if (off[k] % 16) { printf("%s: offset %d not 16-byte aligned\n", name, k); bad = 1; }
  if (off[PS::TOTAL] > 232448) { printf("%s: %d bytes exceed the opt-in limit\n", name, off[PS::TOTAL]); bad = 1; }
  printf("%s: N=%d box %dx%d Nlarge %d -> %d bytes\n", name, N, NR, NC, L, off[PS::TOTAL]);
  return bad;
}
int main() {
  int bad = 0;
  bad |= check<128>("cfg2/cfg4 (fLED-c, dogStomach)", 17, 384);
  bad |= check<64>("cfg1/cfg6 (mono)", 21, 256);
  bad |= check<64>("cfg3b (cellScope, 64)", 24, 384);
  bad |= !(!PhasedShape<128>::box_ok(-42, 42, -42, 42));          // cfg5's 85 x 85 box is not narrow
  return bad;
}
