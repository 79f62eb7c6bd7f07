// Host-side check of ClusterLayout<N, C> (csrc/fpm_update_cluster.cuh): the shared-memory carve-up of
// fpm_update_cluster_kernel for the geometries the library runs it on -- 16-byte alignment of every array (8 bytes for the
// mbarriers), no overlaps, the window-slice buffers present exactly for the narrow instances, totals within the 227 KB a
// B200 CTA can opt into, the aliasing rule of W (|O_new|^2 of the slice lives in the idle row slab).
#include <cstdio>
#include <algorithm>
#include "fpm_update_cluster.cuh"
using namespace fpm;
template <int N, int C> static int check(const char* name, int r, int L, int cs, bool ws) {
  const int NR = 2 * r + 1, NC = 2 * r + 1, CPC = (NC + C - 1) / C;
  const ClusterLayout<N, C> lay(NR, NC, CPC, L, cs, ws);
  struct A { const char* n; size_t off, bytes, align; };
  const size_t slice8 = sizeof(float2) * (size_t)NR * CPC;
  const A arr[] = {
    {"rslab", lay.rslab, sizeof(float2) * (size_t)(N / C) * Shape<N>::PITCH, 16}, {"cslab", lay.cslab, sizeof(float2) * (size_t)CPC * (N + 1), 16},
    {"twA", lay.twA, sizeof(float2) * N, 16}, {"twB", lay.twB, sizeof(float2) * N, 16},
    {"Pc", lay.Pc, slice8, 16}, {"Qc", lay.Qc, slice8, 16}, {"Sc", lay.Sc, slice8 / 2, 16},
    {"Wb", lay.Wb, ws ? 2 * ((slice8 + 15) / 16 * 16) : 0, 16},
    {"U", lay.U, sizeof(float) * (size_t)lay.gro * (L >> 4), 16}, {"Tm", lay.Tm, sizeof(unsigned) * (size_t)lay.tmr * lay.tmc, 16},
    {"red", lay.red, sizeof(float) * 64, 16}, {"pmx", lay.pmx, sizeof(float) * 16, 16}, {"omx", lay.omx, sizeof(float) * 16, 16},
    {"bars", lay.bars, sizeof(uint64_t) * 5, 8},
  };
  const int n = (int)(sizeof arr / sizeof arr[0]);
  int bad = 0;
  for (int k = 0; k < n; ++k) {
    if (arr[k].off % arr[k].align) { printf("%s: %s not %zu-byte aligned\n", name, arr[k].n, arr[k].align); bad = 1; }
    const size_t end = arr[k].off + arr[k].bytes, next = (k + 1 < n) ? arr[k + 1].off : lay.total;
    if (end > next) { printf("%s: %s overlaps its successor (%zu > %zu)\n", name, arr[k].n, end, next); bad = 1; }
  }
  // touched cells: (NR >> cs) + 2 cell rows and (NC >> 4) + 2 cell columns are enough for any position of the rectangle
  for (int r0 = 0; r0 < (1 << cs); ++r0) if (((r0 + NR - 1) >> cs) + 1 > lay.tmr) { printf("%s: Tm rows\n", name); bad = 1; }
  for (int c0 = 0; c0 < 16; ++c0) if (((c0 + NC - 1) >> 4) + 1 > lay.tmc) { printf("%s: Tm columns\n", name); bad = 1; }
  if (sizeof(float) * (size_t)NR * CPC > arr[0].bytes) { printf("%s: W does not fit the row slab\n", name); bad = 1; }
  if (lay.total > 232448) { printf("%s: %zu bytes exceed the opt-in limit\n", name, lay.total); bad = 1; }
  if (CPC > 32) { printf("%s: more than one column per lane\n", name); bad = 1; }
  printf("%s: N=%d C=%d box %dx%d Nlarge %d cells %dx16 %s-> %zu bytes\n", name, N, C, NR, NC, L, 1 << cs, ws ? "window slice on chip " : "", lay.total);
  return bad;
}
int main() {
  int bad = 0;
  bad |= check<128, 4>("cfg2/cfg4 on 4 CTAs", 17, 384, 1, true);
  bad |= check<128, 2>("cfg2/cfg4 on 2 CTAs", 17, 384, 2, true);
  bad |= check<128, 4>("widest narrow box", 23, 512, 1, true);
  bad |= check<128, 2>("widest narrow box, 2 CTAs", 23, 512, 2, true);
  bad |= check<128, 4>("cfg5 (cellscope2, 85 x 85)", 42, 512, 1, false);
  bad |= check<128, 4>("121 x 121 box", 60, 384, 1, false);
  bad |= check<256, 8>("cfg5b (cellscope2, 256)", 84, 1024, 2, false);
  bad |= check<256, 8>("cfg3 (cellScope dome, 256)", 94, 1536, 4, false);
  return bad;
}
