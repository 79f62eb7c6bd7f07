"""Index arithmetic of the cluster kernel's forwarded window slices (fpm_update_cluster.cuh, narrow boxes on 128 x 128
tiles), restated in numpy and checked on the golden crop tables of every BASELINE configuration with 128 x 128 tiles:

* the number of values a CTA expects on its `wbar` (computed from the two crop boxes) equals the number of values the
  other CTAs forward to it in C2 -- a mismatch would leave the barrier phase open for ever or complete it early;
* forwarded elements + the elements `next_slice` fetches from the spectrum cover every element of the next slice exactly
  once;
* the 16-bit reciprocals the kernel divides with are exact over the ranges it uses them on.

The update order of the LEDs wraps around at the end of an iteration (last LED -> first LED), which is the far jump with
nothing to forward.  No GPU involved: this pins the host-checkable part of the protocol; the GPU side is
tests/test_gpu_parity.py::test_cluster_kernel_repeated_runs / test_cluster_kernel_box_shapes.
"""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def crops(name):
    g = json.load(open(os.path.join(ROOT, "tests/golden/geometry_%s.json" % name)))
    byn = {l["n"]: l for l in g["leds"]}
    return [byn[n]["cropX"] for n in g["order"]], [byn[n]["cropY"] for n in g["order"]]


def slice_geometry(NC, C):
    cpc = (NC + C - 1) // C
    return cpc, [max(0, min(cpc, NC - k * cpc)) for k in range(C)]


@pytest.mark.parametrize("name,half", [("cfg2_fLEDc_np128", 17), ("cfg4_dogStomach_np128", 17), ("cfg5_cellscope2_np128", 23)])
@pytest.mark.parametrize("C", [2, 4])
def test_forwarded_counts_and_coverage(name, half, C):
    """half = half-width of a (square) pupil box; 23 is the widest box the narrow instances take."""
    N, H = 128, 64
    cx, cy = crops(name)
    ylo = xlo = -half
    NR = NC = 2 * half + 1
    cpc, ncls = slice_geometry(NC, C)
    n = len(cx)
    for u in range(n):
        v = (u + 1) % n
        r0, c0 = cy[u] + H + ylo, cx[u] + H + xlo
        r1, c1 = r0 + NR - 1, c0 + NC - 1
        r0n, c0n = cy[v] + H + ylo, cx[v] + H + xlo
        # receiver side (top of the update): ovr * ovc values expected
        expect = []
        for k in range(C):
            jc0, ncl = k * cpc, ncls[k]
            ovr = max(0, min(r1, r0n + NR - 1) - max(r0, r0n) + 1)
            ovc = max(0, min(c1, c0n + jc0 + ncl - 1) - max(c0, c0n + jc0) + 1) if ncl > 0 else 0
            expect.append(ovr * ovc)
        # sender side (C2): element (ir, jc) of this rectangle goes to next-slice (rn, cn - dk * cpc) of CTA dk
        cover = [np.zeros((NR, max(ncls[k], 1)), np.int32) for k in range(C)]
        sent = [0] * C
        cpc_inv = (65536 + cpc - 1) // cpc
        for ir in range(NR):
            rn = r0 + ir - r0n
            if not 0 <= rn < NR:
                continue
            for jc in range(NC):
                cn = c0 + jc - c0n
                if 0 <= cn < NC:
                    dk = (cn * cpc_inv) >> 16
                    assert dk == cn // cpc
                    lc = cn - dk * cpc
                    assert lc < ncls[dk]
                    sent[dk] += 1
                    cover[dk][rn, lc] += 1
        assert sent == expect, (name, C, u, sent, expect)
        # next_slice: the elements of the next slice outside this rectangle come from the spectrum
        for k in range(C):
            jc0, ncl = k * cpc, ncls[k]
            if ncl == 0:
                continue
            ncl_inv = (65536 + ncl - 1) // ncl
            dr, dc = r0n - r0, c0n + jc0 - c0
            for t in range(NR * ncl):
                rn = (t * ncl_inv) >> 16
                lc = t - rn * ncl
                assert rn == t // ncl
                inside = 0 <= rn + dr < NR and 0 <= lc + dc < NC
                if not inside:
                    cover[k][rn, lc] += 1
            assert np.all(cover[k][:, :ncl] == 1), (name, C, u, k)


def test_reciprocals_exact():
    """t / d by (t * ceil(65536 / d)) >> 16: the kernel uses it for d <= 32 with t < 47 * 32 (slice elements of a narrow
    box) and t < 256 (box columns)."""
    for d in range(1, 33):
        inv = (65536 + d - 1) // d
        t = np.arange(0, 2048, dtype=np.int64)
        assert np.array_equal((t * inv) >> 16, t // d), d
