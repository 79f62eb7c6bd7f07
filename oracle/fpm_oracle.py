"""CPU oracle for the FPM reconstruction path -- TEST INFRASTRUCTURE ONLY.

This file is a float64 restatement of the reference's `runFPM()` loop and of the
integer LED-geometry path of `loadFPMDataset()` / `main()`.  It is the *checker*:
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may import it.  The product (libfpmb200.so, fpmMain)
never calls into `oracle/`.

Parity status: **unpinned for the arithmetic** -- the reference ships no tests, no
golden vectors and cannot be compiled here (`cvComplex` is not vendored,
fpmMain.cpp:15, makefile:23; no OpenCV C++ headers).  What *is* pinned:
  * the JSON reading + LED geometry + LED order, against `oracle/_ref/ref_geometry`
    (reference's own vendored jsoncpp compiled in place + libstdc++ std::sort);
    fixtures in tests/golden/geometry_*.json;
  * K1/K2 of the reference's profile (output.svg): 157 LEDs pass the NA filter for
    dogStomach, Np=200 -> Nlarge=600;
  * this windowed numpy restatement against a 1:1 op-sequence mirror of
    fpmMain.cpp:345-482 executed on OpenCV's own cv::dft/arithm via `cv2`
    (oracle/cv2_mirror.py), fixtures in tests/golden/loop_*.npz.

Storage convention: `objFc` is the *centred* spectrum; the
reference stores DC-at-corner and calls fftShift around every access
(fpmMain.cpp:358,427,447), so the two are equivalent.  `P`, `S` are DC-at-corner.
"""
from __future__ import annotations

import json
import math
import re
from dataclasses import dataclass, field

import numpy as np

f32 = np.float32


# --------------------------------------------------------------------------- #
# JSON reading with jsoncpp-1.6.5 accessor semantics (include/jsoncpp.cpp)
# --------------------------------------------------------------------------- #
def load_json_lenient(path: str) -> dict:
    """Parse a dataset JSON the way the reference's `Json::Reader` ends up seeing it.

    jsoncpp recovers from the trailing comma that ends `holeCoordinates` in
    dataset_cellScope.json / dataset_dogStomach.json (include/jsoncpp.cpp:698-734,
    941-953): parse() returns false (ignored at fpmMain.cpp:515) and the array gets
    one extra null element.  An extra null behaves exactly like an out-of-range
    index in `led_coords()` below, so dropping the comma is equivalent.
    """
    txt = open(path).read()
    txt = re.sub(r",(\s*[\]}])", r"\1", txt)
    return json.loads(txt)


def _as_int(v) -> int:
    # Value::asInt(): realValue -> int(value) truncation (jsoncpp.cpp:3056-3059)
    if isinstance(v, bool):
        return int(v)
    if isinstance(v, float):
        return int(v)  # trunc toward zero
    return int(v)


def _c_round(x: float) -> float:
    """C `round()`: half away from zero."""
    return math.copysign(math.floor(abs(x) + 0.5), x)


# --------------------------------------------------------------------------- #
# derived optics parameters  (fpmMain.cpp:517-575, 305-306)
# --------------------------------------------------------------------------- #
@dataclass
class Config:
    Np: int
    pixelSize: f32
    objectiveMag: f32
    objectiveNA: f32
    maxIlluminationNA: f32
    lam: f32
    arrayRotation: float
    delta1: f32
    delta2: f32
    ledCount: int
    flipX: bool
    flipY: bool
    ps_eff: f32 = f32(0)
    du: f32 = f32(0)
    factor: int = 0
    Nlarge: int = 0
    naRadius: int = 0
    eps: f32 = f32(0.0000000001)  # fpmMain.h:99


def config_from_json(j: dict, Np_override: int | None = None) -> Config:
    """fpmMain.cpp:519-575 with the reference's C types."""
    Np = _as_int(j.get("cropSizeX", 90)) if Np_override is None else int(Np_override)
    c = Config(
        Np=Np,
        pixelSize=f32(float(j.get("pixelSize", 6.5))),
        objectiveMag=f32(float(j.get("objectiveMag", 8))),
        objectiveNA=f32(float(j.get("objectiveNA", 0.2))),
        maxIlluminationNA=f32(float(j.get("maxIlluminationNA", 0.7604))),
        lam=f32(float(j.get("lambda", 0.5))),
        arrayRotation=float(_as_int(j.get("arrayRotation", 0))),
        delta1=f32(_as_int(j.get("delta1", 5))),
        delta2=f32(_as_int(j.get("delta2", 10))),
        ledCount=_as_int(j.get("ledCount", 508)),
        flipX=bool(j.get("flipDatasetX", False)),
        flipY=bool(j.get("flipDatasetY", False)),
    )
    c.ps_eff = f32(c.pixelSize / c.objectiveMag)                     # :529
    c.du = f32(f32(f32(1) / c.ps_eff) / f32(Np))                     # :530
    t = f32(f32(f32(f32(2) * c.ps_eff) * f32(c.maxIlluminationNA + c.objectiveNA)) / c.lam)
    c.factor = 1 + int(math.ceil(float(t)))                          # :556-558
    c.Nlarge = Np * c.factor                                         # :564
    r = f32(f32(f32(c.objectiveNA * c.ps_eff) * f32(Np)) / c.lam)
    c.naRadius = int(math.ceil(float(r)))                            # :305-306
    return c


def led_coords(j: dict, n: int):
    """`holeCoordinates[n-1][k].get(axis,0).asFloat()` (fpmMain.cpp:77-79).

    Non-const `Value::operator[]` auto-creates nulls, so an out-of-range / null row
    yields (0,0,0).
    """
    hc = j.get("holeCoordinates", 0)
    if not isinstance(hc, list):
        raise ValueError("holeCoordinates is not an array (Json::LogicError in the reference)")
    i = n - 1
    if i < 0 or i >= len(hc) or hc[i] is None:
        return f32(0), f32(0), f32(0)
    row = hc[i]
    out = []
    for k, ax in enumerate("xyz"):
        e = row[k] if k < len(row) and row[k] is not None else {}
        out.append(f32(float(e.get(ax, 0))))
    return tuple(out)


@dataclass
class Geometry:
    led_nums: list          # file numbers that pass the NA filter (any order)
    na: dict                # led -> float32 illumination NA
    idx_u: dict
    idx_v: dict
    cropX: dict
    cropY: dict
    order: list             # sortedIndicies (first ledUsedCount entries)
    na_list: np.ndarray = field(default=None)


def led_geometry(cfg: Config, j: dict, present_leds) -> Geometry:
    """fpmMain.cpp:60-61,77-106,146-168,238-258 for the LED numbers in `present_leds`
    (= the numbers for which an image file exists)."""
    ang = cfg.arrayRotation
    c_, s_ = math.cos(ang * math.pi / 180), math.sin(ang * math.pi / 180)
    R = [[c_, -s_, 0.0], [s_, c_, 0.0], [0.0, 0.0, 1.0]]
    na_list = np.full(cfg.ledCount + 1, f32(99.0), dtype=np.float32)
    g = Geometry([], {}, {}, {}, {}, {}, [])
    for n in present_leds:
        px, py, pz = (float(v) for v in led_coords(j, n))
        # 1x3 (double) * 3x3 (double): cv::gemm accumulates k = 0,1,2 in order
        v = [px * R[0][k] + py * R[1][k] + pz * R[2][k] for k in range(3)]
        if cfg.flipX:
            fl = (-1.0, 1.0, 1.0)
        else:
            fl = (1.0, 1.0, 1.0)
        if cfg.flipY:                       # Y overrides X (fpmMain.cpp:89-92)
            fl = (1.0, -1.0, 1.0)
        v = [v[k] * fl[k] for k in range(3)]
        sx = math.sin(math.atan2(v[0], v[2]))
        sy = math.sin(math.atan2(v[1], v[2]))
        na = f32(math.sqrt(sx * sx + sy * sy))
        if not (na < cfg.maxIlluminationNA):
            continue
        if n > cfg.ledCount or n < 0:
            raise IndexError("imageStack.at(led_num) out of range in the reference")
        uled = f32(sx / float(cfg.lam))
        vled = f32(sy / float(cfg.lam))
        iu = int(_c_round(float(f32(uled / cfg.du))))
        iv = int(_c_round(float(f32(vled / cfg.du))))
        g.led_nums.append(n)
        g.na[n] = na
        g.idx_u[n], g.idx_v[n] = iu, iv
        g.cropX[n] = cfg.Nlarge // 2 + iu - cfg.Np // 2
        g.cropY[n] = cfg.Nlarge // 2 + iv - cfg.Np // 2
        na_list[n] = na
    g.na_list = na_list
    order = libstdcxx_sort_indexes(na_list)
    g.order = [int(i) for i in order[: len(g.led_nums)]]
    return g


# --------------------------------------------------------------------------- #
# libstdc++ std::sort restated (bits/stl_algo.h: __introsort_loop, threshold 16)
# so the *unstable* tie order of fpmMain.h:103-115 is reproduced exactly.
# --------------------------------------------------------------------------- #
def libstdcxx_sort_indexes(v) -> list:
    v = [float(x) for x in v]
    a = list(range(len(v)))
    comp = lambda i, j: v[i] < v[j]
    n = len(a)
    if n == 0:
        return a
    S_THRESHOLD = 16

    def move_median_to_first(res, ia, ib, ic):
        if comp(a[ia], a[ib]):
            if comp(a[ib], a[ic]):
                a[res], a[ib] = a[ib], a[res]
            elif comp(a[ia], a[ic]):
                a[res], a[ic] = a[ic], a[res]
            else:
                a[res], a[ia] = a[ia], a[res]
        elif comp(a[ia], a[ic]):
            a[res], a[ia] = a[ia], a[res]
        elif comp(a[ib], a[ic]):
            a[res], a[ic] = a[ic], a[res]
        else:
            a[res], a[ib] = a[ib], a[res]

    def unguarded_partition(first, last, pivot):
        while True:
            while comp(a[first], a[pivot]):
                first += 1
            last -= 1
            while comp(a[pivot], a[last]):
                last -= 1
            if not (first < last):
                return first
            a[first], a[last] = a[last], a[first]
            first += 1

    def heap_sort(first, last):  # std::__partial_sort(first,last,last) == heapsort
        import heapq  # noqa: F401  (not equivalent in tie order; never reached for our sizes)
        raise NotImplementedError("introsort depth limit hit; restate __heap_select first")

    def introsort_loop(first, last, depth):
        while last - first > S_THRESHOLD:
            if depth == 0:
                heap_sort(first, last)
                return
            depth -= 1
            mid = first + (last - first) // 2
            move_median_to_first(first, first + 1, mid, last - 1)
            cut = unguarded_partition(first + 1, last, first)
            introsort_loop(cut, last, depth)
            last = cut

    def unguarded_linear_insert(last):
        val = a[last]
        nxt = last - 1
        while comp(val, a[nxt]):
            a[last] = a[nxt]
            last = nxt
            nxt -= 1
        a[last] = val

    def insertion_sort(first, last):
        if first == last:
            return
        for i in range(first + 1, last):
            if comp(a[i], a[first]):
                val = a[i]
                a[first + 1 : i + 1] = a[first:i]
                a[first] = val
            else:
                unguarded_linear_insert(i)

    introsort_loop(0, n, 2 * (n.bit_length() - 1))
    if n > S_THRESHOLD:
        insertion_sort(0, S_THRESHOLD)
        for i in range(S_THRESHOLD, n):
            unguarded_linear_insert(i)
    else:
        insertion_sort(0, n)
    return a


# --------------------------------------------------------------------------- #
# pupil support (synthetic stacks live in fpm-opencv_b200/synth.py: input plumbing, not oracle)
# --------------------------------------------------------------------------- #
def sh(a: np.ndarray) -> np.ndarray:
    """cvComplex fftShift on even sizes: circular shift by half (R2)."""
    return np.roll(a, (a.shape[0] // 2, a.shape[1] // 2), axis=(0, 1))


def pupil_support(N: int, r: int) -> np.ndarray:
    """fpmMain.cpp:304-310: filled cv::circle centre (N/2,N/2) radius r, then fftShift.
    cv::circle(filled) == (dx^2+dy^2 <= r^2) for r=1..199 in cv2 4.13;
    tests/test_oracle.py re-checks that against cv2.circle."""
    y, x = np.mgrid[0:N, 0:N]
    disc = ((x - N // 2) ** 2 + (y - N // 2) ** 2 <= r * r).astype(np.float64)
    return sh(disc)


# --------------------------------------------------------------------------- #
# the loop  (fpmMain.cpp:301-482), windowed float64 restatement
# --------------------------------------------------------------------------- #
@dataclass
class State:
    objFc: np.ndarray   # [L][L] complex128, centred
    P: np.ndarray       # [N][N] complex128, DC-at-corner
    S: np.ndarray       # [N][N] float64 support, DC-at-corner


def init_state(stack: np.ndarray, L: int, naRadius: int, init_slot: int = 1) -> State:
    """fpmMain.cpp:301-343.  `stack` is in update order; init image = slot 1
    (`sortedIndicies.at(1)`, fpmMain.cpp:319)."""
    N = stack.shape[1]
    S = pupil_support(N, naRadius)
    c = np.fft.fft2(np.sqrt(stack[init_slot].astype(np.float64))) * S
    objFc = np.zeros((L, L), dtype=np.complex128)
    o = L // 2 - N // 2
    objFc[o : o + N, o : o + N] = sh(c)
    return State(objFc, S.astype(np.complex128), S)


def update(st: State, I: np.ndarray, xs: int, ys: int, delta1, delta2, eps, kappa=1) -> None:
    """One sub-aperture update, fpmMain.cpp:358-475 (appendix A of SURVEY.md)."""
    N = st.P.shape[0]
    d1, d2, e = float(delta1), float(delta2), float(eps)
    P = st.P
    O = sh(st.objFc[ys : ys + N, xs : xs + N])                      # :358-362
    Phi = O * P                                                     # :364
    psi = np.fft.ifft2(Phi)                                         # :365
    a = np.sqrt(I.astype(np.float64))                               # :378-387
    psi2 = a * psi / np.abs(psi + e * (1 + 1j * kappa))             # :390-393
    Phi2 = np.fft.fft2(psi2)                                        # :394
    dPhi = Phi2 - Phi                                               # :409,463
    absP = np.abs(P)
    D_O = absP.max() * ((absP ** 2 + d2) + 1j * kappa * d2)         # :415-418
    dO = dPhi * absP * np.conj(P) / D_O                             # :406-410,419
    st.objFc[ys : ys + N, xs : xs + N] += sh(dO)                    # :427-447
    absO = np.abs(O)
    D_P = np.abs(st.objFc).max() * ((absO ** 2 + d1) + 1j * kappa * d1)   # :459-470
    st.P = P + dPhi * absO * np.conj(O) / D_P * st.S                # :461-475


def run(stack, cropX, cropY, L, naRadius, delta1, delta2, eps, iters, kappa=1,
        init_slot=1, trace=None) -> State:
    """`cropX/cropY/stack` are in update order (sortedIndicies)."""
    st = init_state(stack, L, naRadius, init_slot)
    for it in range(iters):
        for k in range(stack.shape[0]):
            update(st, stack[k], int(cropX[k]), int(cropY[k]), delta1, delta2, eps, kappa)
            if trace is not None:
                trace(it, k, st)
    return st


def obj_crop(st: State) -> np.ndarray:
    """fpmMain.cpp:481: dft(objF, DFT_INVERSE|DFT_SCALE) on the DC-at-corner spectrum."""
    return np.fft.ifft2(np.fft.ifftshift(st.objFc))


def rel_l2(a, b) -> float:
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def mosaic(obj_crops, nx, ny, step, Np):
    """Feathered amplitude mosaic of a regular nx x ny grid of tiles (test oracle of fpmb200_mosaic; the reference
    has no tile loop, fpmMain.cpp:519,532-533).  obj_crops: [nx*ny][L][L] complex; tile (ix,iy) sits at hi-res
    offset (ix*step*f, iy*step*f), f = L/Np; overlaps are cross-faded with separable ramps
    w(u) = min(d+1, ov+1)/(ov+1), d = distance to the nearer tile edge, ov = hi-res overlap."""
    L = obj_crops[0].shape[0]
    f = L // Np
    sh = step * f
    ov = L - sh
    u = np.arange(L)
    d = np.minimum(u, L - 1 - u)
    w1 = np.ones(L) if ov <= 0 else np.where(d >= ov, 1.0, (d + 1.0) / (ov + 1.0))
    w2 = np.outer(w1, w1)
    Hm, Wm = (ny - 1) * sh + L, (nx - 1) * sh + L
    acc = np.zeros((Hm, Wm))
    ws = np.zeros((Hm, Wm))
    for iy in range(ny):
        for ix in range(nx):
            a = np.abs(np.asarray(obj_crops[iy * nx + ix], dtype=np.complex128))
            acc[iy * sh:iy * sh + L, ix * sh:ix * sh + L] += w2 * a
            ws[iy * sh:iy * sh + L, ix * sh:ix * sh + L] += w2
    return acc / ws
