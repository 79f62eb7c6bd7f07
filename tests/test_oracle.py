"""CPU tests of the checker itself: the oracle against every pinned vector we have
(golden geometry from the reference's own jsoncpp + std::sort,
cv2 op-sequence mirror fixtures)."""
import glob
import json
import os
import struct

import numpy as np
import pytest

import fpm_oracle as orc
import fpm_testlib as T


def bits(f):
    return struct.unpack("I", struct.pack("f", float(f)))[0]


@pytest.mark.parametrize("name", T.ALL_CFGS + T.QUIRKS)
def test_oracle_geometry_matches_reference_jsoncpp_driver(name):
    g = T.golden(name)
    j = orc.load_json_lenient(T.embedded_path(name))
    cfg = orc.config_from_json(j)
    assert (cfg.Np, cfg.factor, cfg.Nlarge, cfg.naRadius) == (g["Np"], g["factor"], g["Nlarge"], g["naRadius"])
    assert (bits(cfg.ps_eff), bits(cfg.du)) == (g["ps_eff_bits"], g["du_bits"])
    assert float(cfg.delta1) == g["delta1"] and float(cfg.delta2) == g["delta2"]
    assert cfg.arrayRotation == g["arrayRotation"]
    geo = orc.led_geometry(cfg, j, range(g["present_first"], g["present_last"] + 1))
    got = [(n, bits(geo.na[n]), geo.idx_u[n], geo.idx_v[n], geo.cropX[n], geo.cropY[n]) for n in geo.led_nums]
    want = [(l["n"], l["na_bits"], l["idx_u"], l["idx_v"], l["cropX"], l["cropY"]) for l in g["leds"]]
    assert got == want                      # bit-exact NA, indices and crop boxes
    assert geo.order == g["order"]          # libstdc++ introsort tie order reproduced


def test_known_answers_from_reference_profile():
    """K1/K2 (output.svg:880,866,498): dogStomach as shipped -> 157 LEDs pass, Np=200 -> Nlarge=600."""
    g = T.golden("cfg4s_dogStomach_np200")
    assert g["ledUsedCount"] == 157 and g["Np"] == 200 and g["Nlarge"] == 600 and g["factor"] == 3
    assert g["parse_ok"] is True            # our derived JSON has no trailing comma ...
    assert T.golden("quirks_rot")["parse_ok"] is False   # ... the quirks fixture does (jsoncpp recovery)


def test_crop_boxes_stay_inside_spectrum():
    for name in T.ALL_CFGS:
        g = T.golden(name)
        for l in g["leds"]:
            assert 0 <= l["cropX"] <= g["Nlarge"] - g["Np"] and 0 <= l["cropY"] <= g["Nlarge"] - g["Np"]


def test_dome_table_matches_holePositions_permuted():
    """K3: include/domeHoleCoordinates.h (c0,c1,c2) == dataset_cellscope2.json holePositions (z,y,x)."""
    a = orc.load_json_lenient(os.path.join(T.GOLD, "cfg6_mono_dome_np64.embedded.json"))["holeCoordinates"]
    b = orc.load_json_lenient(os.path.join(T.GOLD, "cfg5_cellscope2_np128.embedded.json"))["holeCoordinates"]
    assert len(a) == len(b) == 508
    A = np.array([[r[0]["x"], r[1]["y"], r[2]["z"]] for r in a])
    B = np.array([[r[0]["x"], r[1]["y"], r[2]["z"]] for r in b])
    assert np.abs(A - B).max() < 6e-5       # header is the same table rounded to 4 decimals


def test_introsort_restatement_on_tie_heavy_vectors():
    rng = np.random.default_rng(3)
    for n in (1, 2, 15, 16, 17, 33, 200, 509, 582):
        v = np.round(rng.uniform(0, 1, n), 1).astype(np.float32)     # many exact ties
        idx = orc.libstdcxx_sort_indexes(v)
        assert sorted(idx) == list(range(n))
        assert all(v[idx[k]] <= v[idx[k + 1]] for k in range(n - 1))


def test_pupil_support_matches_cv2_circle():
    cv2 = pytest.importorskip("cv2")
    for N, r in [(64, 21), (64, 24), (128, 17), (128, 42), (256, 94), (90, 26)] + [(64, r) for r in range(1, 31)]:
        m = np.zeros((N, N), np.float64)
        cv2.circle(m, (N // 2, N // 2), r, 1.0, -1, 8, 0)
        assert np.array_equal(orc.sh(m), orc.pupil_support(N, r)), (N, r)


def test_numpy_oracle_matches_cv2_mirror():
    """The windowed restatement against the 1:1 OpenCV op sequence of fpmMain.cpp:345-482."""
    pytest.importorskip("cv2")
    import cv2_mirror
    c = T.Case("cfg1_mono_np64", 11, n_leds=16)
    for kappa in (1, 0):
        a = c.oracle_run(2, kappa)
        m = cv2_mirror.run(c.stack, c.cx, c.cy, c.L, c.r, c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 2, kappa)
        assert orc.rel_l2(a.objFc, m.objFc()) < 1e-13
        assert orc.rel_l2(a.P, m.P()) < 1e-13
        assert orc.rel_l2(orc.obj_crop(a), m.objCropC()) < 1e-13


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(T.GOLD, "loop_*.npz"))))
def test_oracle_matches_committed_cv2_fixtures(path):
    z = np.load(path)
    c = T.Case(str(z["name"]), int(z["seed"]), int(z["n_leds"]))
    st = c.oracle_run(int(z["iters"]), int(z["kappa"]))
    assert orc.rel_l2(T.corner(st.objFc), z["objF"]) < 2e-7      # fixtures are stored as complex64
    assert orc.rel_l2(st.P, z["pupil"]) < 2e-7
    assert orc.rel_l2(orc.obj_crop(st), z["objCrop"]) < 2e-7


def test_synthetic_stack_properties():
    c = T.case("cfg1_mono_np64")
    assert c.stack.dtype == np.uint16 and c.stack.shape == (117, 64, 64) and c.stack.max() == 60000
    c2 = T.Case("cfg1_mono_np64", 1234)
    assert np.array_equal(c.stack, c2.stack)      # seeded, reproducible


# ---- the plain-C restatement (oracle/fpm_oracle.c), used by the full-size GPU parity tests ---------------------
@pytest.mark.parametrize("n", [2, 4, 6, 8, 10, 12, 30, 50, 64, 90, 100, 108, 128, 200, 256, 384, 600])
def test_c_oracle_fft_matches_numpy(n):
    import c_oracle
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert orc.rel_l2(c_oracle.fft1d(x, -1), np.fft.fft(x)) < 1e-14
    assert orc.rel_l2(c_oracle.fft1d(x, +1), np.fft.ifft(x) * n) < 1e-14


@pytest.mark.parametrize("name,n_leds,iters", [("cfg1_mono_np64", 30, 2), ("cfg7_mono_np90", 20, 2), ("cfg2_fLEDc_np128", 20, 1),
                                               ("cfg4s_dogStomach_np200", 12, 1), ("cfg5b_cellscope2_np256", 6, 1)])
@pytest.mark.parametrize("kappa", [1, 0])
def test_c_oracle_matches_numpy_oracle(name, n_leds, iters, kappa):
    """fpm_oracle.c == fpm_oracle.py to 1e-13 (power-of-two and mixed-radix tiles, both scalar readings), including a
    start in the middle of the LED list (slot_begin) as the per-step tests use it."""
    import c_oracle
    c = T.Case(name, 5, n_leds)
    a = c.oracle_run(iters, kappa)
    b = c_oracle.run(c.stack, c.cx, c.cy, c.L, c.r, c.cfg.delta1, c.cfg.delta2, c.cfg.eps, iters, kappa)
    assert orc.rel_l2(b.objFc, a.objFc) < 1e-13 and orc.rel_l2(b.P, a.P) < 1e-13
    # two more updates from slot 3 on both
    for k in (3, 4):
        orc.update(a, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, kappa)
    c_oracle.update_inplace(b, c.stack, c.cx, c.cy, 2, c.cfg.delta1, c.cfg.delta2, c.cfg.eps, kappa, slot_begin=3)
    assert orc.rel_l2(b.objFc, a.objFc) < 1e-13 and orc.rel_l2(b.P, a.P) < 1e-13


# ---- the only reference-held pin of the un-vendored cvComplex arithmetic: the per-LED op multiset ----------------
def test_cvcomplex_op_multiset_matches_reference_profile(monkeypatch):
    """/root/reference/output.svg:831-880 (committed as tests/golden/profile_call_counts.json by
    tests/golden/make_profile_counts.py) records how often runFPM called every cvComplex helper in one iteration over
    the 157 LEDs of dataset_dogStomach.json.  The op-sequence mirror that anchors the oracle must issue exactly that
    multiset on the same geometry: 9 complexMultiply, 4 complexAbs, 3 complexDivide, 2 complexConj, 5 fftShift,
    1 ifft2, 1 fft2 per LED, +1 complexMultiply +3 fftShift +1 fft2 at initialisation (fpmMain.cpp:310-343)."""
    pytest.importorskip("cv2")
    import cv2_mirror
    want = json.load(open(os.path.join(T.GOLD, "profile_call_counts.json")))
    counts = {}

    def counted(name):
        fn = getattr(cv2_mirror, name)

        def wrapper(*a, **k):
            counts[name] = counts.get(name, 0) + 1
            return fn(*a, **k)
        return wrapper

    for name in want["runFPM"]:
        monkeypatch.setattr(cv2_mirror, name, counted(name))
    n_dft = [0]
    real_dft = cv2_mirror.cv2.dft

    def dft(src, *a, **k):
        n_dft[0] += src.shape[0] + src.shape[1]          # 1-D transforms of a 2-D cv::dft: rows + columns
        return real_dft(src, *a, **k)
    monkeypatch.setattr(cv2_mirror.cv2, "dft", dft)
    c = T.Case("cfg4s_dogStomach_np200", 3)              # the shipped dogStomach tile: Np 200, Nlarge 600
    assert len(c.cx) == 157 and (c.N, c.L) == (200, 600)
    cv2_mirror.run(c.stack, c.cx, c.cy, c.L, c.r, c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1, 1)
    assert counts == {k: v["calls"] for k, v in want["runFPM"].items()}, counts
    # K2: 127 200 one-dimensional DFT_64f = 315 x (200 + 200) + (600 + 600)
    assert n_dft[0] == want["other"]["cv::DFT_64f"]["calls"]
    # and per LED (initialisation subtracted)
    per_led = {k: (counts[k] - init) // 157 for k, init in
               (("complexMultiply", 1), ("complexAbs", 0), ("complexDivide", 0), ("complexConj", 0), ("fftShift", 3), ("ifft2", 0), ("fft2", 1))}
    assert per_led == {"complexMultiply": 9, "complexAbs": 4, "complexDivide": 3, "complexConj": 2, "fftShift": 5, "ifft2": 1, "fft2": 1}
