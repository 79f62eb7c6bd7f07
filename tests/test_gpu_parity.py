"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI of include/fpmb200.h,
against the float64 oracle on the same seeded synthetic stacks.

Tolerances (BASELINE.json north_star): relative L2 <= 1e-5 per update step, <= 1e-3 after the full
iteration count, on objF / pupil / objCrop.  kappa = 1 (literal OpenCV scalar broadcast, SURVEY 8c
R5) unless stated.  Integer inputs (crop tables, LED order) come from the golden geometry, which the
CPU tests pin bit-exactly."""
import glob
import os
import subprocess

import numpy as np
import pytest

import fpm_oracle as orc
import fpm_testlib as T

pytestmark = pytest.mark.gpu

STEP_TOL = 1e-5
FULL_TOL = 1e-3


def _need_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.fixture(autouse=True)
def _gpu():
    _need_gpu()


def note(msg):
    """Parity numbers worth keeping: printed, and appended to gpurun_out/parity_notes.txt when that directory exists
    (the GPU runner brings it back; summaries are committed under profiles/)."""
    print(msg)
    d = os.path.join(T.ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_notes.txt"), "a") as f:
            f.write(msg + "\n")


def compare(ctx, st, tile=0, tol=FULL_TOL, crop=True):
    gF, gC, gP = ctx.download(tile, objCrop=crop)
    eF, eP = orc.rel_l2(gF, T.corner(st.objFc)), orc.rel_l2(gP, st.P)
    assert np.isfinite(gF).all() and np.isfinite(gP).all()
    assert eF < tol and eP < tol, (eF, eP)
    if crop:
        eC = orc.rel_l2(gC, orc.obj_crop(st))
        assert eC < tol, eC
        # amplitude and phase separately (phase where amplitude is significant)
        oc = orc.obj_crop(st)
        amp_err = orc.rel_l2(np.abs(gC), np.abs(oc))
        m = np.abs(oc) > 1e-3 * np.abs(oc).max()
        ph_err = np.abs(np.angle(gC[m] * np.conj(oc[m]))).max()
        assert amp_err < tol and ph_err < tol, (amp_err, ph_err)
    return eF, eP


# ---- initialisation (fpmMain.cpp:301-343) ----------------------------------------------------
@pytest.mark.parametrize("name", ["cfg1_mono_np64", "cfg2_fLEDc_np128", "cfg5_cellscope2_np128"])
def test_init_state(name):
    c = T.case(name)
    ctx = c.make_ctx()
    st = orc.init_state(c.stack, c.L, c.r)
    gF, _, gP = ctx.download(0, objCrop=False)
    assert np.array_equal(gP, st.P.astype(np.complex64))          # pupil = support, exactly
    assert orc.rel_l2(gF, T.corner(st.objFc)) < 1e-6
    ctx.close()


# ---- per-step parity: every update starts from the oracle's float64 state ---------------------
@pytest.mark.parametrize("name,n_steps", [("cfg1_mono_np64", 117), ("cfg2_fLEDc_np128", 89), ("cfg3b_cellScope_np64", 60),
                                          ("cfg4_dogStomach_np128", 157), ("cfg5_cellscope2_np128", 40),
                                          ("cfg5b_cellscope2_np256", 16), ("cfg3_cellScope_np256", 8)])
@pytest.mark.parametrize("kappa", [1, 0])
@pytest.mark.parametrize("ctas", [1, 0])
def test_per_step_parity(name, n_steps, kappa, ctas):
    """ctas = CTAs per tile: 1 = fpm_update_kernel, 0 = the library's choice (a 4-CTA cluster for one 128x128 tile, a
    cluster of 8 for a 256x256 tile).  cfg4 is the bench configuration (every one of its 157 LEDs); the Np=256 cases
    start from the float64 state after a whole pass of the C oracle."""
    import c_oracle
    c = T.case(name)
    if ctas == 0 and c.N == 64:
        pytest.skip("same kernel as ctas=1")
    if c.N == 256 and kappa == 0 and ctas == 1:
        pytest.skip("kappa is a run-time scalar of the same kernel; covered at ctas=0")
    ctx = c.make_ctx(kappa=kappa, cluster=ctas)
    assert ("cluster_kernel" in ctx.variant) == (ctas == 0)
    st = orc.init_state(c.stack, c.L, c.r)
    # warm the oracle state up so that the pupil is not the trivial binary mask
    c_oracle.update_inplace(st, c.stack, c.cx, c.cy, len(c.cx), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, kappa)
    worst = 0.0
    for k in range(min(n_steps, len(c.cx))):
        ctx.upload_state(0, T.corner(st.objFc), st.P)
        orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, kappa)
        ctx.step(0, k)
        worst = max(worst, *compare(ctx, st, tol=STEP_TOL, crop=False))
    note("per-step %s kappa=%d ctas=%d: worst rel-L2 over %d updates %.2e [%s]" % (name, kappa, ctas, min(n_steps, len(c.cx)), worst, ctx.variant[:60]))
    ctx.close()


# ---- full runs at the BASELINE configurations --------------------------------------------------
@pytest.mark.parametrize("name,iters", [("cfg1_mono_np64", 10), ("cfg2_fLEDc_np128", 10), ("cfg3b_cellScope_np64", 3),
                                        ("cfg4_dogStomach_np128", 10), ("cfg5_cellscope2_np128", 4),
                                        ("cfg6_mono_dome_np64", 3)])
@pytest.mark.parametrize("ctas", [1, 0])
def test_full_run_parity(name, iters, ctas):
    c = T.case(name)
    if ctas == 0 and c.N != 128:
        pytest.skip("same kernel as ctas=1")
    ctx = c.make_ctx(cluster=ctas)
    assert ("cluster_kernel" in ctx.variant) == (ctas == 0)
    ctx.run(iters)
    ctx.finalize()
    st = c.oracle_run(iters)
    e = compare(ctx, st)
    note("full run %s %d iterations x %d LEDs: rel-L2 objF %.2e pupil %.2e [%s]" % (name, iters, len(c.cx), e[0], e[1], ctx.variant))
    ctx.close()


@pytest.mark.parametrize("name,n_steps,iters", [("cfg1_mono_np64", 30, 3), ("cfg4_dogStomach_np128", 30, 2)])
def test_older_update_kernel_parity(name, n_steps, iters, monkeypatch):
    """The three-phase kernel (fpm_update_phased_kernel) is the library's choice for 64 x 64 tiles and for narrow pupils
    on 128 x 128 tiles; FPMB200_UPDATE_V1=1 keeps fpm_update_kernel, which still serves every other geometry (coarser
    max-cells, Nlarge not a multiple of 64).  Both are held to the same per-step and full-run tolerances, and agree
    with each other to rounding (the last column stage is a direct sum in one and a butterfly in the other)."""
    c = T.case(name)
    new = c.make_ctx(cluster=1)
    assert "fpm_update_phased_kernel" in new.variant, new.variant
    monkeypatch.setenv("FPMB200_UPDATE_V1", "1")
    old = c.make_ctx(cluster=1)
    assert "fpm_update_kernel<" in old.variant, old.variant
    st = orc.init_state(c.stack, c.L, c.r)
    worst = 0.0
    for k in range(n_steps):
        for ctx in (old, new):
            ctx.upload_state(0, T.corner(st.objFc), st.P)
        orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
        for ctx in (old, new):
            ctx.step(0, k)
            worst = max(worst, *compare(ctx, st, tol=STEP_TOL, crop=False))
    for ctx in (old, new):
        ctx.init_tiles()
        ctx.run(iters)
        ctx.finalize()
    ref = c.oracle_run(iters)
    e_old, e_new = compare(old, ref), compare(new, ref)
    for x, y in zip(old.download(0), new.download(0)):
        assert orc.rel_l2(x, y) < 1e-5
    note("older vs three-phase kernel %s: per-step worst %.2e; %d iterations: objF %.2e / %.2e, pupil %.2e / %.2e" % (
        name, worst, iters, e_old[0], e_new[0], e_old[1], e_new[1]))
    old.close()
    new.close()


def test_full_run_parity_kappa0():
    c = T.case("cfg1_mono_np64")
    ctx = c.make_ctx(kappa=0)
    ctx.run(5)
    ctx.finalize()
    compare(ctx, c.oracle_run(5, kappa=0))
    ctx.close()


@pytest.mark.parametrize("name,ctas", [("cfg1_mono_np64", 1), ("cfg2_fLEDc_np128", 1), ("cfg2_fLEDc_np128", 0), ("cfg7_mono_np90", 1)])
def test_dark_and_saturated_pixels(name, ctas):
    """Zero and full-scale intensities: sqrt(0) = 0 wipes the pixel (the device keeps 1/I = inf and must not produce
    NaN), 65535 is the largest uint16 the loop reinterprets (fpmMain.cpp:380).  Whole dark images, dark rows, a
    checkerboard of zeros and saturated blocks; every kernel family (fused, cluster, general path)."""
    import copy
    c = copy.copy(T.Case(name, 77, n_leds=24))
    st_ = c.stack.copy()
    st_[3] = 0                                   # a completely dark frame
    st_[5, ::2, :] = 0                           # dark rows
    st_[7, ::2, ::2] = 0; st_[7, 1::2, 1::2] = 0  # checkerboard
    st_[9, :8, :8] = 65535                       # saturated block
    st_[11] = 65535                              # a saturated frame
    st_[13][st_[13] < np.median(st_[13])] = 0    # thresholded frame (typical after background subtraction)
    c.stack = st_
    ctx = c.make_ctx(cluster=ctas)
    ctx.run(3)
    ctx.finalize()
    st = c.oracle_run(3)
    e = compare(ctx, st)
    print("%s dark/saturated: rel-L2 objF %.2e pupil %.2e [%s]" % (name, e[0], e[1], ctx.variant))
    ctx.close()


@pytest.mark.parametrize("name,n_leds,iters", [("cfg5b_cellscope2_np256", 40, 2), ("cfg3_cellScope_np256", 24, 1)])
@pytest.mark.parametrize("ctas", [8, 1])
def test_np256_tiles(name, n_leds, iters, ctas):
    """Np=256, Nlarge 1024 / 1536 = 3*2^9.  The field does not fit one SM's shared memory: 8 CTAs of a cluster share
    the tile (default); with one CTA per tile the field lives in a global scratch buffer."""
    c = T.case(name, n_leds=n_leds)
    ctx = c.make_ctx(cluster=ctas)
    assert ("cluster=8" in ctx.variant) == (ctas == 8)
    ctx.run(iters)
    ctx.finalize()
    e = compare(ctx, c.oracle_run(iters))
    print(name, e, ctx.variant)
    ctx.close()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(T.GOLD, "loop_*.npz"))))
def test_against_committed_opencv_fixtures(path):
    """States produced by OpenCV's own cv::dft/arithmetic op sequence (oracle/cv2_mirror.py)."""
    z = np.load(path)
    c = T.Case(str(z["name"]), int(z["seed"]), int(z["n_leds"]))
    ctx = c.make_ctx(kappa=int(z["kappa"]))
    ctx.run(int(z["iters"]))
    ctx.finalize()
    gF, gC, gP = ctx.download(0)
    assert orc.rel_l2(gF, z["objF"]) < 1e-5 and orc.rel_l2(gP, z["pupil"]) < 1e-5 and orc.rel_l2(gC, z["objCrop"]) < 1e-5
    ctx.close()


# ---- the shipped JSONs with their literal cropSizeX (90, 100: fused general kernel; 200: unfused path) ---------
@pytest.mark.parametrize("name,n_leds,iters", [("cfg7_mono_np90", 117, 3), ("cfg8_cellScope_np100", 60, 2),
                                               ("cfg4s_dogStomach_np200", 40, 2)])
@pytest.mark.parametrize("kappa", [1, 0])
def test_literal_tile_sizes(name, n_leds, iters, kappa):
    """Np = 90 / 100 / 200 (Nlarge 360 / 600 / 600): mixed-radix transforms, cells of the max grid cut by the spectrum
    border.  Per-step parity from the oracle's state for the first LEDs, then a full run."""
    c = T.case(name, n_leds=n_leds)
    ctx = c.make_ctx(kappa=kappa)
    assert "general path" in ctx.variant
    st = orc.init_state(c.stack, c.L, c.r)
    gF, _, gP = ctx.download(0, objCrop=False)
    assert orc.rel_l2(gF, T.corner(st.objFc)) < 1e-6 and np.array_equal(gP, st.P.astype(np.complex64))
    for k in range(len(c.cx)):
        orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, kappa)
    worst = 0.0
    for k in range(min(12, len(c.cx))):
        ctx.upload_state(0, T.corner(st.objFc), st.P)
        orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, kappa)
        ctx.step(0, k)
        worst = max(worst, *compare(ctx, st, tol=STEP_TOL, crop=False))
    ctx.init_tiles()
    ctx.run(iters)
    ctx.finalize()
    e = compare(ctx, c.oracle_run(iters, kappa=kappa))
    print("%s kappa=%d per-step %.2e, %d iterations: objF %.2e pupil %.2e" % (name, kappa, worst, iters, e[0], e[1]))
    ctx.close()


@pytest.mark.parametrize("N,factor,expect", [(60, 4, "radix 10 x 6"), (72, 3, "radix 9 x 8"), (80, 3, "radix 10 x 8"),
                                             (96, 3, "radix 16 x 6"), (50, 4, "run-time radices"), (108, 3, "run-time radices")])
def test_general_fused_other_sizes(N, factor, expect):
    """Every compiled two-stage radix plan of fpm_update_general_kernel besides the shipped 90 / 100, and the run-time
    radix variant (Np = 50 = 2*5*5, 108 = 4*3*3*3): per-step and full-run parity on synthetic geometry."""
    c = T.SyntheticCase(N, factor, 100 + N, 24)
    ctx = c.make_ctx()
    assert "general path, fused" in ctx.variant and expect in ctx.variant, ctx.variant
    st = orc.init_state(c.stack, c.L, c.r)
    for k in range(len(c.cx)):
        orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
    for k in range(6):
        ctx.upload_state(0, T.corner(st.objFc), st.P)
        orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
        ctx.step(0, k)
        compare(ctx, st, tol=STEP_TOL, crop=False)
    ctx.init_tiles()
    ctx.run(3)
    ctx.finalize()
    compare(ctx, c.oracle_run(3))
    ctx.close()


@pytest.mark.parametrize("N,factor,r,expect", [(160, 3, 20, "radix 16 x 10"), (240, 2, 22, "radix 20 x 12"), (300, 2, 18, "radix 20 x 15"),
                                               (200, 3, 40, "radix 20 x 10"), (128, 3, 17, "radix 16 x 8")])
def test_pruned_fused_other_sizes(N, factor, r, expect, monkeypatch):
    """Every compiled plan of fpm_update_pruned_kernel (box rows + column batches in shared memory) besides the shipped
    Np = 200 case: per-step and full-run parity on synthetic geometry; Np = 128 reaches it through the developer switch
    that sends power-of-two tiles down the general path."""
    if N == 128:
        monkeypatch.setenv("FPMB200_FORCE_GENERAL", "1")
    c = T.SyntheticCase(N, factor, 100 + N, 16, r=r)
    ctx = c.make_ctx()
    assert "fpm_update_pruned_kernel" in ctx.variant and expect in ctx.variant, ctx.variant
    st = orc.init_state(c.stack, c.L, c.r)
    gF, _, gP = ctx.download(0, objCrop=False)
    assert orc.rel_l2(gF, T.corner(st.objFc)) < 1e-6
    for k in range(len(c.cx)):
        orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
    for k in range(5):
        ctx.upload_state(0, T.corner(st.objFc), st.P)
        orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
        ctx.step(0, k)
        compare(ctx, st, tol=STEP_TOL, crop=False)
    ctx.init_tiles()
    ctx.run(2)
    ctx.finalize()
    e = compare(ctx, c.oracle_run(2))
    note("pruned fused Np=%d: 2 iterations x 16 LEDs rel-L2 objF %.2e pupil %.2e [%s]" % (N, e[0], e[1], ctx.variant[:110]))
    ctx.close()


def test_general_fused_variants_agree(monkeypatch):
    """Np = 90: the two-stage plan, the run-time radices and the unfused per-step kernels are three implementations of
    the same update; each within the full-run tolerance of the oracle (checked against one oracle run)."""
    c = T.Case("cfg7_mono_np90", 5, 40)
    ref = c.oracle_run(2)
    for env, expect in (({}, "radix 10 x 9"), ({"FPMB200_GENERAL_PLAN": "0"}, "run-time radices"),
                        ({"FPMB200_GENERAL_UNFUSED": "1"}, "unfused")):
        for k in ("FPMB200_GENERAL_PLAN", "FPMB200_GENERAL_UNFUSED"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = c.make_ctx()
        assert expect in ctx.variant, ctx.variant
        ctx.run(2)
        ctx.finalize()
        compare(ctx, ref)
        ctx.close()


def test_general_path_many_tiles():
    """Batched over tiles: identical inputs -> identical bits in every slot; sub-ranges of tiles are independent."""
    c = T.Case("cfg7_mono_np90", 11, 20)
    a = c.make_ctx(n_tiles=5)
    a.run(2, 0, 5)
    b = c.make_ctx(n_tiles=5)
    b.run(2, 0, 2)
    b.run(2, 2, 3)
    ref = a.download(0, objCrop=False)
    for t in range(5):
        for ctx in (a, b):
            for x, y in zip(ctx.download(t, objCrop=False), ref):
                if x is not None:
                    assert np.array_equal(x, y)
    a.close(), b.close()


# ---- full field of view: device-side frame ingest and mosaic ---------------------
@pytest.mark.parametrize("name", ["cfg1_mono_np64", "cfg7_mono_np90"])
def test_frame_ingest_matches_host_loader(name):
    """fpmb200_ingest_frame == loadFPMDataset's per-tile preprocessing (fpmMain.cpp:124-144), bit for bit, for every
    tile and LED: ROI cut, cv::divide for dark-field LEDs, background estimate / clamp / saturating subtract; and the
    1/I stack it writes drives the update kernels to the same bits as fpmb200_upload_stack of the same images."""
    import fpmhost
    c = T.Case(name, 21, 10)
    N, n_leds = c.N, 10
    rng = np.random.default_rng(5)
    W, H, ov = 3 * N + 37, 2 * N + 21, N // 4
    nx, ny, xs, ys = fpmhost.tile_grid(W, H, N, ov)
    n_tiles = nx * ny
    assert n_tiles >= 6
    frames = (rng.integers(0, 3000, (n_leds, H, W)) + (rng.random((n_leds, H, W)) < 0.01) * 60000).astype(np.uint16)
    frames[3] += 900                                              # background above the clamp
    bk1, bk2, thr = (W - N - 3, 2), (5, H - N - 1), 2300
    divisors = [1, 1, 3, 1, 0, 7, 1, 2, 1, 5]
    a = c.make_ctx(n_tiles=n_tiles)
    a.set_tile_origins(xs, ys)
    for k in range(n_leds):
        a.ingest_frame(k, frames[k], divisors[k], bk1, bk2, thr)
    bg = a.ingest_bg()
    ref = np.zeros((n_tiles, n_leds, N, N), np.uint16)
    for t in range(n_tiles):
        for k in range(n_leds):
            ref[t, k], b = fpmhost.preprocess_frame(frames[k], N, (xs[t], ys[t]), bk1, bk2, divisors[k], thr)
            assert b == bg[k]
    assert bg[3] == thr and len(set(bg.tolist())) > 2
    for t in (0, n_tiles // 2, n_tiles - 1):
        assert np.array_equal(a.raw_stack(t), ref[t])
    # a context that owns only tiles 2..4 of the grid and receives only the frame rows they need, with the background
    # value computed by the host (fpmb200_ingest_rows: the multi-GPU path of run_fov.cpp); markers order the copies
    sub = list(range(2, 5))
    r_ctx = c.make_ctx(n_tiles=len(sub))
    r_ctx.set_tile_origins([xs[t] for t in sub], [ys[t] for t in sub])
    row0 = min(ys[t] for t in sub)
    n_rows = max(ys[t] for t in sub) + N - row0
    for k in range(n_leds):
        r_ctx.ingest_rows(k, frames[k], row0, n_rows, divisors[k], bg[k])
        r_ctx.event_record(k % 4)
    for k in range(4):
        r_ctx.event_sync(k)
    r_ctx.sync()
    assert np.array_equal(r_ctx.ingest_bg(), bg)
    for i, t in enumerate(sub):
        assert np.array_equal(r_ctx.raw_stack(i), ref[t])
    with pytest.raises(Exception):
        r_ctx.ingest_rows(0, frames[0], row0 + 1, n_rows - 1, 1, 0)        # rows that do not cover the tiles are refused
    r_ctx.close()
    b_ctx = c.make_ctx(n_tiles=n_tiles)
    for t in range(n_tiles):
        b_ctx.upload_stack(t, ref[t])
    for ctx in (a, b_ctx):
        ctx.init_tiles()
        ctx.run(2)
        ctx.finalize()
    for t in range(n_tiles):
        for x, y in zip(a.download(t), b_ctx.download(t)):
            assert np.array_equal(x, y)
    a.close(), b_ctx.close()


@pytest.mark.parametrize("overlap", [0, 16, 24])
def test_mosaic(overlap):
    c = T.Case("cfg1_mono_np64", 31, 12)
    nx, ny, N = 3, 2, c.N
    ctx = c.make_ctx(n_tiles=nx * ny)
    rng = np.random.default_rng(overlap)
    for t in range(nx * ny):                                    # different content per tile
        ctx.upload_stack(t, (c.stack * rng.uniform(0.5, 1.0)).astype(np.uint16))
    ctx.init_tiles()
    ctx.run(1)
    ctx.finalize()
    crops = [ctx.download(t)[1] for t in range(nx * ny)]
    got = ctx.mosaic(nx, ny, N - overlap)
    ref = orc.mosaic(crops, nx, ny, N - overlap, N)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    ctx.close()


# ---- structure / edge cases ----------------------------------------------------------------------
def test_dense_support_uses_general_path():
    """A support mask covering the whole window: no bbox pruning, pupil kept in global memory at
    Np=128 (does not fit next to the field).  Same arithmetic as the oracle with S = 1."""
    for name, n_leds in (("cfg1_mono_np64", 20), ("cfg2_fLEDc_np128", 12)):
        c = T.Case(name, 3, n_leds)
        full = np.ones((c.N, c.N), np.float32)
        ctx = c.make_ctx(support=full)
        assert "bbox=[-%d..%d]" % (c.N // 2, c.N // 2 - 1) in ctx.variant
        st = orc.init_state(c.stack, c.L, c.r)
        st.S = full.astype(np.float64)
        ctx.upload_state(0, T.corner(st.objFc), st.P)
        for it in range(2):
            for k in range(n_leds):
                orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
        ctx.run(2)
        ctx.finalize()
        compare(ctx, st)
        ctx.close()


@pytest.mark.parametrize("name,ell,unfused", [("cfg1_mono_np64", (3, 14.0, -5, 9.0), 0), ("cfg7_mono_np90", (3, 14.0, -5, 9.0), 0),
                                              ("cfg7_mono_np90", (20, 9.0, 12, 6.0), 0), ("cfg8_cellScope_np100", (-24, 11.0, 0, 30.0), 0),
                                              ("cfg7_mono_np90", (3, 14.0, -5, 9.0), 1), ("cfg7_mono_np90", (20, 9.0, 12, 6.0), 1),
                                              ("cfg4s_dogStomach_np200", (-60, 25.0, 33, 40.0), 0), ("cfg4s_dogStomach_np200", (-60, 25.0, 33, 40.0), 1),
                                              ("cfg4s_dogStomach_np200", (5, 30.0, -9, 12.0), 0)])
def test_asymmetric_support_bbox(name, ell, unfused, monkeypatch):
    """An off-centre elliptical support exercises the wrapped bbox arithmetic (fused power-of-two kernel, the fused
    general kernel and the unfused general path; boxes that straddle the origin or lie on one side of it)."""
    if unfused:
        monkeypatch.setenv("FPMB200_GENERAL_UNFUSED", "1")
    c = T.Case(name, 4, 16)
    N = c.N
    y, x = np.mgrid[0:N, 0:N]
    yw, xw = np.where(y < N // 2, y, y - N), np.where(x < N // 2, x, x - N)
    x0, ax, y0, ay = ell
    S = ((((xw - x0) / ax) ** 2 + ((yw - y0) / ay) ** 2) <= 1).astype(np.float32)
    ctx = c.make_ctx(support=S)
    st = orc.State(orc.init_state(c.stack, c.L, c.r).objFc, S.astype(np.complex128), S.astype(np.float64))
    ctx.upload_state(0, T.corner(st.objFc), st.P)
    for k in range(16):
        orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
    ctx.run(1)
    ctx.finalize()
    compare(ctx, st)
    ctx.close()


@pytest.mark.parametrize("name,ell", [("cfg2_fLEDc_np128", (3, 14.0, -5, 9.0)),      # 29 x 19 box: no full group of 32 columns
                                      ("cfg2_fLEDc_np128", (-2, 5.0, 1, 5.0)),       # 11 x 11: a rectangle inside one or two max-cells
                                      ("cfg2_fLEDc_np128", (8, 9.0, -6, 11.0)),      # 19 x 23, off-centre: one-sided in x
                                      ("cfg1_mono_np64", (0, 31.0, 0, 31.0)),        # 63 x 63: every thread has a phase-A butterfly
                                      ("cfg1_mono_np64", (-20, 10.0, 0, 31.0)),      # off-centre, tall
                                      ("cfg1_mono_np64", (0, 4.0, 0, 4.0))])
def test_phased_kernel_box_shapes(name, ell):
    """fpm_update_phased_kernel over the shapes its work decomposition distinguishes (column groups of 32 + packed
    left-overs in the column B stages and in phase C, dedicated / shared threads for the max-cell rebuild, one to five
    touched cell columns), with windows at the spectrum border among the LEDs; 2 passes against the oracle."""
    c = T.Case(name, 6, 14)
    N = c.N
    c.cx[10:14] = np.array([0, c.L - N, 3, c.L - N - 1], np.int16)
    c.cy[10:14] = np.array([0, c.L - N, c.L - N - 2, 1], np.int16)
    y, x = np.mgrid[0:N, 0:N]
    yw, xw = np.where(y < N // 2, y, y - N), np.where(x < N // 2, x, x - N)
    x0, ax, y0, ay = ell
    S = ((((xw - x0) / ax) ** 2 + ((yw - y0) / ay) ** 2) <= 1).astype(np.float32)
    ctx = c.make_ctx(support=S, cluster=1)
    assert "fpm_update_phased_kernel" in ctx.variant, ctx.variant
    st = orc.State(orc.init_state(c.stack, c.L, c.r).objFc, S.astype(np.complex128), S.astype(np.float64))
    ctx.upload_state(0, T.corner(st.objFc), st.P)
    for _ in range(2):
        for k in range(len(c.cx)):
            orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
    ctx.run(2)
    ctx.finalize()
    e = compare(ctx, st)
    note("three-phase kernel %s support %s: rel-L2 objF %.2e pupil %.2e [%s]" % (name, ell, e[0], e[1], ctx.variant[:100]))
    ctx.close()


@pytest.mark.timeout(180, method="thread")
@pytest.mark.parametrize("ctas", [4, 2])
@pytest.mark.parametrize("ell", [(3, 5.0, -2, 5.0),       # 11 x 11 off-centre: narrow instance, three columns per CTA (four CTAs: 3, 3, 3, 2)
                                 (0, 1.0, 0, 1.0),        # 3 x 3: with four CTAs the last one owns no column at all
                                 (8, 9.0, -6, 11.0),      # 19 x 23 off-centre, one-sided in x: narrow instance
                                 (0, 30.0, 0, 30.0),      # 61 x 61: wide instance (window in L2), helper threads
                                 (0, 60.0, 0, 60.0)])     # 121 x 121: every thread has a column item, no helper threads
def test_cluster_kernel_box_shapes(ell, ctas):
    """fpm_update_cluster_kernel over the shapes its work decomposition distinguishes: narrow boxes (window slice on chip,
    forwarded between the CTAs; six-sample butterflies) and wide ones, CTAs without columns, with and without helper
    threads; windows at the spectrum border and far jumps among the LEDs (nothing to forward); 2 passes in one
    persistent launch against the oracle."""
    c = T.Case("cfg2_fLEDc_np128", 6, 14)
    N = c.N
    c.cx[10:14] = np.array([0, c.L - N, 3, c.L - N - 1], np.int16)
    c.cy[10:14] = np.array([0, c.L - N, c.L - N - 2, 1], np.int16)
    y, x = np.mgrid[0:N, 0:N]
    yw, xw = np.where(y < N // 2, y, y - N), np.where(x < N // 2, x, x - N)
    x0, ax, y0, ay = ell
    S = ((((xw - x0) / ax) ** 2 + ((yw - y0) / ay) ** 2) <= 1).astype(np.float32)
    try:
        ctx = c.make_ctx(support=S, cluster=ctas)
    except RuntimeError as e:
        pytest.skip("no %d-CTA cluster for this box: %s" % (ctas, e))
    assert "cluster=%d" % ctas in ctx.variant, ctx.variant
    st = orc.State(orc.init_state(c.stack, c.L, c.r).objFc, S.astype(np.complex128), S.astype(np.float64))
    ctx.upload_state(0, T.corner(st.objFc), st.P)
    for _ in range(2):
        for k in range(len(c.cx)):
            orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
    ctx.run(2)
    ctx.finalize()
    e = compare(ctx, st)
    note("cluster kernel support %s: rel-L2 objF %.2e pupil %.2e [%s]" % (ell, e[0], e[1], ctx.variant[:110]))
    ctx.close()


@pytest.mark.parametrize("name", ["cfg1_mono_np64", "cfg7_mono_np90"])
def test_windows_touching_the_spectrum_border(name):
    """Crop origins 0 and Nlarge-Np (the legal extremes, fpmMain.cpp:157-165)."""
    c = T.Case(name, 5, 8)
    c.cx = np.array([0, c.L - c.N, 0, c.L - c.N, 96, 7, c.L - c.N - 1, 1], np.int16)
    c.cy = np.array([0, 0, c.L - c.N, c.L - c.N, 96, c.L - c.N - 3, 5, 1], np.int16)
    ctx = c.make_ctx()
    ctx.run(2)
    ctx.finalize()
    compare(ctx, c.oracle_run(2))
    ctx.close()


@pytest.mark.parametrize("name,ctas,n_leds", [("cfg2_fLEDc_np128", 2, 20), ("cfg2_fLEDc_np128", 4, 20), ("cfg5_cellscope2_np128", 4, 12),
                                              ("cfg5b_cellscope2_np256", 8, 10)])
def test_cluster_kernel(name, ctas, n_leds):
    """The cluster kernel on several tiles at once (more clusters than fit: they run in waves), windows at the
    spectrum border, and n single-update launches == one persistent launch bit for bit (the DSMEM max merges are
    order-independent)."""
    c = T.Case(name, 8, n_leds)
    c.cx = c.cx.copy(); c.cy = c.cy.copy()
    c.cx[2], c.cy[2] = 0, 0
    c.cx[3], c.cy[3] = c.L - c.N, c.L - c.N
    c.cx[4], c.cy[4] = 1, c.L - c.N - 1
    c.stack = None
    import synth
    c.stack = synth.synth_stack(c.N, c.L, c.r, c.cx, c.cy, 8)
    n_tiles = 40 if c.N == 128 else 3
    a, b = c.make_ctx(n_tiles=n_tiles, cluster=ctas), c.make_ctx(cluster=ctas)
    assert "cluster=%d" % ctas in a.variant
    a.run(1)
    for k in range(n_leds):
        b.step(0, k)
    a.finalize(); b.finalize()
    ref = b.download(0)
    for t in (0, n_tiles - 1):
        for x, y in zip(a.download(t), ref):
            assert np.array_equal(x, y)
    compare(a, c.oracle_run(1), tile=n_tiles - 1)
    a.close(), b.close()


@pytest.mark.timeout(120, method="thread")
@pytest.mark.parametrize("name,ctas", [("cfg2_fLEDc_np128", 2), ("cfg2_fLEDc_np128", 4), ("cfg5_cellscope2_np128", 4), ("cfg5b_cellscope2_np256", 8)])
def test_cluster_kernel_repeated_runs(name, ctas):
    """Persistent launches of the cluster kernel over whole iterations, repeated: the window slices forwarded between the
    CTAs (st.async + the receiver's mbarrier) include LED steps after which a CTA's next slice lies outside the current
    rectangle -- nothing to forward, the barrier phase completes when it is armed.  Arming it before every waiter of the
    previous phase had polled made that waiter skip a phase and the launch never finished (timing dependent; found with
    tools/dev_stress_cluster.py).  Every run must finish and give the same bits."""
    c = T.Case(name, 1)
    for n_tiles in (1, 5):
        ctx = c.make_ctx(n_tiles=n_tiles, cluster=ctas)
        assert "cluster=%d" % ctas in ctx.variant
        ref = None
        for rep in range(6):
            ctx.init_tiles(); ctx.run(2); ctx.sync()
            got = ctx.download(n_tiles - 1, objCrop=False)
            if ref is None:
                ref = got
            else:
                assert all(np.array_equal(a, b) for a, b in zip(got, ref) if a is not None)
        ctx.close()


def test_steps_equal_run_bitwise():
    """n_leds single-update launches == one persistent launch, bit for bit (state round-trips exactly)."""
    c = T.Case("cfg1_mono_np64", 6, 30)
    a, b = c.make_ctx(), c.make_ctx()
    a.run(1)
    for k in range(30):
        b.step(0, k)
    for x, y in zip(a.download(0, objCrop=False), b.download(0, objCrop=False)):
        if x is not None:
            assert np.array_equal(x, y)
    a.close(), b.close()


def test_balanced_passes_equal_one_launch(monkeypatch):
    """More tiles than SMs, not a multiple: fpmb200_run re-cuts the run into passes of one iteration over at most
    sm_count tiles (tiles with the most iterations left first) instead of leaving the last wave partly empty.  An
    iteration boundary is a clean cut, so the results equal the single persistent launch bit for bit."""
    c = T.Case("cfg1_mono_np64", 9, 12)
    n_tiles, iters = 150, 3
    out = {}
    for bal in ("0", "1"):
        monkeypatch.setenv("FPMB200_RUN_BALANCED", bal)
        ctx = c.make_ctx(n_tiles=n_tiles)
        l0 = ctx.kernel_launches
        ctx.run(iters)
        out[bal] = (ctx.kernel_launches - l0, [ctx.download(t, objCrop=False) for t in (0, 1, 147, 148, 149)])
        ctx.close()
    assert out["0"][0] == 1 and out["1"][0] > 1, (out["0"][0], out["1"][0])
    for ta, tb in zip(out["0"][1], out["1"][1]):
        for a, b in zip(ta, tb):
            if a is not None:
                assert np.array_equal(a, b)
    note("balanced passes: %d tiles x %d iterations in %d launches, bit-identical to the single launch" % (n_tiles, iters, out["1"][0]))


def test_tiles_are_independent_and_deterministic():
    """Many tiles in one launch (more CTAs than fit at once): identical inputs give identical bits in
    every slot, different inputs do not interfere -- the property the multi-GPU sharding relies on."""
    import fpmb200
    c1, c2 = T.Case("cfg2_fLEDc_np128", 21, 10), T.Case("cfg2_fLEDc_np128", 22, 10)
    n_tiles = 300
    ctx = fpmb200.Context(0)
    ctx.tiles_alloc(n_tiles, c1.N, c1.L, 10)
    ctx.set_params(c1.cfg.delta1, c1.cfg.delta2, c1.cfg.eps, 1)
    ctx.upload_leds(c1.cx, c1.cy)
    ctx.upload_pupil_support(c1.support)
    for t in range(n_tiles):
        ctx.upload_stack(t, c2.stack if t % 7 == 3 else c1.stack)
    ctx.init_tiles()
    ctx.run(2)
    ctx.finalize()
    ref1, ref2 = ctx.download(0), ctx.download(3)
    for t in (1, 2, 147, 148, 149, 295, 299, 10, 17, 290):
        got = ctx.download(t)
        want = ref2 if t % 7 == 3 else ref1
        assert all(np.array_equal(g, w) for g, w in zip(got, want)), t
    compare(ctx, c1.oracle_run(2), tile=299)
    compare(ctx, c2.oracle_run(2), tile=290)
    ctx.close()


def test_argument_errors_are_reported_not_masked():
    import fpmb200
    ctx = fpmb200.Context(0)
    with pytest.raises(fpmb200.FpmError):
        ctx.tiles_alloc(1, 98, 392, 10)             # 98 = 2 * 7^2: prime factor 7 in the tile edge
    with pytest.raises(fpmb200.FpmError):
        ctx.tiles_alloc(1, 75, 300, 10)             # odd tile edge
    with pytest.raises(fpmb200.FpmError):
        ctx.tiles_alloc(1, 64, 448, 10)             # 448 = 2^6 * 7: prime factor 7
    ctx.tiles_alloc(1, 64, 256, 4)
    with pytest.raises(fpmb200.FpmError):
        ctx.upload_leds([0, 0, 0, 193], [0, 0, 0, 0])   # window would leave the spectrum
    with pytest.raises(fpmb200.FpmError):
        ctx.run(1)                                   # nothing uploaded yet
    with pytest.raises(fpmb200.FpmError):
        fpmb200.Context(1 << 20)
    ctx.close()


def _colour_frame(grey, rng):
    """chunky RGB frame whose red plane (sample 0 = channels[2] of OpenCV's BGR, fpmMain.cpp:112-115) carries the data;
    green and blue are noise that must not leak into the reconstruction."""
    rgb = rng.integers(0, 60000, grey.shape + (3,)).astype(np.uint16)
    rgb[..., 0] = grey
    return rgb


@pytest.mark.parametrize("colour", [False, True])
def test_fpmMain_end_to_end(tmp_path, colour):
    """The reference's entry point on a directory of TIFFs: `fpmMain <dataset.json> <itrCount>`; with isColor
    (dataset_cellScope.json) on 3-sample TIFFs the red plane is reconstructed (fpmMain.cpp:109-116)."""
    import json
    from test_host import write_tiff16
    c = T.Case("cfg1_mono_np64", 9, 20)
    j = json.load(open(os.path.join(T.GOLD, "cfg1_mono_np64.embedded.json")))
    root = tmp_path / "frames"
    root.mkdir()
    j.update(datasetRoot=str(root) + "/", cropX=5, cropY=7, bk1cropX=0, bk1cropY=0, bk2cropX=0, bk2cropY=0, bgThresh=0,
             isColor=colour)
    (tmp_path / "d.json").write_text(json.dumps(j))
    rng = np.random.default_rng(2)
    for k, n in enumerate(c.order):
        fr = np.zeros((80, 90), np.uint16)
        fr[7:7 + 64, 5:5 + 64] = c.stack[k]
        write_tiff16(str(root / ("iLED_%04d.tif" % n)), _colour_frame(fr, rng) if colour else fr)
    out = tmp_path / "out"
    out.mkdir()
    exe = os.path.join(T.ROOT, "fpm-opencv_b200", "bin", "fpmMain")
    env = dict(os.environ, OPENCV_OPENCL_DEVICE="GPU:0")
    r = subprocess.run([exe, str(tmp_path / "d.json"), "3", str(out)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    for line in ("Loading Images...", "resImprovementFactor: 4", "Iteration 3 Completed (Time:", "FP Processing Completed (Time:"):
        assert line in r.stdout
    assert r.stdout.count("Loaded: iLED_") == 20
    # result files: amplitude of the object against the oracle
    import fpmhost  # noqa: F401
    data = open(out / "object_amp.tif", "rb").read()
    amp = np.frombuffer(data[8:8 + 4 * c.L * c.L], np.float32).reshape(c.L, c.L)
    st = c.oracle_run(3)
    assert orc.rel_l2(amp, np.abs(orc.obj_crop(st))) < FULL_TOL


@pytest.mark.parametrize("gpus,colour", [("", False), ("", True), ("0,1", False)])
def test_fpmMain_full_fov(tmp_path, gpus, colour):
    """FPM_FOV_OVERLAP: the whole frame tiled, every frame read once and cut on the device, mosaic written; checked
    against per-tile oracle runs on the host loader's preprocessing of the same frames.  FPM_GPUS=0,1 shards the tiles
    over two GPUs of the box (final peer-copy gather on the first).  colour: isColor + 3-sample TIFFs through
    fpmb200_ingest_frame (the red plane is what the reference keeps)."""
    if gpus:
        import torch
        if torch.cuda.device_count() < 2:
            pytest.skip("needs 2 GPUs")
    import json
    import fpmhost
    import synth
    from test_host import write_tiff16
    c = T.Case("cfg1_mono_np64", 9, 12)
    N, L, nx, ny, ov = c.N, c.L, 3, 2, 16
    step = N - ov
    W, H = (nx - 1) * step + N + 9, (ny - 1) * step + N + 5
    j = json.load(open(os.path.join(T.GOLD, "cfg1_mono_np64.embedded.json")))
    root = tmp_path / "frames"
    root.mkdir()
    bk1, bk2 = (W - N, 0), (0, H - N)
    j.update(datasetRoot=str(root) + "/", cropX=0, cropY=0, bk1cropX=bk1[0], bk1cropY=bk1[1], bk2cropX=bk2[0], bk2cropY=bk2[1],
             bgThresh=120, isColor=colour)
    (tmp_path / "d.json").write_text(json.dumps(j))
    # one big synthetic object: a low-res frame per LED = a wide stack cut from independent tiles' forward models
    frames = np.zeros((len(c.order), H, W), np.uint16)
    for t in range(6):
        st = synth.synth_stack(N, L, c.r, c.cx, c.cy, 50 + t)
        x0, y0 = (t % 3) * 70, (t // 3) * 60
        for k in range(len(c.order)):
            h, w = min(N, H - y0), min(N, W - x0)
            frames[k, y0:y0 + h, x0:x0 + w] = st[k][:h, :w] // 2
    frames += 100
    rng = np.random.default_rng(4)
    for k, n in enumerate(c.order):
        write_tiff16(str(root / ("iLED_%04d.tif" % n)), _colour_frame(frames[k], rng) if colour else frames[k])
    out = tmp_path / "out"
    out.mkdir()
    exe = os.path.join(T.ROOT, "fpm-opencv_b200", "bin", "fpmMain")
    env = dict(os.environ, OPENCV_OPENCL_DEVICE="GPU:0", FPM_FOV_OVERLAP=str(ov))
    if gpus:
        env["FPM_GPUS"] = gpus
    r = subprocess.run([exe, str(tmp_path / "d.json"), "2", str(out)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Full FOV: %dx%d frame -> 3x2 tiles of 64 (overlap 16), 12 LEDs, %d GPU(s)" % (W, H, 2 if gpus else 1) in r.stdout
    assert r.stdout.count("Loaded: iLED_") == 12 and "Iteration 2 Completed (Time:" in r.stdout
    f = L // N
    Wm, Hm = ((nx - 1) * step + N) * f, ((ny - 1) * step + N) * f
    data = open(out / "mosaic_amp.tif", "rb").read()
    got = np.frombuffer(data[8:8 + 4 * Wm * Hm], np.float32).reshape(Hm, Wm)
    crops = []
    for t in range(nx * ny):
        ox, oy = (t % nx) * step, (t // nx) * step
        stack = np.stack([fpmhost.preprocess_frame(frames[k], N, (ox, oy), bk1, bk2, 1, 120)[0] for k in range(len(c.order))])
        st = orc.run(stack, c.cx, c.cy, L, c.r, c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 2, 1)
        crops.append(orc.obj_crop(st))
    ref = orc.mosaic(crops, nx, ny, step, N)
    assert orc.rel_l2(got, ref) < FULL_TOL


# ---- BASELINE configs[2] / [4] (and the shipped dogStomach tile) at FULL size against the float64 oracle -------------
FULL_SIZE = [("cfg3_cellScope_np256", 10), ("cfg5b_cellscope2_np256", 50), ("cfg5_cellscope2_np128", 50),
             ("cfg4s_dogStomach_np200", 10), ("cfg4_dogStomach_np128", 10)]
_full_size_jobs = {}


def _full_size_oracle(name, iters):
    """All full-size C-oracle runs (oracle/fpm_oracle.c, float64; 20-80 s each) start together on first use, in
    threads (ctypes drops the GIL), and overlap each other and the GPU work of the tests that wait for them."""
    import concurrent.futures as cf
    import c_oracle
    if not _full_size_jobs:
        pool = cf.ThreadPoolExecutor(len(FULL_SIZE))

        def job(nm, it):
            c = T.case(nm)
            return c_oracle.run(c.stack, c.cx, c.cy, c.L, c.r, c.cfg.delta1, c.cfg.delta2, c.cfg.eps, it, 1)
        for nm, it in FULL_SIZE:
            _full_size_jobs[nm] = pool.submit(job, nm, it)
    return _full_size_jobs[name].result()


@pytest.mark.parametrize("name,iters", FULL_SIZE)
def test_full_size_oracle_parity(name, iters):
    """north_star: "1e-3 after the full iteration count" -- every LED, every iteration of BASELINE configs[2]
    (cellScope dome, Np 256, Nlarge 1536, 241 LEDs x 10), configs[4] (cellscope2, 193 LEDs x 50, Np 128 / Nlarge 512
    and Np 256 / Nlarge 1024), configs[3]'s tile and the shipped dogStomach tile (Np 200, Nlarge 600, 157 LEDs x 10):
    objF, pupil, objCrop (amplitude and masked phase) against the float64 C oracle.  Where two implementations of the
    update exist for the size (one CTA per tile / thread-block cluster) both are checked, and they agree bit for bit."""
    c = T.case(name)
    variants = [None] if c.N not in (128, 256) else [1, 8 if c.N == 256 else 4]
    got = []
    for ctas in variants:
        ctx = c.make_ctx(cluster=ctas)
        ctx.run(iters)
        ctx.finalize()
        got.append((ctx, ctx.download(0)))
    st = _full_size_oracle(name, iters)
    for ctx, _ in got:
        e = compare(ctx, st)
        note("full size %s %d iterations x %d LEDs = %d updates: rel-L2 objF %.2e pupil %.2e [%s]" % (
            name, iters, len(c.cx), iters * len(c.cx), e[0], e[1], ctx.variant))
    if len(got) == 2:
        assert "cluster_kernel" in got[1][0].variant and "cluster_kernel" not in got[0][0].variant
        if "pruned radix-16" in got[0][0].variant:
            # the one-CTA kernel's pruned first butterfly layer adds in another order than the cluster kernel's full one
            for x, y in zip(got[0][1], got[1][1]):
                assert orc.rel_l2(x, y) < 1e-5
        else:
            # same butterflies in the same order per element, exact maxima: how a tile is spread over SMs does not change a bit
            for x, y in zip(got[0][1], got[1][1]):
                assert np.array_equal(x, y)
    for ctx, _ in got:
        ctx.close()


def test_support_reupload_rebuilds_the_unfused_graph(monkeypatch):
    """Np = 200 replays a captured CUDA graph per pass over the LEDs; the graph bakes in the bounding box of the pupil
    support.  Uploading another support on the same context must not replay the old one."""
    monkeypatch.setenv("FPMB200_GENERAL_UNFUSED", "1")
    c = T.Case("cfg4s_dogStomach_np200", 4, 10)
    ctx = c.make_ctx()
    assert "unfused" in ctx.variant
    ctx.run(1)
    N = c.N
    y, x = np.mgrid[0:N, 0:N]
    yw, xw = np.where(y < N // 2, y, y - N), np.where(x < N // 2, x, x - N)
    S = ((((xw + 30) / 20.0) ** 2 + ((yw - 12) / 33.0) ** 2) <= 1).astype(np.float32)
    ctx.upload_pupil_support(S)
    st = orc.State(orc.init_state(c.stack, c.L, c.r).objFc, S.astype(np.complex128), S.astype(np.float64))
    ctx.upload_state(0, T.corner(st.objFc), st.P)
    for k in range(10):
        orc.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
    ctx.run(1)
    ctx.finalize()
    compare(ctx, st)
    ctx.close()
