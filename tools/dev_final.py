"""Developer script: the planned objCrop transform (fpm_fft2d.cuh) against the run-time-radix path, and init / finalize
timings at the bench tile count."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import fpmb200
if os.environ.get("FPM_LIB"):
    fpmb200.lib_path = lambda: os.path.join(fpmb200.LIB_DIR, os.environ["FPM_LIB"])
import fpm_testlib as T
import fpm_oracle as orc
for name in sys.argv[1:] or ["cfg1_mono_np64", "cfg4_dogStomach_np128", "cfg5_cellscope2_np128"]:
    c = T.Case(name, 1, 12)
    ctx = c.make_ctx()
    ctx.run(1)
    out = {}
    for gen in ("1", "0"):
        os.environ["FPMB200_FINALIZE_GENERIC"] = gen
        ctx.finalize(); ctx.sync()
        out[gen] = ctx.download(0)[1].copy()
    F = ctx.download(0)[0]
    ref = np.fft.ifft2(F.astype(np.complex128))
    print(name, "L", c.L, "planned vs generic rel-L2 %.2e, planned vs numpy %.2e, generic vs numpy %.2e" % (
        orc.rel_l2(out["0"], out["1"]), orc.rel_l2(out["0"], ref), orc.rel_l2(out["1"], ref)), flush=True)
    ctx.close()
n_tiles = int(os.environ.get("FPM_TILES", "592"))
c = T.Case("cfg4_dogStomach_np128", 1)
ctx = c.make_ctx(n_tiles=n_tiles)
for gen in ("1", "0", "1", "0"):
    os.environ["FPMB200_FINALIZE_GENERIC"] = gen
    ctx.finalize(); ctx.sync()
    t0 = time.perf_counter()
    for _ in range(5): ctx.finalize()
    ctx.sync(); dt = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for _ in range(5): ctx.init_tiles()
    ctx.sync(); di = (time.perf_counter() - t0) / 5
    print("generic" if gen == "1" else "planned", "finalize %.3f ms, init %.3f ms (%d tiles)" % (dt * 1e3, di * 1e3, n_tiles), flush=True)
ctx.close()
