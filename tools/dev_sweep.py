"""Developer script: update rate of every shipped configuration (product library, CUDA events via torch)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import torch
import fpmb200
if os.environ.get("FPM_LIB"):
    fpmb200.lib_path = lambda: os.path.join(fpmb200.LIB_DIR, os.environ["FPM_LIB"])
import fpm_testlib as T
names = sys.argv[1:] or ["cfg1_mono_np64", "cfg2_fLEDc_np128", "cfg3b_cellScope_np64", "cfg3_cellScope_np256",
                         "cfg4_dogStomach_np128", "cfg5_cellscope2_np128", "cfg5b_cellscope2_np256"]
for name in names:
    c = T.Case(name, 1)
    for n_tiles in (1, 148):
        try:
            ctx = c.make_ctx(n_tiles=n_tiles)
        except Exception as e:
            print(name, n_tiles, "alloc failed:", e); continue
        ctx.run(1); ctx.sync()
        t0 = time.perf_counter(); ctx.run(2); ctx.sync(); dt = time.perf_counter() - t0
        n_upd = 2 * len(c.cx) * n_tiles
        print("%-26s tiles %4d  %8.2f us/update/tile  %10.0f upd/s   %s" % (name, n_tiles, dt / (2 * len(c.cx)) * 1e6, n_upd / dt, ctx.variant))
        ctx.close()
