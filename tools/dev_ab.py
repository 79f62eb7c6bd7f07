"""Developer script: A/B of product-library builds on the bench workload shape (cfg4, many tiles, 10 iterations).
usage: FPM_TILES=592 python tools/dev_ab.py libA.so libB.so ...   (each timed FPM_REPS times, interleaved)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import torch
import fpmb200
import fpm_testlib as T
libs = sys.argv[1:]
name = os.environ.get("FPM_CFG", "cfg4_dogStomach_np128")
n_tiles = int(os.environ.get("FPM_TILES", "592"))
iters = int(os.environ.get("FPM_ITERS", "10"))
reps = int(os.environ.get("FPM_REPS", "3"))
c = T.Case(name, 1)
res = {l: [] for l in libs}
for rep in range(reps):
    for l in libs:
        fpmb200.lib_path = lambda l=l: os.path.join(fpmb200.LIB_DIR, l)
        fpmb200._lib = None
        ctx = c.make_ctx(n_tiles=n_tiles)
        ctx.run(1); ctx.sync()
        t0 = time.perf_counter(); ctx.run(iters); ctx.sync(); dt = time.perf_counter() - t0
        res[l].append(iters * len(c.cx) * n_tiles / dt)
        ctx.close()
for l in libs:
    print("%-28s %s  best %.4f M upd/s" % (l, " ".join("%.4f" % (v / 1e6) for v in res[l]), max(res[l]) / 1e6))
