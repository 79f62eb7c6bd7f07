import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import fpm_testlib as T
c = T.Case("cfg4_dogStomach_np128", 1)
for n_tiles, cl in ((40, 1), (40, 2), (37, 4), (36, 4), (40, 4), (74, 2), (74, 1), (80, 1), (20, 4), (20, 2), (160, 1)):
    ctx = c.make_ctx(n_tiles=n_tiles, cluster=cl)
    ctx.run(1); ctx.sync()
    t0 = time.perf_counter(); ctx.run(4); ctx.sync(); dt = time.perf_counter() - t0
    print("tiles %3d cluster %d: %.2f ms per 4 iterations, %.2f us/update" % (n_tiles, cl, dt * 1e3, dt / (4 * len(c.cx)) * 1e6), flush=True)
    ctx.close()
