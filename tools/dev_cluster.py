"""Developer script: time per update of one configuration at a given tile count for several CTAs-per-tile choices.
usage: python tools/dev_cluster.py <cfg> <n_tiles> <ctas,ctas,...> [iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import fpmb200
if os.environ.get("FPM_LIB"):
    fpmb200.lib_path = lambda: os.path.join(fpmb200.LIB_DIR, os.environ["FPM_LIB"])
import fpm_testlib as T
name, n_tiles = sys.argv[1], int(sys.argv[2])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
c = T.Case(name, 1)
for ctas in [int(x) for x in sys.argv[3].split(",")]:
    try:
        ctx = c.make_ctx(n_tiles=n_tiles, cluster=ctas)
    except Exception as e:
        print(name, n_tiles, ctas, "failed:", e); continue
    ctx.run(1); ctx.sync()
    best = 1e30
    for rep in range(3):
        t0 = time.perf_counter(); ctx.run(iters); ctx.sync(); best = min(best, time.perf_counter() - t0)
    print("%-24s tiles %4d ctas %d  %8.2f us/update (wall per tile-update)  %8.2f ms per %d iterations  %10.0f upd/s  %s" % (
        name, n_tiles, ctas, best / (iters * len(c.cx)) * 1e6, best * 1e3, iters, iters * len(c.cx) * n_tiles / best, ctx.variant[:70]))
    ctx.close()
