"""Developer script: fpmb200_run with and without the balanced one-iteration passes (tile counts that are not a multiple
of the SM count): identical results, time per run."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import fpmb200
if os.environ.get("FPM_LIB"):
    fpmb200.lib_path = lambda: os.path.join(fpmb200.LIB_DIR, os.environ["FPM_LIB"])
import fpm_testlib as T
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4_dogStomach_np128"
iters = int(os.environ.get("FPM_ITERS", "10"))
c = T.Case(name, 1)
for n_tiles in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "320,160,200,592").split(",")]:
    res = {}
    for bal in ("0", "1"):
        os.environ["FPMB200_RUN_BALANCED"] = bal
        ctx = c.make_ctx(n_tiles=n_tiles, cluster=1)
        ctx.run(iters); ctx.sync()
        ctx.init_tiles(); ctx.sync()
        l0 = ctx.kernel_launches
        t0 = time.perf_counter(); ctx.run(iters); ctx.sync(); dt = time.perf_counter() - t0
        launches = ctx.kernel_launches - l0
        ctx.finalize(); ctx.sync()
        res[bal] = (dt, [ctx.download(t) for t in (0, n_tiles // 2, n_tiles - 1)], launches)
        ctx.close()
    same = all(np.array_equal(a, b) for ta, tb in zip(res["0"][1], res["1"][1]) for a, b in zip(ta, tb))
    print("%s %4d tiles x %d iterations: one launch %.2f ms (%d launches), balanced %.2f ms (%d launches), identical results: %s" % (
        name, n_tiles, iters, res["0"][0] * 1e3, res["0"][2], res["1"][0] * 1e3, res["1"][2], same), flush=True)
