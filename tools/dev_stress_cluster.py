"""Developer script: repeated persistent launches of the cluster kernels (every shape the library picks them for, several
tile counts so that clusters run in waves); results must be identical from run to run.  Run it under `timeout`: a launch
that does not finish is a protocol bug (a signal handler cannot interrupt the blocking CUDA call).
usage: timeout 120 python tools/dev_stress_cluster.py [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import fpmb200
if os.environ.get("FPM_LIB"):
    fpmb200.lib_path = lambda: os.path.join(fpmb200.LIB_DIR, os.environ["FPM_LIB"])
import fpm_testlib as T
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for name, ctas, tiles in (("cfg2_fLEDc_np128", 4, (1, 5, 37, 60)), ("cfg2_fLEDc_np128", 2, (1, 74, 100)), ("cfg4_dogStomach_np128", 4, (1, 36)),
                          ("cfg5_cellscope2_np128", 4, (1, 7, 40)), ("cfg5b_cellscope2_np256", 8, (1, 3, 16, 20)), ("cfg3_cellScope_np256", 8, (1, 18))):
    c = T.Case(name, 1)
    for n_tiles in tiles:
        ctx = c.make_ctx(n_tiles=n_tiles, cluster=ctas)
        ref = None
        t0 = time.perf_counter()
        for r in range(reps):
            ctx.init_tiles(); ctx.run(2); ctx.sync()
            got = ctx.download(n_tiles - 1, objCrop=False)
            if ref is None: ref = got
            else: assert all(np.array_equal(a, b) for a, b in zip(got, ref) if a is not None), "run %d differs from run 0" % r
        print("%-24s ctas %d tiles %3d: %d x 2 iterations identical, %.2f s  %s" % (name, ctas, n_tiles, reps, time.perf_counter() - t0, ctx.variant[:60]), flush=True)
        ctx.close()
