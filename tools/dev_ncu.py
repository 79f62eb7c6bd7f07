"""Developer script: a short launch sequence of one configuration for ncu captures.
usage: python tools/dev_ncu.py <cfg> <n_tiles> [iters per launch] [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import fpm_testlib as T
name, n_tiles = sys.argv[1], int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 1
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 2
c = T.Case(name, 1)
ctx = c.make_ctx(n_tiles=n_tiles)
for _ in range(launches):
    ctx.run(iters)
ctx.sync()
print(ctx.variant)
ctx.close()
