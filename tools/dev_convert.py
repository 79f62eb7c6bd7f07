"""Developer script: one 148-tile upload (copy + uint16 -> 1/I conversion kernel) for an ncu launch-time capture.
usage: ncu --metrics gpu__time_duration.sum -k regex:stack_convert ... python tools/dev_convert.py [n_tiles]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import fpmb200
if os.environ.get("FPM_LIB"):
    fpmb200.lib_path = lambda: os.path.join(fpmb200.LIB_DIR, os.environ["FPM_LIB"])
import fpm_testlib as T
n_tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 148
c = T.Case("cfg4_dogStomach_np128", 1)
ctx = fpmb200.Context(0)
ctx.tiles_alloc(n_tiles, c.N, c.L, len(c.cx))
buf = fpmb200.HostBuffer((n_tiles, len(c.cx) * c.N * c.N), np.uint16)
for t in range(n_tiles):
    buf.array[t] = c.stack.reshape(-1)
for _ in range(3):
    ctx.upload_stack_ptr(0, n_tiles, buf.ptr, None)
ctx.sync()
ctx.close()
print("done")
