import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import fpmb200
fpmb200.lib_path = lambda: os.path.join(fpmb200.LIB_DIR, os.environ.get("FPM_LIB", "libfpmb200_timing.so"))
import fpm_testlib as T
c = T.Case(sys.argv[1] if len(sys.argv) > 1 else "cfg1_mono_np64", 1, 8)
ctx = c.make_ctx()
print(ctx.variant)
ctx.step(0, 0)
if False:
    ctx.run(0)
else:
    pass
buf = (C.c_longlong * 16)()
ctx.L.fpmb200_stage_clocks.argtypes = [C.c_void_p, C.c_void_p]
print("rc", ctx.L.fpmb200_stage_clocks(ctx._h, buf))
print([hex(v) for v in list(buf)[11:16]])
