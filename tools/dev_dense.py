import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import fpm_oracle as o, fpm_testlib as T
name, n_leds = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("cfg2_fLEDc_np128", 6)
dense = (len(sys.argv) <= 3) or sys.argv[3] == "dense"
c = T.Case(name, 3, n_leds)
S = np.ones((c.N, c.N), np.float32) if dense else c.support.astype(np.float32)
for rep in range(2):
    ctx = c.make_ctx(support=S)
    print(ctx.variant)
    st = o.init_state(c.stack, c.L, c.r); st.S = S.astype(np.float64)
    for k in range(n_leds):
        ctx.upload_state(0, T.corner(st.objFc), st.P)
        before = st.objFc.copy(); Pb = st.P.copy()
        o.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
        ctx.step(0, k)
        gF, _, gP = ctx.download(0, objCrop=False)
        gFc = np.fft.fftshift(gF)
        eF, eP = o.rel_l2(gFc, st.objFc), o.rel_l2(gP, st.P)
        d = np.abs(gFc - st.objFc); iy, ix = np.unravel_index(d.argmax(), d.shape)
        print(" step", k, "crop", c.cx[k], c.cy[k], "eF %.2e eP %.2e" % (eF, eP), "maxdiff at", iy - c.cy[k], ix - c.cx[k], "%.3g" % d.max(),
              "incr rel err %.2e" % o.rel_l2(gFc - before, st.objFc - before), "P incr rel err %.2e" % o.rel_l2(gP - Pb, st.P - Pb))
    ctx.close()
