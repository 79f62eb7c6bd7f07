import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import fpm_oracle as o, fpm_testlib as T
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3b_cellScope_np64"
c = T.case(name)
ctx = c.make_ctx()
print(ctx.variant)
st = o.init_state(c.stack, c.L, c.r)
for k in range(len(c.cx)):
    o.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
for k in range(min(80, len(c.cx))):
    ctx.upload_state(0, T.corner(st.objFc), st.P)
    before = st.objFc.copy(); Pb = st.P.copy()
    o.update(st, c.stack[k], int(c.cx[k]), int(c.cy[k]), c.cfg.delta1, c.cfg.delta2, c.cfg.eps, 1)
    ctx.step(0, k)
    gF, _, gP = ctx.download(0, objCrop=False)
    gFc = np.fft.fftshift(gF)
    eF, eP = o.rel_l2(gFc, st.objFc), o.rel_l2(gP, st.P)
    if eF > 1e-5 or eP > 1e-5:
        d = np.abs(gFc - st.objFc)
        iy, ix = np.unravel_index(d.argmax(), d.shape)
        print("step", k, "crop", c.cx[k], c.cy[k], "eF", eF, "eP", eP, "max diff at", iy, ix, d.max(),
              "oracle delta there", abs(st.objFc[iy, ix] - before[iy, ix]), "gpu delta", abs(gFc[iy, ix] - before[iy, ix]),
              "max|objF|", np.abs(st.objFc).max(), "argmax", np.unravel_index(np.abs(st.objFc).argmax(), d.shape),
              "maxP", np.abs(Pb).max())
        nz = np.argwhere(d > 1e-3 * d.max())
        print("   diff bbox rows", nz[:, 0].min(), nz[:, 0].max(), "cols", nz[:, 1].min(), nz[:, 1].max(), "count", len(nz))
        ratio = (gFc - before)[nz[:, 0], nz[:, 1]] / (st.objFc - before)[nz[:, 0], nz[:, 1]]
        print("   gpu/oracle increment ratio: median", np.median(np.abs(ratio)), "min", np.abs(ratio).min(), "max", np.abs(ratio).max())
        dP = np.abs(gP - st.P); print("   P diff max", dP.max(), "P incr ratio", np.median(np.abs((gP - Pb)[c.support > 0] / ((st.P - Pb)[c.support > 0] + 1e-300))))
        break
else:
    print("all steps fine")
