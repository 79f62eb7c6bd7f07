"""Developer script: per-stage cycle breakdown of the update kernel (timing build)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "fpm-opencv_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import fpmb200
fpmb200.lib_path = lambda: os.path.join(fpmb200.LIB_DIR, os.environ.get("FPM_LIB", "libfpmb200_timing.so"))
import fpm_testlib as T
names = sys.argv[1:] or ["cfg2_fLEDc_np128"]
for name in names:
    c = T.Case(name, 1)
    for n_tiles in [int(x) for x in os.environ.get("FPM_TILES", "1,148").split(",")]:
        ctx = c.make_ctx(n_tiles=n_tiles, cluster=int(os.environ["FPM_CLUSTER"]) if os.environ.get("FPM_CLUSTER") else None)
        ctx.run(1); ctx.sync()
        buf = (C.c_longlong * 16)()
        ctx.L.fpmb200_stage_clocks.argtypes = [C.c_void_p, C.c_void_p]
        ctx.L.fpmb200_stage_clocks(ctx._h, buf)
        ctx.run(2); ctx.sync()
        ctx.L.fpmb200_stage_clocks(ctx._h, buf)
        n_upd = 2 * len(c.cx)
        v = np.array(list(buf), dtype=np.float64) / n_upd
        labels = ["-", "S1 colA(inv)+O*P", "S2 colB(inv)", "S3 rowA(inv)", "S4 rowB+amp+rowB'", "S5 rowA'", "S6 colB'", "S7 colA'", "C2 object update", "D max|objF|", "E pupil+next window"]
        if "fpm_update_pruned_kernel" in ctx.variant:
            print(name, "tiles", n_tiles, ctx.variant)
            for k, lab in enumerate(["-", "A P+=Q, O*P (box)", "IR-A inv rows A", "IR-B inv rows B", "IC-A inv cols A (X->S)", "MID inv B + amp + fwd B'",
                                     "FC-A' fwd cols A' (S->X)", "FR-B' fwd rows B'", "FR-A' fwd rows A'", "C object/pupil incr.", "D cell rebuild", "D grid scan"]):
                if k:
                    print("   %-28s %8.0f cyc  %5.1f%%" % (lab, v[k], 100 * v[k] / v[1:16].sum()))
            print("   total %.0f cycles/update" % v[1:16].sum())
            ctx.close()
            continue
        if "general path, fused" in ctx.variant:
            # ticks of fpm_update_general_kernel: 11 = A's element loop, 1 = its barrier + max|P|; 14 = C; 12 = D's cell
            # rebuild, 13 = its barrier, 10 = the grid scan
            labels = ["-", "A barrier, max|P|", "I row stage 1", "I row stage 2", "I col stage 1", "I col stage 2 + M", "F row stage 1",
                      "F row stage 2", "F col stage 1", "F col stage 2", "D grid scan"]
            print(name, "tiles", n_tiles, ctx.variant)
            for k, lab in ((11, "A P+=Q, O*P (box)"), (1, labels[1])) + tuple((k, labels[k]) for k in range(2, 10)) + \
                    ((14, "C object/pupil incr."), (12, "D cell rebuild"), (13, "D barrier"), (10, labels[10])):
                print("   %-22s %8.0f cyc  %5.1f%%" % (lab, v[k], 100 * v[k] / v[1:16].sum()))
            print("   total %.0f cycles/update" % v[1:16].sum())
            ctx.close()
            continue
        if "fpm_update_phased_kernel" in ctx.variant:
            # ticks of fpm_update_phased_kernel; a tick right after a block barrier records its issue, the wait lands in the
            # next tick
            print(name, "tiles", n_tiles, ctx.variant)
            for k, lab in ((9, "top: max|objF| (+ wait of the C barrier)"), (1, "A  P+=Q, O*P, inv col A | cell rebuild"), (10, "B  untouched-cell maximum (+ A barrier)"),
                           (2, "S2 inv col B"), (3, "S3 inv row A"), (4, "S4 inv row B + amplitude + fwd row B'"), (5, "S5 fwd row A'"),
                           (6, "S6 fwd col B'"), (11, "-"), (12, "C  window wait (+ B barrier)"), (13, "C  edge loads, max|P|"),
                           (15, "C  fwd col A' (direct) + object update"), (14, "C  edges -> W, warp max"), (8, "C  fence, barrier issue")):
                print("   %-44s %8.0f cyc  %5.1f%%" % (lab, v[k], 100 * v[k] / v[1:16].sum()))
            print("   total %.0f cycles/update (thread %s of CTA 0)" % (v[1:16].sum(), os.environ.get("FPM_TICK_TID", "0")))
            ctx.close()
            continue
        if "fpm_update_cluster_kernel" in ctx.variant:
            # ticks of fpm_update_cluster_kernel in program order (a tick right after a wait holds the wait)
            print(name, "tiles", n_tiles, ctx.variant)
            tot = v[0:16].sum()
            for k, lab in ((11, "top"), (0, "wbar wait (helper threads: the side work)"), (1, "S1 P+=Q, O*P, inv col A (+ named barrier)"),
                           (14, "S2 inv col B + remote stores"), (2, "   wait: row slab"), (3, "S3 inv row A"), (4, "S4 inv row B + amplitude + fwd row B'"),
                           (15, "S5 fwd row A' + remote stores"), (5, "   wait: column slab"), (6, "S6 fwd col B'"), (7, "S7 fwd col A'"),
                           (8, "C2 object update, Q, forwards"), (12, "D  partial cell maxima of the slice"), (13, "D  reduce + send (warp 0), merge into the owners"),
                           (9, "   release fence + arrivals (duty thread)"), (10, "E  wait: exchange of the rank maxima")):
                print("   %-46s %8.0f cyc  %5.1f%%" % (lab, v[k], 100 * v[k] / tot))
            print("   total %.0f cycles/update (thread 0 of CTA 0)" % tot)
            ctx.close()
            continue
        print(name, "tiles", n_tiles, ctx.variant)
        for k in range(1, 11):
            print("   %-20s %8.0f cyc  %5.1f%%" % (labels[k], v[k], 100 * v[k] / v[1:16].sum()))
        print("   total %.0f cycles/update" % v[1:11].sum(), " extra ticks 11..15:", " ".join("%.0f" % x for x in v[11:16]))
        pass
        ctx.close()
