"""Developer script: quick GPU-vs-oracle parity and timing (run under gpurun)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "fpm-opencv_b200"))
import numpy as np
import fpm_oracle as o
import fpmb200
if os.environ.get("FPM_LIB"): fpmb200.lib_path = lambda: os.path.join(fpmb200.LIB_DIR, os.environ["FPM_LIB"])
import synth

def setup(name, seed):
    g = json.load(open(os.path.join(ROOT, "tests/golden/geometry_%s.json" % name)))
    j = o.load_json_lenient(os.path.join(ROOT, "tests/golden/%s.embedded.json" % name))
    cfg = o.config_from_json(j)
    byn = {l["n"]: l for l in g["leds"]}
    cx = [byn[n]["cropX"] for n in g["order"]]; cy = [byn[n]["cropY"] for n in g["order"]]
    st = synth.synth_stack(cfg.Np, cfg.Nlarge, cfg.naRadius, cx, cy, seed)
    return cfg, cx, cy, st

def main():
    names = sys.argv[1:] or ["cfg1_mono_np64", "cfg2_fLEDc_np128"]
    for name in names:
        cfg, cx, cy, st = setup(name, 1234)
        N, L, n = cfg.Np, cfg.Nlarge, len(cx)
        ctx = fpmb200.Context(0)
        ctx.tiles_alloc(1, N, L, n)
        ctx.set_params(cfg.delta1, cfg.delta2, cfg.eps, 1)
        if os.environ.get("FPM_CLUSTER"): ctx.set_cluster(int(os.environ["FPM_CLUSTER"]))
        ctx.upload_leds(cx, cy)
        ctx.upload_pupil_support(o.pupil_support(N, cfg.naRadius))
        print(name, ctx.variant, flush=True)
        ctx.upload_stack(0, st)
        ctx.init_tiles()
        ctx.sync()
        ost = o.init_state(st, L, cfg.naRadius)
        gF, _, gP = ctx.download(0, objCrop=False)
        print(" init objF", o.rel_l2(gF, np.fft.ifftshift(ost.objFc)), "P", o.rel_l2(gP, ost.P), flush=True)
        # per-step from the oracle's state
        worst = 0
        for k in range(min(n, int(os.environ.get("FPM_STEPS", "40")))):
            ctx.upload_state(0, np.fft.ifftshift(ost.objFc), ost.P)
            o.update(ost, st[k], cx[k], cy[k], cfg.delta1, cfg.delta2, cfg.eps, 1)
            ctx.step(0, k)
            gF, _, gP = ctx.download(0, objCrop=False)
            eF, eP = o.rel_l2(gF, np.fft.ifftshift(ost.objFc)), o.rel_l2(gP, ost.P)
            worst = max(worst, eF, eP)
            if k < 3 or not np.isfinite(eF): print("  step", k, eF, eP, flush=True)
        print(" per-step worst rel-L2", worst, flush=True)
        # full run
        iters = int(os.environ.get("FPM_ITERS", "10"))
        ctx.init_tiles(); ctx.sync()
        t = time.perf_counter(); ctx.run(iters); ctx.sync(); dt = time.perf_counter() - t
        ctx.finalize(); ctx.sync()
        gF, gC, gP = ctx.download(0)
        t = time.perf_counter(); ost = o.run(st, cx, cy, L, cfg.naRadius, cfg.delta1, cfg.delta2, cfg.eps, iters, 1); dto = time.perf_counter() - t
        oc = o.obj_crop(ost)
        print(" full: objF", o.rel_l2(gF, np.fft.ifftshift(ost.objFc)), "P", o.rel_l2(gP, ost.P), "objCrop", o.rel_l2(gC, oc),
              "| gpu %.1f us/update (%.0f upd/s), oracle %.0f upd/s" % (dt / (iters * n) * 1e6, iters * n / dt, iters * n / dto), flush=True)
        ctx.close()

if __name__ == "__main__":
    main()
