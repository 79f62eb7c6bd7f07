#!/usr/bin/env python
"""bench.py -- sub-aperture updates/s of the FPM reconstruction loop (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[3] geometry, `dataset_dogStomach.json` optics with
128x128 tiles (Nlarge 384, 157 LEDs in the reference's NA order, 10 iterations, EPRY pupil update on),
`--tiles-per-gpu` independent tiles per GPU (default 592 = 4 waves of 148 CTAs).  One "step" = one
complete reconstruction of every tile of the rank: spectrum/pupil initialisation, iterations x LEDs
fused updates, final Nlarge x Nlarge inverse FFT.  Tiles are sharded over ranks with no data-path
collective (weak scaling: per-GPU work fixed); the only communication is the timing reduction and,
in the full-FOV leg, the final gather.

`value`   : updates/s with the stacks resident in HBM, CUDA events on the launching stream, max over ranks.
`e2e`     : same metric through the C ABI from pinned HOST buffers: H2D of every stack and D2H of every
            objCrop inside the timed region (chunked; uploads spread over several copy streams).
            `e2e.copy_only` repeats the identical copy schedule WITHOUT the reconstruction, so the line says
            what the host link alone allows (the e2e leg cannot be faster than that).
`roofline`: the fused update kernel against the binding on-chip ceiling, FP32 (SURVEY 8d FLOP count / kernel time,
            peak = FFMA throughput measured on this GPU by bin/fpm_peaks right before the run); `roofline_hbm`
            (SURVEY 8d's algorithmic 18*Np^2 bytes against the measured copy bandwidth, plus the DRAM bytes ncu
            actually saw) and `roofline_smem` are the secondary views.
other legs: `single_tile` (configs[0], [1], [2]: latency of one tile), `stress` (configs[4]: cellscope2, 193 LEDs x 50
            iterations), `full_fov` (configs[3]'s 2560x2160 frame: 320 tiles sharded strongly over the ranks),
            `full_fov_shares` (N = 1: the per-GPU share of that frame at 2 / 4 / 8 GPUs, run on this GPU).
`--impl reference`: the reference's CPU path (1:1 OpenCV op-sequence mirror, oracle/cv2_mirror.py --
            the reference binary itself cannot be built here) on all host cores; loads no
            library of this repo.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, "fpm-opencv_b200")]

import numpy as np  # noqa: E402

CFG_JSON = os.path.join(ROOT, "configs", "cfg4_dogStomach_np128.json")
WORKLOAD = "dogStomach optics (configs[3]), 128x128 tiles, Nlarge 384, 157 LEDs, {iters} iterations, {tiles} tiles/GPU"
METRIC = "sub-aperture updates/sec (aggregate over GPUs)"
UNIT = "updates/s"
FP32_PEAK_TFLOPS_NOMINAL = 148 * 128 * 2 * 1.965e9 / 1e12        # #SM * 128 lanes * 2 * f_max
SMEM_PEAK_TBS_NOMINAL = 148 * 128 * 1.965e9 / 1e12              # #SM * 128 B/clk * f_max


def flops_per_update(N):
    return 20.0 * N * N * np.log2(N) + 86.0 * N * N


def bytes_per_update(N):            # window read + write (fp32 complex) + uint16 intensity
    return 18.0 * N * N


def config_dict(tiles, iters, N=128, L=384, n_leds=157):
    """`config` of the JSON line -- identical for both arms (the driver compares them key by key)."""
    return {"workload": WORKLOAD.format(iters=iters, tiles=tiles), "tiles_per_gpu": tiles, "Np": N, "Nlarge": L,
            "n_leds": n_leds, "iterations": iters, "kappa": 1, "parallelism": "tiles sharded, no collective",
            "l2": "inputs larger than L2 (%.1f GB of stacks + %.1f GB of spectra per GPU)" % (
                tiles * n_leds * N * N * 2 / 1e9, tiles * L * L * 8 / 1e9)}


def geometry(cfg_json=None):
    """LED tables through the product's own host layer (C++ libfpmhost, not the oracle)."""
    import fpmhost
    ds = fpmhost.Dataset(cfg_json or CFG_JSON, 10)
    s = ds.scalars
    n = ds.geometry(1, min(293, s.ledCount) if cfg_json is None else s.ledCount)
    cx, cy = ds.crop_tables()
    return dict(N=s.Np, L=s.Nlarge, r=s.naRadius, n_leds=n, cx=cx, cy=cy, delta1=s.delta1, delta2=s.delta2, eps=s.eps,
                support=fpmhost.pupil_support(s.Np, s.naRadius), order=[int(v) for v in ds.order])


def geometry_reference_arm(name="cfg4_dogStomach_np128"):
    """The same tables for the reference arm WITHOUT loading any library of this repo: the committed golden geometry
    (produced by the reference's own jsoncpp + std::sort, tests/golden/make_golden.py) and the oracle's config reader."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fpm_oracle as orc
    gold = os.path.join(ROOT, "tests", "golden")
    g = json.load(open(os.path.join(gold, "geometry_%s.json" % name)))
    cfg = orc.config_from_json(orc.load_json_lenient(os.path.join(gold, name + ".embedded.json")))
    byn = {l["n"]: l for l in g["leds"]}
    cx = np.array([byn[n]["cropX"] for n in g["order"]], np.int16)
    cy = np.array([byn[n]["cropY"] for n in g["order"]], np.int16)
    return dict(N=cfg.Np, L=cfg.Nlarge, r=cfg.naRadius, n_leds=len(cx), cx=cx, cy=cy, delta1=float(cfg.delta1),
                delta2=float(cfg.delta2), eps=float(cfg.eps))


def distinct_stacks(g, k, seed0=4000):
    import synth
    return [synth.synth_stack(g["N"], g["L"], g["r"], g["cx"], g["cy"], seed0 + i) for i in range(k)]


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons while the timed region runs (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.sm_max, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                 "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def finish(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measure_peaks(device):
    """FP32 FMA and shared-memory load throughput of this GPU, measured now (fpm-opencv_b200/bin/fpm_peaks)."""
    exe = os.path.join(ROOT, "fpm-opencv_b200", "bin", "fpm_peaks")
    try:
        out = subprocess.run([exe, str(device)], capture_output=True, text=True, timeout=60)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:
        return {"error": repr(e)}


def fov_e2e_leg(n_gpus, g, stacks, iters, W=2560, H=2160):
    """Full-FOV reconstruction END TO END through the reference's entry point: `fpmMain <dataset.json> <itrCount>` with
    FPM_FOV_OVERLAP=0 and FPM_GPUS=0..n-1 on a directory of synthetic 16-bit TIFF frames (one 2560x2160 frame per LED,
    157 frames = 1.7 GB): directory scan + LED geometry, TIFF read, device-side tile cut, reconstruction of all 320
    tiles, final gather and amplitude mosaic back in host memory.  Wall clock of the whole process (CUDA context creation
    included) and fpmMain's own phase timing; two passes (the second reads the frames from the page cache)."""
    import shutil
    import tempfile
    import synth
    N, n_leds = g["N"], g["n_leds"]
    need = n_leds * W * H * 2 * 1.1
    root = None
    for base in ("/dev/shm", tempfile.gettempdir()):
        try:
            st = os.statvfs(base)
            if st.f_bavail * st.f_frsize > need:
                root = tempfile.mkdtemp(prefix="fpm_fov_", dir=base)
                break
        except OSError:
            pass
    if root is None:
        return {"error": "no scratch directory with %.1f GB free for the synthetic frames" % (need / 1e9)}
    try:
        t0 = time.perf_counter()
        nx, ny = W // N, H // N
        idx = (np.arange(nx * ny) % len(stacks)).reshape(ny, nx)
        st8 = np.stack(stacks)                                   # [8][n_leds][N][N]
        frame = np.zeros((H, W), np.uint16)
        for k, led in enumerate(g["order"]):
            frame[:ny * N, :nx * N] = st8[idx, k].transpose(0, 2, 1, 3).reshape(ny * N, nx * N)
            synth.write_tiff16(os.path.join(root, "ILED_%04d.tif" % led), frame)
        j = json.load(open(CFG_JSON))
        j.update(datasetRoot=root + "/", cropX=0, cropY=0, bk1cropX=0, bk1cropY=H - N, bk2cropX=W - N, bk2cropY=H - N, bgThresh=0)
        cfg = os.path.join(root, "dataset.json")
        json.dump(j, open(cfg, "w"))
        t_gen = time.perf_counter() - t0
        exe = os.path.join(ROOT, "fpm-opencv_b200", "bin", "fpmMain")
        env = dict(os.environ, OPENCV_OPENCL_DEVICE="GPU:0", FPM_FOV_OVERLAP="0", FPM_GPUS=",".join(str(i) for i in range(n_gpus)))
        passes = []
        for rep in range(2):
            t1 = time.perf_counter()
            r = subprocess.run([exe, cfg, str(iters)], capture_output=True, text=True, env=env, timeout=600)
            wall = time.perf_counter() - t1
            if r.returncode != 0:
                return {"error": "fpmMain exited %d: %s" % (r.returncode, (r.stdout + r.stderr)[-400:])}
            ph = {}
            for ln in r.stdout.splitlines():
                if ln.startswith("Full FOV timing:"):
                    import re
                    m = re.search(r"setup ([0-9.e+-]+) s, load\+ingest ([0-9.e+-]+) s \((\d+) reader threads\), reconstruction ([0-9.e+-]+) s, gather\+mosaic ([0-9.e+-]+) s", ln)
                    if m:
                        ph.update({"setup_s": float(m.group(1)), "load_ingest_s": float(m.group(2)), "reader_threads": int(m.group(3)),
                                   "reconstruction_s": float(m.group(4)), "gather_mosaic_s": float(m.group(5))})
                if ln.startswith("FP Processing Completed"):
                    ph["fpmMain_total_s"] = float(ln.split("Time:")[1].split("sec")[0])
            ph["process_wall_s"] = wall
            passes.append(ph)
        best = min(passes, key=lambda q: q.get("process_wall_s", 1e30))
        return {"what": "fpmMain end to end on %d synthetic %dx%d TIFF frames (%.2f GB), %d tiles, %d GPUs" % (
                    n_leds, W, H, n_leds * W * H * 2 / 1e9, nx * ny, n_gpus),
                "n_gpus": n_gpus, "wall_s": best["process_wall_s"], "fpmMain_total_s": best.get("fpmMain_total_s"),
                "phases": best, "passes": passes, "frames_dir": os.path.dirname(root), "generate_s": t_gen,
                "updates_per_s": nx * ny * n_leds * iters / best["process_wall_s"]}
    except Exception as e:
        return {"error": repr(e)}
    finally:
        shutil.rmtree(root, ignore_errors=True)


# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import fpmb200
    if os.environ.get("FPM_LIB"):
        fpmb200.lib_path = lambda: os.path.join(fpmb200.LIB_DIR, os.environ["FPM_LIB"])
    import sharding
    import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa_cpus = bind_to_gpu_cpus(local)        # before the pinned buffers are allocated (first touch = local node)
    gloo = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo")       # host-side barrier for the leg in which rank 0 drives all GPUs itself
    peaks_live = measure_peaks(local) if rank == 0 else {}
    g = geometry()
    N, L, n_leds, iters = g["N"], g["L"], g["n_leds"], args.iters
    tiles = args.tiles_per_gpu
    ctx = fpmb200.Context(local)
    ctx.tiles_alloc(tiles, N, L, n_leds)
    ctx.set_params(g["delta1"], g["delta2"], g["eps"], 1)
    ctx.upload_leds(g["cx"], g["cy"])
    ctx.upload_pupil_support(g["support"])

    # ---- synthetic input: 8 distinct seeded tiles replicated into a pinned host buffer ----
    per_tile = n_leds * N * N
    in_buf = fpmb200.HostBuffer((tiles, per_tile), np.uint16, write_combined=bool(args.write_combined))
    out_buf = fpmb200.HostBuffer((tiles, L * L * 2), np.float32)
    hin, hout = in_buf.array, out_buf.array
    distinct = distinct_stacks(g, 8, 4000 + 100 * rank)
    for t in range(tiles):
        hin[t] = distinct[t % len(distinct)].reshape(-1)
    ctx.upload_stack_ptr(0, tiles, in_buf.ptr, None)
    ctx.sync()

    # a real (non-NULL) stream: the C ABI treats stream==NULL as "the context's own stream", and
    # torch.cuda.Event only sees work enqueued on the stream it is recorded on
    main = torch.cuda.Stream()
    torch.cuda.set_stream(main)
    sp = main.cuda_stream
    assert sp != 0
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(k_ev=None):
        ctx.init_tiles(0, tiles, 1, sp)
        if k_ev:
            k_ev[0].record(main)
        ctx.run(iters, 0, tiles, sp)
        if k_ev:
            k_ev[1].record(main)
        ctx.finalize(0, tiles, sp)

    # ---- value: inputs resident in HBM ----
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.kernel_launches
    e0, e1 = ev(), ev()
    kev = [(ev(), ev()) for _ in range(args.steps)]
    e0.record(main)
    for s in range(args.steps):
        step_resident(kev[s])
    e1.record(main)
    barrier()
    clocks = sampler.finish()
    launches = ctx.kernel_launches - l0
    ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    ms_max = sharding.max_over_ranks(ms, "cuda")
    updates_per_step_rank = tiles * n_leds * iters
    value = world * updates_per_step_rank * args.steps / (ms_max * 1e-3)

    # ---- e2e: pinned host -> device -> pinned host, chunked; uploads spread over several copy streams ----
    n_in = max(1, args.h2d_streams)
    s_ins = [torch.cuda.Stream() for _ in range(n_in)]
    s_out = torch.cuda.Stream()
    chunk = args.chunk
    sub = max(1, min(args.h2d_tiles, chunk))
    n_chunks = (tiles + chunk - 1) // chunk
    in_bytes, out_bytes = per_tile * 2, L * L * 8

    # Streaming pipeline: chunk c of step s+1 is uploaded while step s still computes (its device buffers are free
    # as soon as chunk c of step s has been reconstructed), results are read back chunk by chunk.  `compute=False`
    # runs the identical copy schedule without the reconstruction (copy-only leg); `d2h=False` drops the read-back.
    state = {"computed": [None] * n_chunks, "fetched": [None] * n_chunks}

    diag_ev = [] if os.environ.get("FPM_E2E_DIAG") else None
    diag_up = []


    def step_e2e(compute=True, d2h=True, sub=sub, n_in=n_in, h2d=True):
        computed, fetched = state["computed"], state["fetched"]
        if computed[0] is None:                    # first step of a leg: uploads start after whatever main has queued so far
            start = torch.cuda.Event()
            start.record(main)
            for si in s_ins:
                si.wait_event(start)
        for c in range(n_chunks):
            a = c * chunk
            n = min(chunk, tiles - a)
            for k, b in enumerate(range(a, a + n, sub)):
                # consecutive upload calls alternate between the copy streams: every upload ends with the 1/I conversion
                # kernel, which cannot start before the persistent update kernel of the chunk in flight leaves the SMs --
                # on a single stream the NEXT chunk's copy would wait behind it
                si = s_ins[(c * ((n + sub - 1) // sub) + k) % n_in]
                if computed[c] is not None:
                    si.wait_event(computed[c])
                if h2d:
                    if diag_ev is not None:
                        u0, u1 = ev(), ev()
                        u0.record(si)
                    ctx.upload_stack_ptr(b, min(sub, a + n - b), in_buf.ptr + b * in_bytes, si.cuda_stream)
                    if diag_ev is not None:
                        u1.record(si)
                        diag_up.append((u0, u1))
            for si in s_ins:
                e_in = torch.cuda.Event()
                e_in.record(si)
                main.wait_event(e_in)
            if fetched[c] is not None:
                main.wait_event(fetched[c])
            if compute:
                ctx.init_tiles(a, n, 1, sp)
                if diag_ev is not None:
                    d0, d1 = ev(), ev()
                    d0.record(main)
                ctx.run(iters, a, n, sp)
                if diag_ev is not None:
                    d1.record(main)
                    diag_ev.append((d0, d1))
                ctx.finalize(a, n, sp)
            e_c = torch.cuda.Event()
            e_c.record(main)
            computed[c] = e_c
            if d2h:
                s_out.wait_event(e_c)
                ctx.download_objcrop_ptr(a, n, out_buf.ptr + a * out_bytes, s_out.cuda_stream)
                e_o = torch.cuda.Event()
                e_o.record(s_out)
                fetched[c] = e_o

    def drain_e2e():
        for e_o in state["fetched"]:
            if e_o is not None:
                main.wait_event(e_o)

    def timed_e2e(n_steps, n_warm, **kw):
        for _ in range(n_warm):
            step_e2e(**kw)
        drain_e2e()
        barrier()
        state["computed"] = [None] * n_chunks
        state["fetched"] = [None] * n_chunks
        a, b = ev(), ev()
        a.record(main)
        for _ in range(n_steps):
            step_e2e(**kw)
        drain_e2e()                  # every result of every timed step is in host memory before the clock stops
        b.record(main)
        barrier()
        state["computed"] = [None] * n_chunks
        state["fetched"] = [None] * n_chunks
        return sharding.max_over_ranks(a.elapsed_time(b), "cuda") / n_steps

    ms_e2e = timed_e2e(args.steps, max(1, args.warmup // 2))
    e2e_value = world * updates_per_step_rank / (ms_e2e * 1e-3)
    checksum = float(np.abs(hout[:: max(1, tiles // 8), :4096]).sum())
    if os.environ.get("FPM_E2E_DIAG"):       # developer: which leg of the pipeline costs what (stderr)
        for lab, kw in (("compute chunked, no copies", dict(h2d=False, d2h=False)), ("compute + D2H", dict(h2d=False)),
                        ("compute + H2D", dict(d2h=False)), ("all", dict())):
            del diag_ev[:]
            del diag_up[:]
            t_leg = timed_e2e(4, 2, **kw)
            if diag_up:
                ups = [a.elapsed_time(b) for a, b in diag_up[-4 * n_chunks:]]
                print("e2e diag: %-28s upload call (copy + conversion) on its stream %.2f ms (min %.2f max %.2f)" % (lab, np.mean(ups), min(ups), max(ups)), file=sys.stderr)
            runs = [a.elapsed_time(b) for a, b in diag_ev[-4 * n_chunks:]]
            print("e2e diag: %-28s %.2f ms/step; update kernel per chunk %.2f ms (min %.2f max %.2f), sum %.2f" % (
                lab, t_leg, np.mean(runs), min(runs), max(runs), np.sum(runs) / 4), file=sys.stderr)
    # the same copies with no reconstruction in between: what the host link alone allows
    ms_copy = timed_e2e(3, 1, compute=False)
    ms_h2d = timed_e2e(3, 1, compute=False, d2h=False)
    # the same bytes as 16 copies per chunk spread over four streams (no reconstruction competes for the SMs here, so
    # the conversion kernels do not stall the copies): does splitting the upload help the host link?
    while len(s_ins) < 4:
        s_ins.append(torch.cuda.Stream())
    ms_h2d_split = timed_e2e(3, 1, compute=False, d2h=False, sub=max(1, (chunk + 15) // 16), n_in=4)
    copy_only = {"ms_per_step": ms_copy, "h2d_only_ms_per_step": ms_h2d, "h2d_only_ms_per_step_4_streams_16_copies_per_chunk": ms_h2d_split,
                 "h2d_gbs_per_gpu": tiles * in_bytes / (ms_h2d * 1e-3) / 1e9,
                 "h2d_gbs_aggregate": world * tiles * in_bytes / (ms_h2d * 1e-3) / 1e9,
                 "both_directions_gbs_aggregate": world * tiles * (in_bytes + out_bytes) / (ms_copy * 1e-3) / 1e9,
                 "updates_per_s_if_copy_bound": world * updates_per_step_rank / (ms_copy * 1e-3),
                 "what": "identical schedule (same pinned buffers, chunks, streams, uint16->1/I conversion kernel) without init/run/finalize"}

    # ---- full-FOV leg (strong scaling): BASELINE configs[3] frame 2560x2160 -> 20x16 = 320 tiles sharded over the
    #      ranks, each rank in a context of its own size (so the library's few-tiles policy sees the real count),
    #      5 timed passes, final gather of objCrop to rank 0 ----
    fov_tiles = len(sharding.tile_grid(2560, 2160, N))
    a, b = sharding.shard_range(fov_tiles, rank, world)
    nl = b - a
    fctx = fpmb200.Context(local)
    fctx.tiles_alloc(max(nl, 1), N, L, n_leds)
    fctx.set_params(g["delta1"], g["delta2"], g["eps"], 1)
    fctx.upload_leds(g["cx"], g["cy"])
    fctx.upload_pupil_support(g["support"])
    for t in range(max(nl, 1)):
        fctx.upload_stack(t, distinct[t % len(distinct)])
    fov_pass = []
    for rep in range(6):                          # first pass = warm-up
        barrier()
        f0, f1 = ev(), ev()
        f0.record(main)
        if nl:
            fctx.init_tiles(0, nl, 1, sp)
            fctx.run(iters, 0, nl, sp)
            fctx.finalize(0, nl, sp)
        f1.record(main)
        torch.cuda.synchronize()
        t_ms = sharding.max_over_ranks(f0.elapsed_time(f1), "cuda")
        if rep:
            fov_pass.append(t_ms)
    gather_ms = 0.0
    if world > 1:
        dev_out = fctx.objcrop_tensor(0, nl) if nl else torch.empty((0, L * L * 2), dtype=torch.float32, device="cuda")
        # first pass untimed (NCCL sets its point-to-point channels up lazily), second pass timed on the device
        for timed in (False, True):
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            full = sharding.gather_tiles(dev_out, fov_tiles, rank, world)     # NCCL send/recv -> rank 0
            g1.record()
            torch.cuda.synchronize()
            del full
            if timed:
                gather_ms = g0.elapsed_time(g1)
    fov_ms = float(np.median(fov_pass))
    gather_s = sharding.max_over_ranks(gather_ms, "cuda") * 1e-3
    fov_variant = fctx.variant
    fctx.close()

    # ---- N = 1 only: the share one GPU holds of the same frame at 2 / 4 / 8 GPUs, run on this GPU in a context of that
    #      size (the library's few-tiles policy picks the kernel: clusters of two CTAs per tile at 40 tiles).  The device
    #      time of the strong-scaling leg at W GPUs is this number plus the gather; the 8-GPU lines of SCALE check it. ----
    fov_shares = None
    if world == 1:
        fov_shares = []
        for w in (2, 4, 8):
            nl_w = (fov_tiles + w - 1) // w
            sctx = fpmb200.Context(local)
            sctx.tiles_alloc(nl_w, N, L, n_leds)
            sctx.set_params(g["delta1"], g["delta2"], g["eps"], 1)
            sctx.upload_leds(g["cx"], g["cy"])
            sctx.upload_pupil_support(g["support"])
            for t in range(nl_w):
                sctx.upload_stack(t, distinct[t % len(distinct)])
            passes = []
            for rep in range(4):                      # first pass = warm-up
                f0, f1 = ev(), ev()
                f0.record(main)
                sctx.init_tiles(0, nl_w, 1, sp)
                sctx.run(iters, 0, nl_w, sp)
                sctx.finalize(0, nl_w, sp)
                f1.record(main)
                torch.cuda.synchronize()
                if rep:
                    passes.append(f0.elapsed_time(f1))
            fov_shares.append({"gpus": w, "tiles_per_gpu": nl_w, "recon_ms": float(np.median(passes)), "kernel": sctx.variant})
            sctx.close()

    # ---- stress leg: BASELINE configs[4] (cellscope2 dome, 128x128 tiles, Nlarge 512, 193 LEDs, 50 iterations),
    #      148 tiles per GPU (one wave), weak scaling like the headline ----
    stress = None
    try:
        g5 = geometry(os.path.join(ROOT, "configs", "cfg5_cellscope2_np128.json"))
        st_tiles, st_iters = 148, 50
        c5 = fpmb200.Context(local)
        c5.tiles_alloc(st_tiles, g5["N"], g5["L"], g5["n_leds"])
        c5.set_params(g5["delta1"], g5["delta2"], g5["eps"], 1)
        c5.upload_leds(g5["cx"], g5["cy"])
        c5.upload_pupil_support(g5["support"])
        s5 = [synth.synth_stack(g5["N"], g5["L"], g5["r"], g5["cx"], g5["cy"], 7000 + 10 * rank + i) for i in range(2)]
        for t in range(st_tiles):
            c5.upload_stack(t, s5[t % 2])
        best = None
        for rep in range(3):
            barrier()
            s0, s1 = ev(), ev()
            s0.record(main)
            c5.init_tiles(0, st_tiles, 1, sp)
            c5.run(st_iters, 0, st_tiles, sp)
            c5.finalize(0, st_tiles, sp)
            s1.record(main)
            torch.cuda.synchronize()
            t_ms = sharding.max_over_ranks(s0.elapsed_time(s1), "cuda")
            if rep:
                best = min(best or 1e30, t_ms)
        stress = {"workload": "cellscope2 optics (configs[4]), 128x128 tiles, Nlarge %d, %d LEDs, %d iterations, %d tiles/GPU" % (
                      g5["L"], g5["n_leds"], st_iters, st_tiles),
                  "ms_per_pass": best, "updates_per_s": world * st_tiles * g5["n_leds"] * st_iters / (best * 1e-3),
                  "scaling": "weak", "kernel": c5.variant}
        c5.close()
    except Exception as e:
        stress = {"error": repr(e)}

    # ---- single-tile legs: BASELINE configs[0], [1], [2] -- latency numbers: the LED order is sequential, one tile can
    #      use one SM (or one thread-block cluster) ----
    single = None
    if rank == 0:
        single = []
        for cfg, what in (("cfg1_mono_np64", "mono optics (configs[0]), one 64x64 tile"),
                          ("cfg2_fLEDc_np128", "fLED-c optics (configs[1]), one 128x128 tile"),
                          ("cfg3_cellScope_np256", "cellScope dome (configs[2]), one 256x256 tile")):
            try:
                g1 = geometry(os.path.join(ROOT, "configs", cfg + ".json"))
                c1 = fpmb200.Context(local)
                c1.tiles_alloc(1, g1["N"], g1["L"], g1["n_leds"])
                c1.set_params(g1["delta1"], g1["delta2"], g1["eps"], 1)
                c1.upload_leds(g1["cx"], g1["cy"])
                c1.upload_pupil_support(g1["support"])
                c1.upload_stack(0, synth.synth_stack(g1["N"], g1["L"], g1["r"], g1["cx"], g1["cy"], 5000))
                best = None
                for rep in range(4):                      # first pass = warm-up
                    c1.init_tiles(0, 1, 1, sp)
                    s0, s1 = ev(), ev()
                    s0.record(main)
                    c1.run(iters, 0, 1, sp)
                    c1.finalize(0, 1, sp)
                    s1.record(main)
                    torch.cuda.synchronize()
                    if rep:
                        best = min(best or 1e30, s0.elapsed_time(s1))
                single.append({"workload": "%s, Nlarge %d, %d LEDs, %d iterations" % (what, g1["L"], g1["n_leds"], iters),
                               "recon_ms": best, "us_per_update": best * 1e3 / (g1["n_leds"] * iters),
                               "updates_per_s": g1["n_leds"] * iters / (best * 1e-3), "kernel": c1.variant})
                c1.close()
            except Exception as e:                        # the headline legs above do not depend on these
                single.append({"workload": what, "error": repr(e)})

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        if "fp32_ffma_tflops" in peaks_live:
            fp32_peak = max(peaks_live["fp32_ffma_tflops"], peaks_live.get("fp32_ffma2_tflops", 0.0))
            smem_peak = peaks_live["smem_lds128_tbs"]
            onchip_src = "measured on this GPU right before the run (bin/fpm_peaks: FFMA/FFMA2 chains, conflict-free LDS.128)"
        else:
            fp32_peak, smem_peak = FP32_PEAK_TFLOPS_NOMINAL, SMEM_PEAK_TBS_NOMINAL
            onchip_src = "nominal 148 SM x 128 lanes x 2 x 1.965 GHz / 148 SM x 128 B/clk (SURVEY 8d); fpm_peaks failed: %s" % peaks_live.get("error")
        upd_per_launch = tiles * n_leds * iters
        ach_gbs = bytes_per_update(N) * upd_per_launch / (kernel_ms * 1e-3) / 1e9
        ach_tf = flops_per_update(N) * upd_per_launch / (kernel_ms * 1e-3) / 1e12
        ach_smem = 48.0 * N * N * upd_per_launch / (kernel_ms * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture
        # (profiles/traffic.json); per launch like `achieved`, scaled by updates when this launch is a different size
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * (upd_per_launch / tj["updates_per_launch"])
        except Exception:
            pass
        cfgd = config_dict(tiles, iters, N, L, n_leds)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (complex64 field, uint16 intensities)", "data": "synthetic (8 seeded tiles/rank replicated; forward model of the reference)",
            "config": cfgd,
            "kernel": ctx.variant,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": tiles * in_bytes * world,
                    "d2h_bytes_per_step": tiles * out_bytes * world, "ms_per_step": ms_e2e,
                    "chunk_tiles": chunk, "h2d_streams": n_in, "h2d_copy_tiles": sub, "write_combined_input": bool(args.write_combined),
                    "checksum": checksum, "host_cpus_bound": numa_cpus, "copy_only": copy_only,
                    "pipeline": "per step: H2D of every stack + reconstruction + D2H of every objCrop; uploads of step s+1 overlap the compute of step s"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": ach_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach_tf / fp32_peak,
                         "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu dram read+write)",
                         "kernel": ctx.variant.split("<")[0], "kernel_ms_per_launch": kernel_ms,
                         "flops_per_update": flops_per_update(N), "updates_per_launch": upd_per_launch,
                         "peak_source": onchip_src, "peak_nominal": FP32_PEAK_TFLOPS_NOMINAL},
            "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                             "algorithmic_bytes_per_update": bytes_per_update(N), "peak_source": peak_src,
                             "dram_gbs_actual": None if traffic is None else traffic / (kernel_ms * 1e-3) / 1e9,
                             "note": "SURVEY 8d counts the whole window in HBM; the window lives in L2, DRAM sees the 1/I stream (dram_gbs_actual)"},
            # SURVEY 8d's shared-memory lower bound (48*Np^2 bytes per update: one transpose per 2-D FFT + stage in/out);
            # the ncu capture under profiles/ has the actual wavefront count
            "roofline_smem": {"bound": "smem", "achieved": ach_smem, "peak": smem_peak, "unit": "TB/s",
                              "frac": ach_smem / smem_peak, "bytes_per_update": 48.0 * N * N, "peak_source": onchip_src},
            "peaks_measured": peaks_live,
            "full_fov": {"frame": "2560x2160", "tiles": fov_tiles, "recon_ms": fov_ms, "recon_ms_passes": fov_pass, "gather_s": gather_s,
                         "updates_per_s": fov_tiles * n_leds * iters / (fov_ms * 1e-3), "scaling": "strong", "kernel": fov_variant},
        }
        if fov_shares is not None:
            line["full_fov_shares"] = fov_shares
        if stress is not None:
            line["stress"] = stress
        if single is not None:
            line["single_tile"] = single
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single(g, distinct[0], iters)
    ctx.close()
    in_buf.close()
    out_buf.close()
    torch.cuda.synchronize()
    # ---- full-FOV end to end (BASELINE metric, second half): rank 0 runs the fpmMain executable over all `world` GPUs
    #      while the other ranks wait on a host-side barrier (their GPUs are idle: every context of this process is closed)
    if gloo is not None:
        dist.barrier(group=gloo)
    if rank == 0:
        if not args.no_fov_e2e:
            line["full_fov_e2e"] = fov_e2e_leg(world, g, distinct, iters)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(group=gloo)
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def cpu_baseline_single(g, stack, iters, budget_s=20.0):
    """Oracle port (OpenCV op-sequence mirror) on ONE host core, bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cv2
    import cv2_mirror
    cv2.setNumThreads(1)
    m = cv2_mirror.Mirror(stack, g["cx"], g["cy"], g["L"], g["r"], g["delta1"], g["delta2"], g["eps"], 1)
    t0 = time.perf_counter()
    n = 0
    while n < iters * g["n_leds"] and time.perf_counter() - t0 < budget_s:
        m.update(n % g["n_leds"])
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "1 tile, %d sequential updates of the same stack (oracle/cv2_mirror.py: cv::dft float64 op sequence of fpmMain.cpp:345-482)" % n}


def _ref_worker(a):
    stack, g, n_updates = a
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cv2
    import cv2_mirror
    cv2.setNumThreads(1)
    m = cv2_mirror.Mirror(stack, g["cx"], g["cy"], g["L"], g["r"], g["delta1"], g["delta2"], g["eps"], 1)
    for k in range(n_updates):
        m.update(k % g["n_leds"])
    return float(np.abs(m.pupil).sum())


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path, all host cores, one tile per
    worker process (tiles are independent).  Under torchrun only rank 0 works.  No library of this repo is loaded:
    geometry comes from the committed golden tables, stacks from numpy."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import multiprocessing as mp
    g = geometry_reference_arm()
    cores = os.cpu_count() or 1
    stacks = distinct_stacks(g, min(cores, 8))
    n_updates = args.ref_updates
    jobs = [(stacks[i % len(stacks)], g, n_updates) for i in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker, [(s, g, max(4, n_updates // 8)) for s, _, _ in jobs])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_ref_worker, jobs)
        dt = time.perf_counter() - t0
    value = cores * n_updates * args.steps / dt
    sample = "%d tiles in parallel (one per core), %d sequential updates each per step" % (cores, n_updates)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 (CV_64FC2, uint16 intensities)", "data": "synthetic",
            "config": config_dict(args.tiles_per_gpu, args.iters, g["N"], g["L"], g["n_leds"]),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_cpus(local):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the e2e leg
    live on the GPU's NUMA node (8 ranks streaming 3 GB per step each otherwise share one socket's memory).
    Returns the number of CPUs in the mask (0 = left unchanged)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = {i for i in range(n) if (mask[i // 64] >> (i % 64)) & 1} & os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tiles-per-gpu", type=int, default=592)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--chunk", type=int, default=148)
    # one upload call per chunk on one stream: measured on one B200, 13 calls of 12 tiles on two streams HALVED e2e
    # (7.96 M vs 13.2 M updates/s): every call ends with the uint16 -> 1/I conversion kernel, which needs an SM and
    # waits for a CTA of the persistent update kernel to retire (16.8 ms) -- the copies queued behind it stall
    ap.add_argument("--h2d-streams", type=int, default=2, help="copy streams the uploads of one chunk are spread over")
    ap.add_argument("--h2d-tiles", type=int, default=148, help="tiles per upload call (12 tiles = 62 MB)")
    ap.add_argument("--write-combined", type=int, default=0, help="1: write-combined pinned input buffer")
    ap.add_argument("--ref-updates", type=int, default=157)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fov-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print("bench.py: note: fewer than 3 warm-up steps requested", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    # stdout carries exactly one JSON line: libraries that write to file descriptor 1 on their own (NCCL prints its
    # version banner there) are pointed at stderr for the duration of the run
    sys.stdout.flush()
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_json_fd, "w", buffering=1)
    main()
