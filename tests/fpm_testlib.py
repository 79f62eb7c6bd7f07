"""Shared helpers for the tests: golden geometry, synthetic stacks, oracle runs."""
import functools
import json
import os

import numpy as np

import fpm_oracle as orc
import synth

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
GOLD = os.path.join(ROOT, "tests", "golden")
CONFIGS = os.path.join(ROOT, "configs")

ALL_CFGS = ["cfg1_mono_np64", "cfg2_fLEDc_np128", "cfg3_cellScope_np256", "cfg3b_cellScope_np64",
            "cfg4_dogStomach_np128", "cfg4s_dogStomach_np200", "cfg5_cellscope2_np128",
            "cfg5b_cellscope2_np256", "cfg6_mono_dome_np64"]
QUIRKS = ["quirks_rot", "quirks_flipxy"]


def golden(name):
    return json.load(open(os.path.join(GOLD, "geometry_%s.json" % name)))


def embedded_path(name):
    return os.path.join(GOLD, (name + ".json") if name.startswith("quirks") else (name + ".embedded.json"))


class Case:
    """One synthetic single-tile problem derived from a golden geometry."""

    def __init__(self, name, seed, n_leds=None):
        g = golden(name)
        j = orc.load_json_lenient(embedded_path(name))
        self.name, self.cfg = name, orc.config_from_json(j)
        byn = {l["n"]: l for l in g["leds"]}
        order = g["order"] if n_leds is None else g["order"][:n_leds]
        self.order = order
        self.cx = np.array([byn[n]["cropX"] for n in order], np.int16)
        self.cy = np.array([byn[n]["cropY"] for n in order], np.int16)
        c = self.cfg
        self.N, self.L, self.r = c.Np, c.Nlarge, c.naRadius
        self.stack = synth.synth_stack(c.Np, c.Nlarge, c.naRadius, self.cx, self.cy, seed)
        self.support = orc.pupil_support(c.Np, c.naRadius)

    def oracle_run(self, iters, kappa=1, trace=None):
        c = self.cfg
        return orc.run(self.stack, self.cx, self.cy, self.L, self.r, c.delta1, c.delta2, c.eps, iters, kappa, trace=trace)

    def make_ctx(self, n_tiles=1, kappa=1, support=None, device=0, cluster=None):
        import fpmb200
        c = self.cfg
        ctx = fpmb200.Context(device)
        ctx.tiles_alloc(n_tiles, self.N, self.L, len(self.cx))
        ctx.set_params(c.delta1, c.delta2, c.eps, kappa)
        if cluster is not None:
            ctx.set_cluster(cluster)
        ctx.upload_leds(self.cx, self.cy)
        ctx.upload_pupil_support(self.support if support is None else support)
        for t in range(n_tiles):
            ctx.upload_stack(t, self.stack)
        ctx.init_tiles()
        ctx.sync()
        return ctx


class SyntheticCase(Case):
    """A tile size no shipped JSON uses: Np and Nlarge chosen freely, windows on a seeded spiral around the centre
    (brightfield first, like the NA-sorted order of the reference); parameters of cfg7."""

    def __init__(self, N, factor, seed, n_leds, r=None):
        j = orc.load_json_lenient(embedded_path("cfg7_mono_np90"))
        self.name, self.cfg = "synthetic_np%d" % N, orc.config_from_json(j)
        self.N, self.L, self.r = N, N * factor, max(3, N // 6) if r is None else r
        rng = np.random.default_rng(seed)
        c0 = (self.L - N) // 2
        k = np.arange(n_leds)
        rad = np.minimum(c0, 0.35 * self.r * np.sqrt(k))
        ang = 2.4 * k + rng.random(n_leds)
        self.cx = np.clip(np.rint(c0 + rad * np.cos(ang)), 0, self.L - N).astype(np.int16)
        self.cy = np.clip(np.rint(c0 + rad * np.sin(ang)), 0, self.L - N).astype(np.int16)
        self.cx[:2] = c0
        self.cy[:2] = c0
        self.order = list(range(n_leds))
        self.stack = synth.synth_stack(N, self.L, self.r, self.cx, self.cy, seed)
        self.support = orc.pupil_support(N, self.r)


@functools.lru_cache(maxsize=16)
def case(name, seed=1234, n_leds=None):
    return Case(name, seed, n_leds)


def corner(objFc):
    """centred spectrum -> the reference's DC-at-corner objF"""
    return np.fft.ifftshift(objFc)
