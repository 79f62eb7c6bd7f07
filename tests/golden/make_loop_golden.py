#!/usr/bin/env python
"""Generates tests/golden/loop_*.npz: states produced by the 1:1 cv2 op-sequence mirror of
fpmMain.cpp:301-482 (oracle/cv2_mirror.py, i.e. OpenCV's own cv::dft / arithmetic) on small seeded
problems.  They pin oracle/fpm_oracle.py on machines without cv2 and are compared with the CUDA path
in tests/test_gpu_parity.py.  Stored as complex64 to keep the fixtures small (the comparison
tolerances are >= 1e-6)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path[:0] = [os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import cv2_mirror  # noqa: E402
import fpm_testlib as T  # noqa: E402

SPECS = [("cfg1_mono_np64", 7, 12, 2, 1), ("cfg1_mono_np64", 8, 12, 2, 0)]

for name, seed, n_leds, iters, kappa in SPECS:
    c = T.Case(name, seed, n_leds)
    m = cv2_mirror.run(c.stack, c.cx, c.cy, c.L, c.r, c.cfg.delta1, c.cfg.delta2, c.cfg.eps, iters, kappa)
    out = os.path.join(ROOT, "tests", "golden", "loop_%s_s%d_k%d.npz" % (name, seed, kappa))
    np.savez_compressed(out, name=name, seed=seed, n_leds=n_leds, iters=iters, kappa=kappa,
                        objF=np.fft.ifftshift(m.objFc()).astype(np.complex64), pupil=m.P().astype(np.complex64),
                        objCrop=m.objCropC().astype(np.complex64))
    print(out, os.path.getsize(out))
