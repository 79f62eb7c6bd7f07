"""1:1 op-sequence mirror of the reference's runFPM() on OpenCV's own primitives (via cv2)
-- TEST INFRASTRUCTURE / CPU BASELINE ONLY (never imported by the product).

Every statement of fpmMain.cpp:301-482 is reproduced in order, on CV_64FC2 interleaved
mats, with cv2.dft / cv2.add / cv2.subtract / cv2.multiply / cv2.minMaxLoc / cv2.circle
-- i.e. the reference's `cv::dft` CPU path including its three full-spectrum shifts, the
full-spectrum |objF| and the full-spectrum max per LED.  The un-vendored `cvComplex`
helpers (fpmMain.cpp:15) are restated per SURVEY.md 8c R1-R5 on top of cv2 calls.

Used (a) by tests to pin oracle/fpm_oracle.py (windowed numpy restatement) and to
generate tests/golden/loop_*.npz, (b) by bench.py as the same-host CPU baseline
(`cpu_baseline.kind = "port"`, `--impl reference`).
"""
from __future__ import annotations

import cv2
import numpy as np

# ---- cvComplex restated -------------------------------------------------
def fftShift(src):                       # R2: circular shift by (cols/2, rows/2)
    return np.roll(src, (src.shape[0] // 2, src.shape[1] // 2), axis=(0, 1))


def fft2(src):                           # R1: cv::dft forward, unscaled, complex output
    return cv2.dft(src, flags=cv2.DFT_COMPLEX_OUTPUT)


def ifft2(src):                          # R1: DFT_INVERSE | DFT_SCALE
    return cv2.dft(src, flags=cv2.DFT_INVERSE | cv2.DFT_SCALE | cv2.DFT_COMPLEX_OUTPUT)


def complexAbs(a):                       # R3: 2-channel (|A|, 0)
    re, im = cv2.split(a)
    return cv2.merge([cv2.magnitude(re, im), np.zeros_like(re)])


def complexConj(a):
    re, im = cv2.split(a)
    return cv2.merge([re, cv2.multiply(im, -1.0)])


def complexMultiply(a, b):               # R4: true complex product
    ar, ai = cv2.split(a)
    br, bi = cv2.split(b)
    re = cv2.subtract(cv2.multiply(ar, br), cv2.multiply(ai, bi))
    im = cv2.add(cv2.multiply(ar, bi), cv2.multiply(ai, br))
    return cv2.merge([re, im])


def complexDivide(a, b):                 # R4: A*conj(B)/|B|^2
    ar, ai = cv2.split(a)
    br, bi = cv2.split(b)
    den = cv2.add(cv2.multiply(br, br), cv2.multiply(bi, bi))
    re = cv2.divide(cv2.add(cv2.multiply(ar, br), cv2.multiply(ai, bi)), den)
    im = cv2.divide(cv2.subtract(cv2.multiply(ai, br), cv2.multiply(ar, bi)), den)
    return cv2.merge([re, im])


def add_scalar(a, s, kappa):
    """cv::add(mat2ch, float lvalue): the scalar is replicated to every channel (R5,
    kappa=1).  kappa=0 is the 'intended' reading (real part only)."""
    if kappa:
        return cv2.add(a, np.array([[float(s)]]))          # 1x1 f64 -> all channels
    return cv2.add(a, (float(s), 0.0, 0.0, 0.0))


class Mirror:
    """State = the reference's FPM_Dataset fields objF (DC-at-corner), pupil, pupilSupport."""

    def __init__(self, stack, cropX, cropY, Nlarge, naRadius, delta1, delta2, eps, kappa=1):
        self.stack, self.cropX, self.cropY = stack, cropX, cropY
        self.Np = Np = stack.shape[1]
        self.L = L = Nlarge
        self.delta1, self.delta2, self.eps, self.kappa = float(delta1), float(delta2), float(eps), kappa
        # :301-313
        planes0 = np.zeros((Np, Np), np.float64)
        cv2.circle(planes0, (Np // 2, Np // 2), int(naRadius), 1.0, -1, 8, 0)
        planes0 = fftShift(planes0)
        self.pupil = cv2.merge([planes0, np.zeros((Np, Np), np.float64)])
        self.pupilSupport = self.pupil.copy()
        # :319-327
        tmp = stack[1].astype(np.float64)
        complexI = cv2.merge([cv2.sqrt(tmp), np.zeros((Np, Np), np.float64)])
        complexI = fft2(complexI)
        complexI = complexMultiply(complexI, self.pupilSupport)
        complexI = fftShift(complexI)
        # :330-343
        objF = np.zeros((L, L, 2), np.float64)
        o = L // 2 - Np // 2
        objF[o:o + Np, o:o + Np] = complexI
        self.objF = fftShift(objF)
        self.objCrop = None

    def update(self, k):
        Np, kap = self.Np, self.kappa
        xs, ys = int(self.cropX[k]), int(self.cropY[k])
        objF_centered = fftShift(self.objF)                                   # :358
        Objfcrop = fftShift(objF_centered[ys:ys + Np, xs:xs + Np])            # :361-362
        ObjfcropP = complexMultiply(Objfcrop, self.pupil)                     # :364
        ObjcropP = ifft2(ObjfcropP)                                           # :365
        objectAmp = cv2.merge([np.sqrt(self.stack[k].astype(np.float64)),     # :378-387
                               np.zeros((Np, Np), np.float64)])
        tmp1 = add_scalar(ObjcropP, self.eps, kap)                            # :390
        tmp3 = complexAbs(tmp1)                                               # :391
        tmp1 = complexDivide(ObjcropP, tmp3)                                  # :392
        tmp3 = complexMultiply(tmp1, objectAmp)                               # :393
        Objfup = fft2(tmp3)                                                   # :394
        pupil_abs = complexAbs(self.pupil)                                    # :406
        pupil_conj = complexConj(self.pupil)                                  # :407
        tmp1 = complexMultiply(pupil_abs, pupil_conj)                         # :408
        tmp2 = cv2.subtract(Objfup, ObjfcropP)                                # :409
        numerator = complexMultiply(tmp2, tmp1)                               # :410
        pupil_abs_max = cv2.minMaxLoc(pupil_abs.reshape(Np, 2 * Np))[1]       # :415
        pupil_abs_sq = complexMultiply(pupil_abs, pupil_abs)                  # :416
        denomSum = add_scalar(pupil_abs_sq, self.delta2, kap)                 # :417
        tmp1 = cv2.multiply(denomSum, pupil_abs_max)                          # :418
        tmp2 = complexDivide(numerator, tmp1)                                 # :419
        objF_centered = fftShift(self.objF)                                   # :427
        tmp2 = fftShift(tmp2)                                                 # :432
        tmp1 = cv2.add(tmp2, objF_centered[ys:ys + Np, xs:xs + Np])           # :433
        objF_centered[ys:ys + Np, xs:xs + Np] = tmp1                          # :444
        self.objF = fftShift(objF_centered)                                   # :447
        Objfcrop_abs = complexAbs(Objfcrop)                                   # :459
        Objf_abs = complexAbs(self.objF)                                      # :460
        Objfcrop_conj = complexConj(Objfcrop)                                 # :461
        tmp1 = complexMultiply(Objfcrop_abs, Objfcrop_conj)                   # :462
        tmp2 = cv2.subtract(Objfup, ObjfcropP)                                # :463
        numerator = complexMultiply(tmp2, tmp1)                               # :464
        Objf_abs_max = cv2.minMaxLoc(Objf_abs.reshape(self.L, 2 * self.L))[1]  # :467
        Objfcrop_abs_sq = complexMultiply(Objfcrop_abs, Objfcrop_abs)         # :468
        denomSum = add_scalar(Objfcrop_abs_sq, self.delta1, kap)              # :469
        tmp1 = cv2.multiply(denomSum, Objf_abs_max)                           # :470
        tmp2 = complexDivide(numerator, tmp1)                                 # :471
        tmp2 = complexMultiply(tmp2, self.pupilSupport)                       # :472
        self.pupil = cv2.add(self.pupil, tmp2)                                # :475

    def iteration(self):
        for k in range(self.stack.shape[0]):                                  # :348-350
            self.update(k)
        self.objCrop = cv2.dft(self.objF, flags=cv2.DFT_INVERSE | cv2.DFT_SCALE)  # :481

    # views in the oracle's conventions
    def objFc(self):
        c = fftShift(self.objF)
        return c[..., 0] + 1j * c[..., 1]

    def P(self):
        return self.pupil[..., 0] + 1j * self.pupil[..., 1]

    def objCropC(self):
        return self.objCrop[..., 0] + 1j * self.objCrop[..., 1]


def run(stack, cropX, cropY, Nlarge, naRadius, delta1, delta2, eps, iters, kappa=1) -> Mirror:
    m = Mirror(stack, cropX, cropY, Nlarge, naRadius, delta1, delta2, eps, kappa)
    for _ in range(iters):
        m.iteration()
    return m
