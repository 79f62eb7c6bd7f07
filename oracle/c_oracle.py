"""ctypes wrapper of oracle/_build/liboracle_c.so (plain-C float64 restatement, fpm_oracle.c).
TEST INFRASTRUCTURE ONLY -- the fast checker for full-size runs (BASELINE configs[2] / [4]: thousands of updates
at Np 256), pinned to oracle/fpm_oracle.py by tests/test_oracle.py."""
import ctypes as C
import os

import numpy as np

import fpm_oracle as orc

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "_build", "liboracle_c.so")
        if not os.path.exists(p):
            raise ImportError("build it with `make -C oracle c_oracle`")
        _lib = C.CDLL(p)
        _lib.fpm_oracle_run.restype = C.c_int
        _lib.fpm_oracle_run.argtypes = [C.c_void_p] * 6 + [C.c_int] * 5 + [C.c_double] * 3 + [C.c_int]
        _lib.fpm_oracle_fft1d.restype = C.c_int
        _lib.fpm_oracle_fft1d.argtypes = [C.c_void_p, C.c_int, C.c_int]
    return _lib


def fft1d(x, sign=-1):
    """sign=-1: forward like cv::dft / numpy.fft.fft; +1: unscaled inverse."""
    x = np.array(x, np.complex128)
    load().fpm_oracle_fft1d(x.ctypes.data, x.size, sign)
    return x


def update_inplace(st, stack, cx, cy, n_updates, delta1, delta2, eps, kappa=1, slot_begin=0):
    """`n_updates` sequential updates (slots slot_begin, slot_begin+1, ... wrapping) on an orc.State, in place."""
    lib = load()
    N, L = st.P.shape[0], st.objFc.shape[0]
    objFc = np.ascontiguousarray(st.objFc, np.complex128)
    P = np.ascontiguousarray(st.P, np.complex128)
    S = np.ascontiguousarray(st.S, np.float64)
    stack = np.ascontiguousarray(stack, np.uint16)
    cx = np.ascontiguousarray(cx, np.int16)
    cy = np.ascontiguousarray(cy, np.int16)
    rc = lib.fpm_oracle_run(objFc.ctypes.data, P.ctypes.data, S.ctypes.data, stack.ctypes.data, cx.ctypes.data, cy.ctypes.data,
                            N, L, stack.shape[0], slot_begin, n_updates, float(delta1), float(delta2), float(eps), int(kappa))
    if rc:
        raise ValueError("fpm_oracle_run: Np=%d has a prime factor other than 2, 3, 5" % N)
    st.objFc, st.P = objFc, P
    return st


def run(stack, cropX, cropY, L, naRadius, delta1, delta2, eps, iters, kappa=1, init_slot=1):
    """Same contract as fpm_oracle.run: initialisation (numpy, fpmMain.cpp:301-343) + iters x n_leds updates (C)."""
    st = orc.init_state(stack, L, naRadius, init_slot)
    return update_inplace(st, stack, cropX, cropY, iters * stack.shape[0], delta1, delta2, eps, kappa)
