"""ctypes wrapper of oracle/_build/liboracle_c.so (plain-C float64 restatement, fpm_oracle.c).
TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def load():
    global _lib
    if _lib is None:
        p = os.path.join(_HERE, "_build", "liboracle_c.so")
        if not os.path.exists(p):
            raise ImportError("build it with `make -C oracle c_oracle`")
        _lib = C.CDLL(p)
        _lib.fpm_oracle_run.restype = None
        _lib.fpm_oracle_run.argtypes = [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_double] * 3 + [C.c_int]
    return _lib


def run(objFc, P, S, stack, cx, cy, L, n_updates, delta1, delta2, eps, kappa=1):
    """In-place `n_updates` sequential updates on (objFc [L][L] c128 centred, P [N][N] c128)."""
    lib = load()
    N = P.shape[0]
    assert objFc.dtype == np.complex128 and P.dtype == np.complex128 and objFc.flags.c_contiguous and P.flags.c_contiguous
    S = np.ascontiguousarray(S, np.float64)
    stack = np.ascontiguousarray(stack, np.uint16)
    cx = np.ascontiguousarray(cx, np.int16)
    cy = np.ascontiguousarray(cy, np.int16)
    lib.fpm_oracle_run(objFc.ctypes.data, P.ctypes.data, S.ctypes.data, stack.ctypes.data, cx.ctypes.data, cy.ctypes.data,
                       N, L, stack.shape[0], n_updates, float(delta1), float(delta2), float(eps), int(kappa))
