"""ctypes binding of include/fpmhost.h (libfpmhost.so) -- test / bench plumbing for the C++ host
layer (dataset JSON, LED geometry, LED order, image loader).  No CUDA dependency."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


class Scalars(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "Np", "Nlarge", "Mlarge", "resImprovementFactor", "naRadius", "ledCount", "ledUsedCount",
        "cropX", "cropY", "bk1cropX", "bk1cropY", "bk2cropX", "bk2cropY", "darkfieldExpMultiplier",
        "flipIlluminationX", "flipIlluminationY", "color", "itrCount", "parse_ok")] + [
        (n, C.c_float) for n in ("ps_eff", "du", "lambda_", "objectiveNA", "maxIlluminationNA", "delta1",
                                 "delta2", "bgThreshold", "eps", "ps")] + [("arrayRotation", C.c_double)]


class Led(C.Structure):
    _fields_ = [("led_num", C.c_int32), ("used", C.c_int32), ("sinTheta_x", C.c_double), ("sinTheta_y", C.c_double),
                ("uled", C.c_float), ("vled", C.c_float), ("illumination_na", C.c_float)] + [
        (n, C.c_int16) for n in ("idx_u", "idx_v", "cropXStart", "cropXEnd", "cropYStart", "cropYEnd", "bg_val", "pad_")]


EXPORTS = ["fpmhost_last_error", "fpmhost_open", "fpmhost_close", "fpmhost_geometry", "fpmhost_load",
           "fpmhost_get_scalars", "fpmhost_get_order", "fpmhost_get_led", "fpmhost_get_image",
           "fpmhost_geometry_source", "fpmhost_pupil_support", "fpmhost_preprocess_frame", "fpmhost_tile_grid",
           "fpmhost_device_from_env"]


def lib_path():
    return os.path.join(_HERE, "lib", "libfpmhost.so")


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path()):
            raise ImportError("%s not built (make -C fpm-opencv_b200)" % lib_path())
        L = C.CDLL(lib_path())
        vp, i = C.c_void_p, C.c_int
        L.fpmhost_last_error.restype = C.c_char_p
        L.fpmhost_open.argtypes = [C.c_char_p, i, C.POINTER(vp)]
        L.fpmhost_close.argtypes = [vp]
        L.fpmhost_close.restype = None
        L.fpmhost_geometry.argtypes = [vp, i, i]
        L.fpmhost_load.argtypes = [vp]
        L.fpmhost_get_scalars.argtypes = [vp, C.POINTER(Scalars)]
        L.fpmhost_get_order.argtypes = [vp, vp, i]
        L.fpmhost_get_led.argtypes = [vp, i, C.POINTER(Led)]
        L.fpmhost_get_image.argtypes = [vp, i, vp]
        L.fpmhost_geometry_source.argtypes = [vp]
        L.fpmhost_geometry_source.restype = C.c_char_p
        L.fpmhost_pupil_support.argtypes = [i, i, vp]
        L.fpmhost_preprocess_frame.argtypes = [vp] + [i] * 11 + [vp, vp]
        L.fpmhost_tile_grid.argtypes = [i, i, i, i, vp, vp]
        _lib = L
    return _lib


class Dataset:
    def __init__(self, json_path: str, itr_count: int = 10):
        self.L = load()
        h = C.c_void_p()
        self._h = None
        rc = self.L.fpmhost_open(json_path.encode(), itr_count, C.byref(h))
        if rc != 0:
            raise RuntimeError("fpmhost_open: " + self.L.fpmhost_last_error().decode())
        self._h = h

    def close(self):
        if self._h:
            self.L.fpmhost_close(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def geometry(self, first, last) -> int:
        n = self.L.fpmhost_geometry(self._h, first, last)
        if n < 0:
            raise RuntimeError("fpmhost_geometry: " + self.L.fpmhost_last_error().decode())
        return n

    def load_images(self) -> int:
        return self.L.fpmhost_load(self._h)

    @property
    def scalars(self) -> Scalars:
        s = Scalars()
        self.L.fpmhost_get_scalars(self._h, C.byref(s))
        return s

    @property
    def order(self):
        buf = np.zeros(70000, np.int16)
        n = self.L.fpmhost_get_order(self._h, C.c_void_p(buf.ctypes.data), len(buf))
        return buf[:n].copy()

    def led(self, n) -> Led:
        l = Led()
        if self.L.fpmhost_get_led(self._h, n, C.byref(l)) != 0:
            raise RuntimeError(self.L.fpmhost_last_error().decode())
        return l

    def image(self, n):
        Np = self.scalars.Np
        a = np.empty((Np, Np), np.uint16)
        if self.L.fpmhost_get_image(self._h, n, C.c_void_p(a.ctypes.data)) != 0:
            raise RuntimeError(self.L.fpmhost_last_error().decode())
        return a

    @property
    def geometry_source(self) -> str:
        return self.L.fpmhost_geometry_source(self._h).decode()

    def crop_tables(self):
        """cropXStart/cropYStart in update order (what fpmb200_upload_leds takes)."""
        o = self.order
        cx = np.array([self.led(int(n)).cropXStart for n in o], np.int16)
        cy = np.array([self.led(int(n)).cropYStart for n in o], np.int16)
        return cx, cy


def pupil_support(Np, radius):
    m = np.zeros((Np, Np), np.float32)
    load().fpmhost_pupil_support(Np, radius, C.c_void_p(m.ctypes.data))
    return m


def preprocess_frame(frame, Np, crop, bk1, bk2, divisor, bg_threshold):
    """loadFPMDataset's per-frame preprocessing (fpmMain.cpp:124-144) -> (image [Np][Np] uint16, bg_val)."""
    f = np.ascontiguousarray(frame, dtype=np.uint16)
    out = np.zeros((Np, Np), np.uint16)
    bg = C.c_int(0)
    L = load()
    rc = L.fpmhost_preprocess_frame(C.c_void_p(f.ctypes.data), f.shape[1], f.shape[0], Np, int(crop[0]), int(crop[1]),
                                    int(bk1[0]), int(bk1[1]), int(bk2[0]), int(bk2[1]), int(divisor), int(bg_threshold),
                                    C.c_void_p(out.ctypes.data), C.byref(bg))
    if rc != 0:
        raise RuntimeError("fpmhost_preprocess_frame: " + L.fpmhost_last_error().decode())
    return out, bg.value


def tile_grid(width, height, Np, overlap=0):
    """(nx, ny, xs, ys): regular grid of tile ROI origins, tile index = iy*nx + ix."""
    nx, ny = C.c_int(0), C.c_int(0)
    L = load()
    if L.fpmhost_tile_grid(width, height, Np, overlap, C.byref(nx), C.byref(ny)) != 0:
        raise RuntimeError("fpmhost_tile_grid: " + L.fpmhost_last_error().decode())
    step = Np - overlap
    xs = np.tile(np.arange(nx.value) * step, ny.value).astype(np.int32)
    ys = np.repeat(np.arange(ny.value) * step, nx.value).astype(np.int32)
    return nx.value, ny.value, xs, ys
