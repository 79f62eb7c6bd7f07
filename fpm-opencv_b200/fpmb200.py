"""ctypes binding of the C ABI in include/fpmb200.h (libfpmb200.so) and include/fpmhost.h
(libfpmhost.so).  Plumbing for tests/ and bench.py only -- the product's host side is the C++
`fpmMain` / `runFPM` in fpm-opencv_b200/host; nothing here computes.

Fails loudly (ImportError / RuntimeError) when the CUDA library is missing: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
INCLUDE_DIR = os.path.abspath(os.path.join(_HERE, "..", "include"))

_lib = None


def lib_path() -> str:
    return os.path.join(LIB_DIR, "libfpmb200.so")


def load():
    """dlopen libfpmb200.so (built in-tree by `make -C fpm-opencv_b200` / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise ImportError("%s not built: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                          "there is no CPU fallback" % p)
    L = C.CDLL(p)
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    sig = {
        "fpmb200_last_error": (C.c_char_p, []),
        "fpmb200_abi_version": (i, []),
        "fpmb200_create": (i, [i, C.POINTER(vp)]),
        "fpmb200_destroy": (None, [vp]),
        "fpmb200_tiles_alloc": (i, [vp, i, i, i, i]),
        "fpmb200_set_params": (i, [vp, f, f, f, i]),
        "fpmb200_set_cluster": (i, [vp, i]),
        "fpmb200_upload_leds": (i, [vp, vp, vp, i]),
        "fpmb200_upload_pupil_support": (i, [vp, vp]),
        "fpmb200_upload_stack": (i, [vp, i, i, vp, vp]),
        "fpmb200_init_tiles": (i, [vp, i, i, i, vp]),
        "fpmb200_run": (i, [vp, i, i, i, vp]),
        "fpmb200_step": (i, [vp, i, i]),
        "fpmb200_finalize": (i, [vp, i, i, vp]),
        "fpmb200_upload_state": (i, [vp, i, vp, vp]),
        "fpmb200_download": (i, [vp, i, vp, vp, vp]),
        "fpmb200_download_objcrop": (i, [vp, i, i, vp, vp]),
        "fpmb200_device_buffer": (i, [vp, i, i, C.POINTER(vp), C.POINTER(C.c_ulonglong)]),
        "fpmb200_set_tile_origins": (i, [vp, vp, vp, i]),
        "fpmb200_ingest_frame": (i, [vp, i, vp, i, i, i, i, i, i, i, i, vp]),
        "fpmb200_ingest_rows": (i, [vp, i, vp, i, i, i, i, i, vp]),
        "fpmb200_ingest_bg": (i, [vp, vp]),
        "fpmb200_event_record": (i, [vp, i, vp]),
        "fpmb200_event_sync": (i, [vp, i]),
        "fpmb200_mosaic": (i, [vp, vp, i, i, i, vp, i, vp]),
        "fpmb200_device_alloc": (i, [vp, C.c_ulonglong, C.POINTER(vp)]),
        "fpmb200_device_free": (i, [vp, vp]),
        "fpmb200_copy_objcrop_to": (i, [vp, i, i, vp, vp, vp]),
        "fpmb200_host_alloc": (i, [C.c_ulonglong, i, C.POINTER(vp)]),
        "fpmb200_host_free": (i, [vp]),
        "fpmb200_sync": (i, [vp]),
        "fpmb200_kernel_launches": (C.c_longlong, [vp]),
        "fpmb200_variant": (C.c_char_p, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


EXPORTS = ["fpmb200_last_error", "fpmb200_abi_version", "fpmb200_create", "fpmb200_destroy",
           "fpmb200_tiles_alloc", "fpmb200_set_params", "fpmb200_set_cluster", "fpmb200_upload_leds",
           "fpmb200_upload_pupil_support", "fpmb200_upload_stack", "fpmb200_init_tiles", "fpmb200_run",
           "fpmb200_step", "fpmb200_finalize", "fpmb200_upload_state", "fpmb200_download",
           "fpmb200_download_objcrop", "fpmb200_device_buffer", "fpmb200_set_tile_origins", "fpmb200_ingest_frame",
           "fpmb200_ingest_rows", "fpmb200_event_record", "fpmb200_event_sync",
           "fpmb200_ingest_bg", "fpmb200_mosaic", "fpmb200_device_alloc", "fpmb200_device_free", "fpmb200_copy_objcrop_to",
           "fpmb200_host_alloc", "fpmb200_host_free", "fpmb200_sync", "fpmb200_kernel_launches", "fpmb200_variant"]


class FpmError(RuntimeError):
    pass


class HostBuffer:
    """Page-locked host memory from fpmb200_host_alloc as a numpy array (freed with the object)."""

    def __init__(self, shape, dtype, write_combined=False):
        self.L = load()
        self.array = None
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        if self.L.fpmb200_host_alloc(n, int(write_combined), C.byref(p)) != 0:
            raise FpmError(self.L.fpmb200_last_error().decode())
        self.ptr = p.value
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_ubyte)), shape=(n,)).view(dtype).reshape(shape)

    def close(self):
        if self.array is not None:
            self.array = None
            self.L.fpmb200_host_free(C.c_void_p(self.ptr))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """Thin object wrapper: one per CUDA device, mirrors the call order runFPM() needs."""

    def __init__(self, device: int = 0):
        self.L = load()
        h = C.c_void_p()
        self._h = None
        self._ck(self.L.fpmb200_create(device, C.byref(h)))
        self._h = h
        self.n_tiles = self.Np = self.Nlarge = self.n_leds = 0

    def _ck(self, rc):
        if rc != 0:
            raise FpmError("fpmb200 error %d: %s" % (rc, self.L.fpmb200_last_error().decode()))

    def close(self):
        if self._h is not None:
            self.L.fpmb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tiles_alloc(self, n_tiles, Np, Nlarge, n_leds):
        self._ck(self.L.fpmb200_tiles_alloc(self._h, n_tiles, Np, Nlarge, n_leds))
        self.n_tiles, self.Np, self.Nlarge, self.n_leds = n_tiles, Np, Nlarge, n_leds

    def set_params(self, delta1, delta2, eps, literal_scalar=1):
        self._ck(self.L.fpmb200_set_params(self._h, float(delta1), float(delta2), float(eps), int(literal_scalar)))

    def set_cluster(self, ctas_per_tile):
        self._ck(self.L.fpmb200_set_cluster(self._h, int(ctas_per_tile)))

    def upload_leds(self, cropX, cropY):
        cx = np.ascontiguousarray(cropX, dtype=np.int16)
        cy = np.ascontiguousarray(cropY, dtype=np.int16)
        self._ck(self.L.fpmb200_upload_leds(self._h, _ptr(cx), _ptr(cy), len(cx)))

    def upload_pupil_support(self, mask):
        m = np.ascontiguousarray(mask, dtype=np.float32)
        assert m.shape == (self.Np, self.Np)
        self._ck(self.L.fpmb200_upload_pupil_support(self._h, _ptr(m)))

    def upload_stack(self, tile_first, stack, stream=None):
        s = np.ascontiguousarray(stack, dtype=np.uint16)
        n = s.size // (self.n_leds * self.Np * self.Np)
        assert n * self.n_leds * self.Np * self.Np == s.size
        self._ck(self.L.fpmb200_upload_stack(self._h, tile_first, n, _ptr(s), stream))
        if stream is None:
            self.sync()

    def upload_stack_ptr(self, tile_first, n, host_ptr, stream=None):
        self._ck(self.L.fpmb200_upload_stack(self._h, tile_first, n, C.c_void_p(host_ptr), stream))

    def set_tile_origins(self, xs, ys):
        x = np.ascontiguousarray(xs, dtype=np.int32)
        y = np.ascontiguousarray(ys, dtype=np.int32)
        self._ck(self.L.fpmb200_set_tile_origins(self._h, _ptr(x), _ptr(y), len(x)))

    def ingest_frame(self, led_slot, frame, divisor, bk1, bk2, bg_threshold, stream=None):
        f = np.ascontiguousarray(frame, dtype=np.uint16)
        self._ck(self.L.fpmb200_ingest_frame(self._h, int(led_slot), _ptr(f), f.shape[1], f.shape[0], int(divisor),
                                             int(bk1[0]), int(bk1[1]), int(bk2[0]), int(bk2[1]), int(bg_threshold), stream))
        if stream is None:
            self.sync()          # `f` may be a temporary

    def ingest_rows(self, led_slot, frame, row0, n_rows, divisor, bg_val, stream=None):
        """rows [row0, row0+n_rows) of `frame` only (a context that owns part of the tile grid), host-computed bg_val"""
        f = np.ascontiguousarray(frame[row0:row0 + n_rows], dtype=np.uint16)
        self._ck(self.L.fpmb200_ingest_rows(self._h, int(led_slot), _ptr(f), f.shape[1], int(row0), int(n_rows), int(divisor),
                                            int(bg_val), stream))
        if stream is None:
            self.sync()

    def event_record(self, slot, stream=None):
        self._ck(self.L.fpmb200_event_record(self._h, int(slot), stream))

    def event_sync(self, slot):
        self._ck(self.L.fpmb200_event_sync(self._h, int(slot)))

    def ingest_bg(self):
        out = np.zeros(self.n_leds, np.int32)
        self._ck(self.L.fpmb200_ingest_bg(self._h, _ptr(out)))
        return out

    def raw_stack(self, tile):
        """uint16 stack of one tile as uploaded / ingested, [n_leds][Np][Np] (device -> host via torch-free memcpy)."""
        import ctypes
        ptr, nbytes = self.device_buffer(4, tile)
        out = np.empty((self.n_leds, self.Np, self.Np), np.uint16)
        rt = ctypes.CDLL("libcudart.so.12")
        rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        self.sync()
        assert rt.cudaMemcpy(_ptr(out), ctypes.c_void_p(ptr), nbytes, 2) == 0
        return out

    def mosaic(self, nx, ny, step):
        f = self.Nlarge // self.Np
        out = np.empty((((ny - 1) * step + self.Np) * f, ((nx - 1) * step + self.Np) * f), np.float32)
        self._ck(self.L.fpmb200_mosaic(self._h, None, nx, ny, step, _ptr(out), 0, None))
        self.sync()
        return out

    def init_tiles(self, tile_first=0, n=None, init_led_slot=1, stream=None):
        self._ck(self.L.fpmb200_init_tiles(self._h, tile_first, self.n_tiles - tile_first if n is None else n,
                                           init_led_slot, stream))

    def run(self, iters, tile_first=0, n=None, stream=None):
        self._ck(self.L.fpmb200_run(self._h, tile_first, self.n_tiles - tile_first if n is None else n, iters, stream))

    def step(self, tile, led_slot):
        self._ck(self.L.fpmb200_step(self._h, tile, led_slot))

    def finalize(self, tile_first=0, n=None, stream=None):
        self._ck(self.L.fpmb200_finalize(self._h, tile_first, self.n_tiles - tile_first if n is None else n, stream))

    def upload_state(self, tile, objF=None, pupil=None):
        o = None if objF is None else np.ascontiguousarray(objF, dtype=np.complex64)
        p = None if pupil is None else np.ascontiguousarray(pupil, dtype=np.complex64)
        self._ck(self.L.fpmb200_upload_state(self._h, tile, _ptr(o), _ptr(p)))

    def download(self, tile, objF=True, objCrop=True, pupil=True):
        o = np.empty((self.Nlarge, self.Nlarge), np.complex64) if objF else None
        c = np.empty((self.Nlarge, self.Nlarge), np.complex64) if objCrop else None
        p = np.empty((self.Np, self.Np), np.complex64) if pupil else None
        self._ck(self.L.fpmb200_download(self._h, tile, _ptr(o), _ptr(c), _ptr(p)))
        return o, c, p

    def download_objcrop_ptr(self, tile_first, n, host_ptr, stream=None):
        self._ck(self.L.fpmb200_download_objcrop(self._h, tile_first, n, C.c_void_p(host_ptr), stream))

    def device_buffer(self, which, tile=0):
        """(device pointer, bytes per tile) of objFc(0) / objCrop(1) / pupil(2) / stack(3)."""
        p, b = C.c_void_p(), C.c_ulonglong()
        self._ck(self.L.fpmb200_device_buffer(self._h, which, tile, C.byref(p), C.byref(b)))
        return p.value, b.value

    def objcrop_tensor(self, tile_first, n):
        """torch view (no copy) of objCrop of n tiles: float32 [n, Nlarge*Nlarge*2] on this device."""
        import torch
        ptr, b = self.device_buffer(1, tile_first)

        class _V:
            __cuda_array_interface__ = {"shape": (n, b // 4), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_V(), device="cuda")

    def sync(self):
        self._ck(self.L.fpmb200_sync(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(self.L.fpmb200_kernel_launches(self._h))

    @property
    def variant(self) -> str:
        return self.L.fpmb200_variant(self._h).decode()
