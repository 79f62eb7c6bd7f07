"""Synthetic intensity stacks -- input-data plumbing for tests and bench.py.

The reference ships no images (`datasetRoot` are absolute /Users/... paths), so stacks are produced
with the reference's own forward model: crop of the centred spectrum of a seeded ground-truth object
at (cropYStart, cropXStart) -> multiply by the disc pupil -> inverse FFT -> |.|^2, scaled so the
global maximum is 60000 and rounded to uint16.  numpy only; nothing here is on the timed path.
"""
from __future__ import annotations

import numpy as np


def _sh(a):
    return np.roll(a, (a.shape[0] // 2, a.shape[1] // 2), axis=(0, 1))


def disc_support(N: int, r: int) -> np.ndarray:
    y, x = np.mgrid[0:N, 0:N]
    return _sh(((x - N // 2) ** 2 + (y - N // 2) ** 2 <= r * r).astype(np.float64))


def synth_object(L: int, seed: int, sigma: float = 0.1) -> np.ndarray:
    """amp = 0.3 + 0.7*G1 (G1 in [0,1]), phase = G2/max|G2| rad; G = Gaussian-low-passed white noise
    (sigma cycles/pixel: 0.1 puts signal under the dark-field LEDs too)."""
    rng = np.random.default_rng(seed)
    fy = np.fft.fftfreq(L)[:, None]
    fx = np.fft.fftfreq(L)[None, :]
    lp = np.exp(-(fx * fx + fy * fy) / (2 * sigma ** 2))

    def smooth():
        w = rng.standard_normal((L, L))
        return np.real(np.fft.ifft2(np.fft.fft2(w) * lp))

    g1 = smooth()
    g1 = (g1 - g1.min()) / (g1.max() - g1.min())
    g2 = smooth()
    g2 = g2 / np.abs(g2).max()
    return (0.3 + 0.7 * g1) * np.exp(1j * g2)


def synth_stack(N: int, L: int, naRadius: int, cropX, cropY, seed: int) -> np.ndarray:
    """uint16 [n_leds][N][N], slot k imaged through the window at (cropY[k], cropX[k])."""
    obj = synth_object(L, seed)
    Fc = np.fft.fftshift(np.fft.fft2(obj))
    S = disc_support(N, naRadius)
    out = np.empty((len(cropX), N, N), dtype=np.float64)
    for k, (xs, ys) in enumerate(zip(cropX, cropY)):
        xs, ys = int(xs), int(ys)
        out[k] = np.abs(np.fft.ifft2(_sh(Fc[ys:ys + N, xs:xs + N]) * S)) ** 2
    out *= 60000.0 / out.max()
    return np.rint(out).astype(np.uint16)


def write_tiff16(path, img):
    """uncompressed single-strip little-endian 16-bit grey TIFF (what the reference's datasets are; libtiff's
    DumpModeDecode in its profile, output.svg:13,133) -- synthetic frames for the full-FOV legs."""
    import struct
    h, w = img.shape
    data = np.ascontiguousarray(img, dtype="<u2").tobytes()
    tags = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, 16), (259, 3, 1, 1), (262, 3, 1, 1), (273, 4, 1, 8),
            (277, 3, 1, 1), (278, 4, 1, h), (279, 4, 1, len(data))]
    ifd = struct.pack("<H", len(tags))
    for tag, typ, cnt, val in tags:
        ifd += struct.pack("<HHI", tag, typ, cnt) + (struct.pack("<HH", val, 0) if typ == 3 else struct.pack("<I", val))
    ifd += struct.pack("<I", 0)
    with open(path, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, 8 + len(data)))
        f.write(data)
        f.write(ifd)
