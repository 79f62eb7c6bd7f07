"""CPU tests of the product's host layer (C++: own JSON reader, geometry, LED order, TIFF loader,
preprocessing) through its C ABI, against the golden vectors produced by the reference's own
jsoncpp + libstdc++ (tests/golden/make_golden.py), plus ABI/symbol checks of both libraries."""
import ctypes
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import fpm_oracle as orc
import fpm_testlib as T
import fpmb200
import fpmhost


def bits(f):
    return struct.unpack("I", struct.pack("f", float(f)))[0]


def _check_against_golden(ds, g):
    n = ds.geometry(g["present_first"], g["present_last"])
    s = ds.scalars
    assert n == g["ledUsedCount"] == s.ledUsedCount
    assert (s.Np, s.resImprovementFactor, s.Nlarge, s.Mlarge, s.naRadius) == (g["Np"], g["factor"], g["Nlarge"], g["Mlarge"], g["naRadius"])
    assert (bits(s.ps_eff), bits(s.du), bits(s.lambda_), bits(s.objectiveNA), bits(s.maxIlluminationNA)) == (
        g["ps_eff_bits"], g["du_bits"], g["lambda_bits"], g["objectiveNA_bits"], g["maxIlluminationNA_bits"])
    assert (s.delta1, s.delta2, s.bgThreshold, s.ledCount, s.darkfieldExpMultiplier) == (
        g["delta1"], g["delta2"], g["bgThreshold"], g["ledCount"], g["darkfieldExpMultiplier"])
    assert s.arrayRotation == g["arrayRotation"]
    assert (bool(s.flipIlluminationX), bool(s.flipIlluminationY)) == (g["flipX"], g["flipY"])
    for l in g["leds"]:
        L = ds.led(l["n"])
        assert L.used == 1 and bits(L.illumination_na) == l["na_bits"]
        assert (L.idx_u, L.idx_v, L.cropXStart, L.cropYStart) == (l["idx_u"], l["idx_v"], l["cropX"], l["cropY"])
        assert (L.cropXEnd, L.cropYEnd) == (l["cropX"] + g["Np"] - 1, l["cropY"] + g["Np"] - 1)
    used = {l["n"] for l in g["leds"]}
    for n_ in range(g["present_first"], g["present_last"] + 1):
        if n_ not in used:
            assert ds.led(n_).used == 0
    assert list(ds.order) == g["order"]


@pytest.mark.parametrize("name", T.ALL_CFGS + T.QUIRKS)
def test_host_geometry_bit_exact_embedded(name):
    """`holeCoordinates` embedded -- the only form HEAD of the reference reads (fpmMain.cpp:77-79)."""
    ds = fpmhost.Dataset(T.embedded_path(name))
    assert ds.geometry_source == "holeCoordinates"
    _check_against_golden(ds, T.golden(name))


@pytest.mark.parametrize("name,source", [("cfg1_mono_np64", "ledList:"), ("cfg2_fLEDc_np128", "ledList:"),
                                         ("cfg5_cellscope2_np128", "holePositions"),
                                         ("cfg5b_cellscope2_np256", "holePositions"),
                                         ("cfg6_mono_dome_np64", "domeHoleCoordinates")])
def test_host_geometry_other_led_formats(name, source):
    """ledArrayMaps `ledList`, cellscope2 `holePositions` and the built-in dome table give the same
    bits as the embedded form."""
    ds = fpmhost.Dataset(os.path.join(T.CONFIGS, name + ".json"))
    assert ds.geometry_source.startswith(source)
    _check_against_golden(ds, T.golden(name))


def test_defaults_when_keys_missing(tmp_path):
    p = tmp_path / "empty.json"
    p.write_text("{}")
    s = fpmhost.Dataset(str(p)).scalars          # fpmMain.cpp:517-575 defaults
    assert (s.Np, s.ledCount, s.cropX, s.cropY, s.darkfieldExpMultiplier) == (90, 508, 1, 1, 1)
    assert (s.delta1, s.delta2, s.bgThreshold) == (5.0, 10.0, 1000.0)
    assert bits(s.lambda_) == bits(np.float32(0.5)) and bits(s.maxIlluminationNA) == bits(np.float32(0.7604))
    assert bits(s.eps) == bits(np.float32(0.0000000001))
    # bgThreh (sic, dataset_mono.json:20) is not the key the reference reads -> default
    p.write_text('{"bgThreh": 20, "delta1": 7.9, "arrayRotation": -7.9}')
    s = fpmhost.Dataset(str(p)).scalars
    assert s.bgThreshold == 1000.0 and s.delta1 == 7.0 and s.arrayRotation == -7.0   # asInt() truncation


def test_json_recovery_drops_keys_after_trailing_comma_array(tmp_path):
    """Reader::recoverFromError is token-level: after `,]` everything up to the next `]` is skipped."""
    p = tmp_path / "t.json"
    p.write_text('{"cropSizeX": 64, "holeCoordinates": [[{"x":1},{"y":2},{"z":50}],], "delta1": 3, "q": [1], "delta2": 4}')
    ds = fpmhost.Dataset(str(p))
    s = ds.scalars
    assert s.Np == 64 and s.delta1 == 5.0 and s.delta2 == 10.0     # both lost -> defaults
    assert ds.geometry(1, 3) == 3                                  # rows 2,3 are null -> (0,0,0) -> NA 0 passes
    assert ds.led(2).illumination_na == 0.0 and ds.led(1).illumination_na > 0


def test_missing_file_is_an_error():
    with pytest.raises(RuntimeError):
        fpmhost.Dataset("/nonexistent/dataset.json")


def test_pupil_support_matches_oracle():
    for N, r in [(64, 21), (128, 17), (128, 42), (256, 94)]:
        assert np.array_equal(fpmhost.pupil_support(N, r), orc.pupil_support(N, r).astype(np.float32))


def test_device_from_env(monkeypatch):
    L = fpmhost.load()
    monkeypatch.delenv("OPENCV_OPENCL_DEVICE", raising=False)
    assert L.fpmhost_device_from_env() == 0
    monkeypatch.setenv("OPENCV_OPENCL_DEVICE", "GPU:3")
    assert L.fpmhost_device_from_env() == 3
    monkeypatch.setenv("OPENCV_OPENCL_DEVICE", "CPU:0")       # use_cpu.sh
    assert L.fpmhost_device_from_env() == -1
    for f, want in (("use_gpu.sh", "GPU:0"), ("use_cpu.sh", "CPU:0")):
        assert open(os.path.join(T.ROOT, f)).read().strip() == "export OPENCV_OPENCL_DEVICE=" + want


# ---- image loader ---------------------------------------------------------------------------
def write_tiff16(path, img, big_endian=False):
    """uncompressed single-strip 16-bit TIFF (what the reference's datasets are); [h][w] grey or [h][w][3] chunky RGB."""
    h, w = img.shape[:2]
    spp = 1 if img.ndim == 2 else img.shape[2]
    e = ">" if big_endian else "<"
    data = img.astype(e + "u2").tobytes()
    ifd_off = 8 + len(data)
    n_tags = 9
    extra_off = ifd_off + 2 + 12 * n_tags + 4            # out-of-line BitsPerSample for RGB (3 shorts do not fit the entry)
    tags = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, spp, 16 if spp == 1 else extra_off), (259, 3, 1, 1),
            (262, 3, 1, 1 if spp == 1 else 2), (273, 4, 1, 8), (277, 3, 1, spp), (278, 4, 1, h), (279, 4, 1, len(data))]
    ifd = struct.pack(e + "H", len(tags))
    for tag, typ, cnt, val in tags:
        inline_short = typ == 3 and cnt == 1
        ifd += struct.pack(e + "HHI", tag, typ, cnt) + (struct.pack(e + "HH", val, 0) if inline_short else struct.pack(e + "I", val))
    ifd += struct.pack(e + "I", 0)
    if spp > 1:
        ifd += struct.pack(e + "%dH" % spp, *([16] * spp))
    with open(path, "wb") as f:
        f.write((b"MM" if big_endian else b"II") + struct.pack(e + "HI", 42, ifd_off) + data + ifd)


def test_loader_preprocessing_matches_opencv(tmp_path):
    """ROI crop, dark-field divide, 2-ROI background mean/clamp/subtract (fpmMain.cpp:124-144)
    against the same cv2 calls, files found by the reference's name rule (fpmMain.cpp:69-75)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    root = tmp_path / "frames"
    root.mkdir()
    pts = [(0.0, 0.0, 60.0), (4.0, 0.0, 60.0), (0.0, -4.0, 60.0), (8.0, 8.0, 60.0), (30.0, 0.0, 60.0), (90.0, 0.0, 60.0)]
    cfg = {"datasetRoot": str(root) + "/", "filePrefix": "iLED_", "fileExtension": ".tif", "cropSizeX": 64,
           "cropX": 40, "cropY": 33, "bk1cropX": 3, "bk1cropY": 100, "bk2cropX": 120, "bk2cropY": 7, "bgThresh": 700,
           "pixelSize": 6.5, "objectiveMag": 8, "objectiveNA": 0.1, "maxIlluminationNA": 0.6, "lambda": 0.5,
           "darkfieldExpMultiplier": 3, "ledCount": 10,
           "holeCoordinates": [[{"x": x}, {"y": y}, {"z": z}] for x, y, z in pts]}
    import json
    (tmp_path / "d.json").write_text(json.dumps(cfg))
    frames = {}
    for n in range(1, 7):
        fr = rng.integers(0, 4000 if n != 4 else 300, (200, 230)).astype(np.uint16)
        frames[n] = fr
        write_tiff16(str(root / ("iLED_%04d.tif" % n)), fr, big_endian=(n == 2))
    (root / "notes.txt").write_text("ignored")
    (root / "other_0001.tif").write_bytes(b"junk")            # wrong prefix: must be ignored
    ds = fpmhost.Dataset(str(tmp_path / "d.json"))
    assert ds.load_images() == 1
    s = ds.scalars
    assert s.ledUsedCount == 5                                  # LED 6 (NA 0.83) is skipped
    for n in range(1, 6):
        L = ds.led(n)
        fr = frames[n]
        img = fr[33:33 + 64, 40:40 + 64].copy()
        if L.illumination_na > s.objectiveNA:
            img = cv2.divide(img, 3.0)
        bg = (cv2.mean(fr[7:7 + 64, 120:120 + 64])[0] + cv2.mean(fr[100:100 + 64, 3:3 + 64])[0]) / 2
        bg = min(bg, 700.0)
        bgv = int(np.floor(bg + 0.5))
        want = cv2.subtract(img, (float(bgv), 0, 0, 0))
        assert L.bg_val == bgv
        assert np.array_equal(ds.image(n), want), n
    assert ds.led(1).illumination_na == 0.0 and ds.led(5).illumination_na > s.objectiveNA
    # no images -> -1 like fpmMain.cpp:241-244 ; missing directory -> -1 like :266-270
    cfg["filePrefix"] = "nothing_"
    (tmp_path / "e.json").write_text(json.dumps(cfg))
    assert fpmhost.Dataset(str(tmp_path / "e.json")).load_images() == -1
    cfg["datasetRoot"] = str(tmp_path / "missing") + "/"
    (tmp_path / "f.json").write_text(json.dumps(cfg))
    assert fpmhost.Dataset(str(tmp_path / "f.json")).load_images() == -1


def test_tiff_reader_reads_opencv_written_files(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    img = rng.integers(0, 65535, (70, 66)).astype(np.uint16)
    root = tmp_path / "r"
    root.mkdir()
    cv2.imwrite(str(root / "iLED_0001.tif"), img, [cv2.IMWRITE_TIFF_COMPRESSION, 1])
    import json
    cfg = {"datasetRoot": str(root) + "/", "cropSizeX": 64, "cropX": 1, "cropY": 2, "bk1cropX": 0, "bk1cropY": 0,
           "bk2cropX": 0, "bk2cropY": 0, "bgThresh": 0, "ledCount": 4, "holeCoordinates": [[{"x": 0}, {"y": 0}, {"z": 50}]]}
    (tmp_path / "d.json").write_text(json.dumps(cfg))
    ds = fpmhost.Dataset(str(tmp_path / "d.json"))
    assert ds.load_images() == 1
    assert np.array_equal(ds.image(1), img[2:66, 1:65])
    # compressed TIFFs are rejected loudly, not mis-read
    cv2.imwrite(str(root / "iLED_0001.tif"), img, [cv2.IMWRITE_TIFF_COMPRESSION, 5])
    assert fpmhost.Dataset(str(tmp_path / "d.json")).load_images() == -1


def test_colour_frames_keep_the_channel_the_reference_keeps(tmp_path):
    """isColor (fpmMain.cpp:109-116, dataset_cellScope.json): imread(ANYDEPTH | COLOR) gives BGR and the reference keeps
    channels[2] -- the red plane = sample 0 of a chunky RGB TIFF.  Checked against cv2.imread on files written by this
    test's own writer (little and big endian) and by cv2.imwrite; a grey TIFF under isColor reads as itself (OpenCV
    replicates it into the three channels)."""
    cv2 = pytest.importorskip("cv2")
    import json
    rng = np.random.default_rng(9)
    root = tmp_path / "frames"
    root.mkdir()
    pts = [(0.0, 0.0, 60.0), (4.0, 0.0, 60.0), (0.0, -4.0, 60.0), (6.0, 6.0, 60.0)]
    cfg = {"datasetRoot": str(root) + "/", "cropSizeX": 32, "cropX": 11, "cropY": 6, "bk1cropX": 0, "bk1cropY": 50,
           "bk2cropX": 60, "bk2cropY": 1, "bgThresh": 500, "isColor": True, "objectiveNA": 0.05, "darkfieldExpMultiplier": 2,
           "maxIlluminationNA": 0.6, "ledCount": 8, "holeCoordinates": [[{"x": x}, {"y": y}, {"z": z}] for x, y, z in pts]}
    (tmp_path / "d.json").write_text(json.dumps(cfg))
    rgb = {n: rng.integers(0, 3000, (90, 100, 3)).astype(np.uint16) for n in (1, 2, 3)}
    write_tiff16(str(root / "iLED_0001.tif"), rgb[1])
    write_tiff16(str(root / "iLED_0002.tif"), rgb[2], big_endian=True)
    cv2.imwrite(str(root / "iLED_0003.tif"), rgb[3][..., ::-1].copy(), [cv2.IMWRITE_TIFF_COMPRESSION, 1])    # imwrite takes BGR
    grey = rng.integers(0, 3000, (90, 100)).astype(np.uint16)
    write_tiff16(str(root / "iLED_0004.tif"), grey)
    ds = fpmhost.Dataset(str(tmp_path / "d.json"))
    assert ds.load_images() == 1 and ds.scalars.ledUsedCount == 4
    for n in (1, 2, 3, 4):
        bgr = cv2.imread(str(root / ("iLED_%04d.tif" % n)), cv2.IMREAD_ANYDEPTH | cv2.IMREAD_COLOR)
        assert bgr.dtype == np.uint16 and bgr.shape == (90, 100, 3)
        full = cv2.split(bgr)[2]                                            # fpmMain.cpp:112-115
        assert np.array_equal(full, rgb[n][..., 0] if n < 4 else grey)
        L = ds.led(n)
        want, bg = fpmhost.preprocess_frame(full, 32, (11, 6), (0, 50), (60, 1), 2 if L.illumination_na > ds.scalars.objectiveNA else 1, 500)
        assert L.bg_val == bg and np.array_equal(ds.image(n), want), n
    # several channels without isColor: refused (the reference would reinterpret interleaved samples as grey pixels)
    cfg["isColor"] = False
    (tmp_path / "e.json").write_text(json.dumps(cfg))
    assert fpmhost.Dataset(str(tmp_path / "e.json")).load_images() == -1


# ---- ABI surface ------------------------------------------------------------------------------
def _declared(header):
    txt = open(os.path.join(T.ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fpm(?:b200|host)_[a-z_0-9]+)\s*\(", txt)))


def test_cuda_library_exports_every_declared_symbol():
    names = _declared("fpmb200.h")
    assert sorted(names) == sorted(fpmb200.EXPORTS)
    lib = ctypes.CDLL(fpmb200.lib_path())           # loads without a GPU (no compute calls here)
    for n in names:
        assert hasattr(lib, n), n
    assert fpmb200.load().fpmb200_abi_version() == 1


def test_host_library_exports_every_declared_symbol():
    names = _declared("fpmhost.h")
    assert sorted(names) == sorted(fpmhost.EXPORTS)
    lib = ctypes.CDLL(fpmhost.lib_path())
    for n in names:
        assert hasattr(lib, n), n


def test_cuda_library_is_sm100a_and_uses_bulk_async_copy():
    out = subprocess.run(["cuobjdump", "-lelf", fpmb200.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="CPU-box behaviour")
def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the product refuses to run -- it never routes to the oracle."""
    with pytest.raises(fpmb200.FpmError):
        fpmb200.Context(0)
    exe = os.path.join(T.ROOT, "fpm-opencv_b200", "bin", "fpmMain")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "Not enough input" in r.stdout          # fpmMain.cpp:501-506
    env = dict(os.environ, OPENCV_OPENCL_DEVICE="CPU:0")
    r = subprocess.run([exe, os.path.join(T.CONFIGS, "cfg1_mono_np64.json"), "1"], capture_output=True, text=True, env=env,
                       cwd=T.ROOT)
    assert r.returncode == 3 and "no CPU reconstruction path" in r.stdout


def test_preprocess_frame_and_tile_grid():
    """fpmhost_preprocess_frame == the cv2 statement of fpmMain.cpp:124-144 (cv::divide, cv::mean, saturating
    cv::subtract); fpmhost_tile_grid counts."""
    import cv2
    rng = np.random.default_rng(3)
    H, W, Np = 150, 210, 48
    for divisor, thr in ((1, 1000), (3, 1000), (7, 40), (0, 1000)):
        frame = (rng.integers(0, 4000, (H, W)) + rng.integers(0, 2, (H, W)) * 61000).astype(np.uint16)
        crop, bk1, bk2 = (31, 17), (100, 90), (5, 60)
        img, bg = fpmhost.preprocess_frame(frame, Np, crop, bk1, bk2, divisor, thr)
        roi = frame[crop[1]:crop[1] + Np, crop[0]:crop[0] + Np].copy()
        if divisor != 1:
            roi = cv2.divide(roi, np.full(roi.shape, divisor, np.uint16) if divisor else np.zeros(roi.shape, np.uint16))
        m1 = cv2.mean(frame[bk1[1]:bk1[1] + Np, bk1[0]:bk1[0] + Np])[0]
        m2 = cv2.mean(frame[bk2[1]:bk2[1] + Np, bk2[0]:bk2[0] + Np])[0]
        b = min((m2 + m1) / 2, thr)
        bgv = int(np.floor(b + 0.5))
        ref = cv2.subtract(roi, np.full(roi.shape, bgv, np.uint16))
        assert bg == bgv and np.array_equal(img, ref), (divisor, thr)
    nx, ny, xs, ys = fpmhost.tile_grid(2560, 2160, 128, 0)
    assert (nx, ny) == (20, 16) and xs[21] == 128 and ys[21] == 128
    nx, ny, xs, ys = fpmhost.tile_grid(2560, 2160, 128, 32)
    assert (nx, ny) == (26, 22) and xs[-1] + 128 <= 2560 and ys[-1] + 128 <= 2160
