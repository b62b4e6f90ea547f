"""Host-side file formats of the application layer, CPU only: the dependency-free PNG reader (own
inflate) against PNGs written here with zlib at several compression levels and every scanline filter, cv::imread's flag
semantics, the boost::program_options-compatible .ini reader on the reference's option set, and the VTK writers."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tool(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("io") / "io_tool")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "include", "compat"),
                           os.path.join(ROOT, "tests", "cpp", "io_tool.cpp"), "-o", exe])
    return exe


def paeth(a, b, c):
    p = a.astype(np.int32) + b - c
    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
    return np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c)).astype(np.uint8)


def write_png(path, arr, bit_depth, colour, palette=None, level=6, filters=(0, 1, 2, 3, 4), idat_split=None):
    """arr: [H, W, C] of uint8 / uint16 (or packed rows for bit depths < 8 given as [H, stride] uint8)"""
    h = arr.shape[0]
    if bit_depth == 16:
        rows = arr.astype(">u2").reshape(h, -1).view(np.uint8)
        bpp = 2 * arr.shape[2]
        width = arr.shape[1]
    elif bit_depth == 8:
        rows = arr.reshape(h, -1).astype(np.uint8)
        bpp = arr.shape[2]
        width = arr.shape[1]
    else:
        rows, bpp, width = arr, 1, arr.shape[1] * 8 // bit_depth
    raw = bytearray()
    prev = np.zeros(rows.shape[1], dtype=np.uint8)
    for y in range(h):
        cur = rows[y]
        ft = filters[y % len(filters)]
        left = np.concatenate([np.zeros(bpp, np.uint8), cur[:-bpp]])
        upleft = np.concatenate([np.zeros(bpp, np.uint8), prev[:-bpp]])
        if ft == 0: out = cur
        elif ft == 1: out = cur - left
        elif ft == 2: out = cur - prev
        elif ft == 3: out = cur - ((left.astype(np.int32) + prev) >> 1).astype(np.uint8)
        else: out = cur - paeth(left, prev, upleft)
        raw.append(ft)
        raw += out.astype(np.uint8).tobytes()
        prev = cur
    z = zlib.compress(bytes(raw), level)

    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body) & 0xffffffff)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", width, h, bit_depth, colour, 0, 0, 0))
    if palette is not None:
        out += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    out += chunk(b"tEXt", b"Comment\0synthetic")                # ancillary chunks are skipped
    parts = [z] if not idat_split else [z[i:i + idat_split] for i in range(0, len(z), idat_split)]
    for p in parts:
        out += chunk(b"IDAT", p)
    out += chunk(b"IEND", b"")
    open(path, "wb").write(out)


def imread(tool, path, flags, tmp):
    raw = os.path.join(tmp, "out.raw")
    r = subprocess.run([tool, "imread", path, str(flags), raw], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    if r.stdout.startswith("EMPTY"):
        return None
    data = open(raw, "rb").read()
    head, body = data.split(b"\n", 1)
    rows, cols, typ = [int(x) for x in head.split()]
    depth, ch = typ & 7, (typ >> 3) + 1
    dt = {0: np.uint8, 2: np.uint16}[depth]
    return np.frombuffer(body, dtype=dt).reshape(rows, cols, ch)


ANYDEPTH, COLOR, GRAY, UNCHANGED = 2, 1, 0, -1


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_depth_png_16bit_round_trip(tool, tmp_path, level):
    """the depth maps of the sequences: 16-bit grey, millimetres -- every zlib level (stored / fixed / dynamic Huffman blocks)"""
    rng = np.random.RandomState(level)
    h, w = 97, 131
    yy, xx = np.mgrid[0:h, 0:w]
    depth = (600 + 40 * np.sin(xx / 9.0) + 25 * np.cos(yy / 7.0)).astype(np.uint16)
    depth[rng.rand(h, w) < 0.1] = 0                               # holes
    depth[10:20, 10:40] = rng.randint(0, 65536, size=(10, 30))    # incompressible patch incl. values >= 2^15
    p = str(tmp_path / "d.png")
    write_png(p, depth[..., None], 16, 0, level=level, idat_split=4000)
    got = imread(tool, p, ANYDEPTH, str(tmp_path))
    assert got.dtype == np.uint16 and got.shape == (h, w, 1) and np.array_equal(got[..., 0], depth)
    got8 = imread(tool, p, GRAY, str(tmp_path))                   # without ANYDEPTH: reduced to 8 bit
    assert got8.dtype == np.uint8 and np.array_equal(got8[..., 0], (depth >> 8).astype(np.uint8))


def test_colour_mask_palette_and_packed_pngs(tool, tmp_path):
    rng = np.random.RandomState(3)
    h, w = 33, 41
    rgb = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
    rgb[:, :20] = (rgb[:, :1] // 2)                                # compressible part: exercises long matches
    p = str(tmp_path / "c.png")
    write_png(p, rgb, 8, 2, level=9)
    bgr = imread(tool, p, COLOR, str(tmp_path))
    assert bgr.shape == (h, w, 3) and np.array_equal(bgr, rgb[..., ::-1])
    grey = imread(tool, p, GRAY, str(tmp_path))[..., 0]
    want = ((rgb[..., 0].astype(np.uint32) * 4899 + rgb[..., 1].astype(np.uint32) * 9617 + rgb[..., 2].astype(np.uint32) * 1868 + 8192) >> 14).astype(np.uint8)
    assert np.array_equal(grey, want)
    rgba = np.concatenate([rgb, rng.randint(0, 256, size=(h, w, 1)).astype(np.uint8)], -1)
    write_png(p, rgba, 8, 6, level=6)
    assert np.array_equal(imread(tool, p, COLOR, str(tmp_path)), rgb[..., ::-1])          # alpha dropped
    assert np.array_equal(imread(tool, p, UNCHANGED, str(tmp_path)), rgba[..., [2, 1, 0, 3]])
    # 8-bit mask (demo.cpp:303: imread(mask, CV_8U))
    mask = ((np.add.outer(np.arange(h), np.arange(w)) % 5) == 0).astype(np.uint8) * 255
    write_png(p, mask[..., None], 8, 0)
    assert np.array_equal(imread(tool, p, GRAY, str(tmp_path))[..., 0], mask)
    # palette image, 8-bit indices
    pal = rng.randint(0, 256, size=(7, 3)).astype(np.uint8)
    idx = rng.randint(0, 7, size=(h, w, 1)).astype(np.uint8)
    write_png(p, idx, 8, 3, palette=pal)
    assert np.array_equal(imread(tool, p, COLOR, str(tmp_path)), pal[idx[..., 0]][..., ::-1])
    # 1-bit grey, packed rows (w = 16 pixels -> 2 bytes per row)
    bits = rng.randint(0, 2, size=(h, 16)).astype(np.uint8)
    write_png(p, np.packbits(bits, axis=1), 1, 0, filters=(0,))
    assert np.array_equal(imread(tool, p, GRAY, str(tmp_path))[..., 0], bits * 255)


def test_unreadable_files_give_an_empty_mat(tool, tmp_path):
    assert imread(tool, str(tmp_path / "missing.png"), ANYDEPTH, str(tmp_path)) is None
    p = str(tmp_path / "bad.png")
    write_png(p, np.zeros((4, 4, 1), np.uint16), 16, 0)
    b = bytearray(open(p, "rb").read())
    b[60] ^= 0xff                                                   # corrupt the image data: CRC mismatch
    open(p, "wb").write(bytes(b))
    assert imread(tool, p, ANYDEPTH, str(tmp_path)) is None
    open(p, "wb").write(b"not a png")
    assert imread(tool, p, ANYDEPTH, str(tmp_path)) is None


def test_imwrite_and_masked_copy(tool, tmp_path):
    p = str(tmp_path / "w.png")
    subprocess.check_call([tool, "imwrite16", p, "37", "23"])
    data = open(p, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, ihdr = 8, b"", None
    while pos < len(data):
        n, t = struct.unpack(">I", data[pos:pos + 4])[0], data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(t + body) & 0xffffffff == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0]
        if t == b"IHDR": ihdr = struct.unpack(">IIBBBBB", body)
        if t == b"IDAT": idat += body
        pos += 12 + n
    assert ihdr == (37, 23, 16, 0, 0, 0, 0)
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(23, 1 + 37 * 2)      # python's zlib accepts our stream
    assert (raw[:, 0] == 0).all()
    got = raw[:, 1:].copy().view(">u2").astype(np.uint16)
    yy, xx = np.mgrid[0:23, 0:37]
    want = np.where((xx + yy) % 3 == 0, (yy * 257 + xx * 3) & 0xffff, 0).astype(np.uint16)
    assert np.array_equal(got, want)
    assert np.array_equal(imread(tool, p, ANYDEPTH, str(tmp_path))[..., 0], want)             # and our own reader


INI = """# TSDF
VOL_DIMS_X=96
VOL_DIMS_Y=80
VOL_DIMS_Z = 64

VOL_SIZE_X=0.9
VOL_SIZE_Y=0.75   # trailing comment
VOL_SIZE_Z=0.6
TSDF_TRUNC_DIST=10
ETA=5
TSDF_MAX_WEIGHT=128
GRADIENT_DELTA_FACTOR=0.5
INTR_FX=517.0
INTR_FY=516.5
INTR_CX=320.0
INTR_CY=240.0
TRUNC_DEPTH=3.0
VOL_POSE_T_Z=0.05
BILATERAL_SIGMA_DEPTH=0.01
BILATERAL_SIGMA_SPATIAL=4.5
BILATERAL_KERNEL_SIZE=7
START_FRAME=4
MAX_ITER=2048
MAX_UPDATE_NORM=1e-3
S=7
LAMBDA=0.1
ALPHA=0.1
W_REG=0.2
ALPHA=0.9
"""


def run_ini(tool, text, tmp_path):
    p = tmp_path / "p.ini"
    p.write_text(text)
    r = subprocess.run([tool, "ini", str(p)], capture_output=True, text=True)
    return r.returncode, dict(l.split("=", 1) for l in r.stdout.splitlines() if "=" in l), r.stdout


def test_ini_reader_follows_program_options(tool, tmp_path):
    rc, _, out = run_ini(tool, INI, tmp_path)                      # ALPHA is given twice: boost throws multiple_occurrences
    assert rc == 3 and "ALPHA" in out and "more than once" in out
    rc, kv, _ = run_ini(tool, INI.replace("ALPHA=0.9\n", ""), tmp_path)
    assert rc == 0
    assert kv["VOL_DIMS"] == "96 80 64" and kv["VOL_SIZE"].split() == ["0.899999976", "0.75", "0.600000024"]
    assert kv["TSDF_TRUNC_DIST"] == "10" and kv["ETA"] == "5" and kv["VOL_POSE_T_Z"] == "0.0500000007"
    assert kv["INTR"].split() == ["517", "516.5", "320", "240"] and kv["BILATERAL"].split()[2] == "7"
    assert kv["MAX_ITER"] == "2048" and kv["MAX_UPDATE_NORM"] == "0.00100000005" and kv["S"] == "7"
    assert kv["ALPHA"] == "0.100000001"
    rc, _, out = run_ini(tool, INI.replace("ALPHA=0.9\n", "") + "RHO_0=1.0\n", tmp_path)      # an option the application does not declare is an error
    assert rc == 3 and "RHO_0" in out
    rc, _, out = run_ini(tool, INI.replace("ALPHA=0.9\n", "").replace("MAX_ITER=2048", "MAX_ITER=20.5"), tmp_path)
    assert rc == 3 and "MAX_ITER" in out                           # lexical_cast<int> rejects it
    rc, _, out = run_ini(tool, "VOL_DIMS_X 96\n", tmp_path)
    assert rc == 3


@pytest.mark.skipif(not os.path.isdir("/root/reference/params"), reason="reference checkout not present")
def test_every_reference_ini_is_read(tool):
    """build container only: the reference's own parameter files.  params_boxing.ini carries RHO_0, which demo.cpp does not
    declare -- boost::program_options rejects that file, and so do we ('RHO_0 line dropped')."""
    seen = 0
    for name in sorted(os.listdir("/root/reference/params")):
        r = subprocess.run([tool, "ini", os.path.join("/root/reference/params", name)], capture_output=True, text=True)
        text = open(os.path.join("/root/reference/params", name)).read()
        if "RHO_0" in text:
            assert r.returncode == 3 and "RHO_0" in r.stdout, name
        else:                                   # params_ours.ini holds comments only: it parses to nothing
            assert r.returncode == 0 and ("S=7" in r.stdout) == ("\nS=7" in text), (name, r.stdout)
        seen += 1
    assert seen >= 6


def test_vtk_writers(tool, tmp_path):
    p = str(tmp_path / "m.vtk")
    subprocess.check_call([tool, "vtk", p])
    lines = open(p).read().split("\n")
    assert lines[:5] == ["# vtk DataFile Version 3.0", "vtk output", "ASCII", "DATASET POLYDATA", "POINTS 6 float"]
    pts = np.array([[float(v) for v in l.split()] for l in lines[5:11]])
    assert np.allclose(pts, [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0.5, 0.25, -1.5], [1e-3, 123456.789, 2], [3, 2, 1]], rtol=1e-4)
    i = lines.index("VERTICES 6 12")
    assert lines[i + 1:i + 7] == ["1 %d" % k for k in range(6)]
    j = lines.index("POLYGONS 2 8")
    assert lines[j + 1:j + 3] == ["3 0 1 2", "3 3 4 5"]
    p = str(tmp_path / "f.vti")
    subprocess.check_call([tool, "vti", p])
    blob = open(p, "rb").read()
    assert b'WholeExtent="0 2 0 1 0 1"' in blob and b'NumberOfComponents="4"' in blob and b'header_type="UInt64"' in blob
    k = blob.index(b"_", blob.index(b"<AppendedData")) + 1
    n = struct.unpack("<Q", blob[k:k + 8])[0]
    assert n == 3 * 2 * 2 * 4 * 4
    assert np.array_equal(np.frombuffer(blob[k + 8:k + 8 + n], dtype="<f4"), 0.5 * np.arange(48, dtype=np.float32))


def test_glob_lists_sorted_regular_files(tool, tmp_path):
    d = tmp_path / "depth"
    d.mkdir()
    (d / "sub").mkdir()
    for n in ("depth_000010.png", "depth_000002.png", "depth_000001.png"):
        (d / n).write_bytes(b"x")
    r = subprocess.run([tool, "glob", str(d)], capture_output=True, text=True)
    assert [os.path.basename(l) for l in r.stdout.splitlines()] == ["depth_000001.png", "depth_000002.png", "depth_000010.png"]


def test_png_reader_survives_corrupt_files(tmp_path):
    """memory safety of the PNG / inflate code on hostile input: the reader (built with AddressSanitizer + UBSan) either decodes or
    reports an empty image, for truncated files and for image data corrupted UNDER a valid chunk CRC (so the inflate code sees it)"""
    exe = str(tmp_path / "io_tool_asan")
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
                           "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "include", "compat"),
                           os.path.join(ROOT, "tests", "cpp", "io_tool.cpp"), "-o", exe])
    rng = np.random.RandomState(7)
    depth = (500 + 30 * np.sin(np.arange(64)[None, :] / 5.0) + np.arange(48)[:, None]).astype(np.uint16)
    good = str(tmp_path / "good.png")
    write_png(good, depth[..., None], 16, 0, level=6)
    blob = open(good, "rb").read()
    i0 = blob.index(b"IDAT")
    n = struct.unpack(">I", blob[i0 - 4:i0])[0]
    body = bytearray(blob[i0 + 4:i0 + 4 + n])

    def rebuild(new_body, ihdr=None):
        out = bytearray(blob[:i0 - 4]) if ihdr is None else bytearray(b"\x89PNG\r\n\x1a\n" + struct.pack(">I", 13) + b"IHDR" + ihdr +
                                                                     struct.pack(">I", zlib.crc32(b"IHDR" + ihdr) & 0xffffffff))
        out += struct.pack(">I", len(new_body)) + b"IDAT" + bytes(new_body) + struct.pack(">I", zlib.crc32(b"IDAT" + bytes(new_body)) & 0xffffffff)
        out += struct.pack(">I", 0) + b"IEND" + struct.pack(">I", zlib.crc32(b"IEND") & 0xffffffff)
        return bytes(out)

    cases = [blob[:k] for k in (0, 7, 8, 20, 33, 40, len(blob) // 2, len(blob) - 5)]
    for _ in range(150):                                   # bit flips / byte edits inside the deflate stream, CRC kept valid
        b = bytearray(body)
        for _ in range(rng.randint(1, 4)):
            b[rng.randint(0, len(b))] = rng.randint(0, 256)
        cases.append(rebuild(b))
    cases.append(rebuild(body[:len(body) // 2]))           # truncated stream
    cases.append(rebuild(body, struct.pack(">IIBBBBB", 60000, 60000, 16, 0, 0, 0, 0)))      # header claims a huge image
    cases.append(rebuild(body, struct.pack(">IIBBBBB", 64, 48, 16, 3, 0, 0, 0)))            # palette type with 16 bits
    cases.append(rebuild(zlib.compress(b"\0" * (1 << 22), 9)))                              # far more data than the header announces
    bad = str(tmp_path / "bad.png")
    decoded = 0
    for c in cases:
        open(bad, "wb").write(c)
        r = subprocess.run([exe, "imread", bad, str(ANYDEPTH), str(tmp_path / "o.raw")], capture_output=True, text=True)
        assert r.returncode == 0 and (r.stdout.startswith("EMPTY") or r.stdout.startswith("OK")), (r.returncode, r.stdout[-300:], r.stderr[-1500:])
        decoded += r.stdout.startswith("OK")
    assert decoded < len(cases) // 2                        # most corruptions are detected (the rest only change pixel values)
    assert imread(exe, good, ANYDEPTH, str(tmp_path)) is not None
