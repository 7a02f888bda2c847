"""Host-side loaders of the reference's input files -- no GPU needed.

brdfgpu_read_cal / _read_obj / _read_png restate CBRDFdata::LoadCameraParameters + WriteValue
(brdfdata.cpp:149-247), igl::readOBJ as LoadModel uses it (:289-312) and cv::imread(IMREAD_COLOR)
(:40,122).  Checked against independent Python readers (regex / split / cv2 / PIL) on generated
files, and on the reference's own cup / bunny files when that tree is mounted."""
import os
import struct
import zlib

import numpy as np
import pytest

import real_scenes as RS
from brdf_b200 import api as A

REF = "/root/reference"
FIELDS = RS.CAL_FIELDS


def _write_cal(path, vals, extra_head="<camera_model>CameraTsai</camera_model>\n", kappa=True, order=None, eol="\n"):
    names = list(FIELDS) if order is None else order
    lines = [extra_head] if extra_head else []
    for k in names:
        lines.append("<%s>%s</%s>%s" % (k, repr(float(vals[FIELDS.index(k)])), k, eol))
        if kappa and k == "sx":
            lines.append("<kappa1>1.6594307295081538e-008</kappa1>" + eol)
    with open(path, "w", newline="") as f:
        f.write("".join(lines))


def test_read_cal_matches_regex_reader(tmp_path):
    rng = np.random.default_rng(1)
    vals = rng.normal(0, 100, 16)
    p = tmp_path / "a.cal"
    _write_cal(p, vals)
    cam, mask = A.read_cal(p)
    assert mask == 0xFFFF
    assert cam.tobytes() == RS.parse_cal(p).tobytes() == np.asarray(vals).tobytes()
    # field order, CRLF line ends and unknown tags do not matter; a missing field stays 0 and is reported
    order = list(reversed(FIELDS))
    order.remove("pz")
    _write_cal(p, vals, order=order, eol="\r\n", extra_head="")
    cam, mask = A.read_cal(p)
    assert mask == 0xFFFF & ~(1 << 15) and cam[15] == 0.0
    assert cam[:15].tobytes() == np.asarray(vals)[:15].tobytes()


def test_read_cal_atof_semantics(tmp_path):
    """atof(): leading blanks skipped, trailing junk ignored, unparsable -> 0 (brdfdata.cpp:197)."""
    p = tmp_path / "b.cal"
    p.write_text("<cx>  12.5abc</cx>\n<cy>junk</cy>\n<f>1e3</f>\n<sx>-0x10</sx>\n")
    cam, mask = A.read_cal(p)
    assert mask == 0b1111
    assert cam[0] == 12.5 and cam[1] == 0.0 and cam[2] == 1000.0 and cam[3] == -16.0


def test_read_cal_missing_file(tmp_path):
    with pytest.raises(A.BrdfGpuError):
        A.read_cal(tmp_path / "nope.cal")


def test_read_obj_forms(tmp_path):
    p = tmp_path / "m.obj"
    p.write_text("# comment\n"
                 "mtllib x.mtl\n"
                 "v 0 0 0\nv 1 0 0.5\nv 0 1 -2e-3\nv 1 1 1 0.25\n"   # 4th vertex carries a w that is ignored
                 "vt 0.5 0.5\nvn 0 0 1\n"
                 "f 1 2 3\n"
                 "f 2/1 3/1 4/1\n"
                 "f 1//1 3//1 4//1\n"
                 "f 1/1/1 2/1/1 4/1/1\r\n"
                 "f -4 -3 -1\n"
                 "f 1 2 3 4\n")            # a quad: the reference only ever reads the first three columns
    V, F = A.read_obj(p)
    assert V.shape == (4, 3) and F.shape == (6, 3)
    np.testing.assert_array_equal(V, [[0, 0, 0], [1, 0, 0.5], [0, 1, -2e-3], [1, 1, 1]])
    np.testing.assert_array_equal(F, [[0, 1, 2], [1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 3], [0, 1, 2]])


def test_read_obj_matches_split_reader(tmp_path):
    rng = np.random.default_rng(2)
    V = rng.normal(0, 50, (500, 3))
    F = rng.integers(0, 500, (900, 3))
    p = tmp_path / "r.obj"
    with open(p, "w") as f:
        f.write("# List of vertices follows\n")
        for v in V:
            f.write("v %r %r %r\n" % tuple(float(x) for x in v))
        for k, t in enumerate(V[:, :2]):
            f.write("vt %.6f %.6f\n" % tuple(t))
        for a, b, c in F:
            f.write("f %d/%d %d/%d %d/%d\n" % (a + 1, a + 1, b + 1, b + 1, c + 1, c + 1))
    gV, gF = A.read_obj(p)
    wV, wF = RS.parse_obj(p)
    assert gV.tobytes() == wV.tobytes() == V.tobytes()
    assert gF.tobytes() == wF.tobytes()


def test_read_obj_errors(tmp_path):
    p = tmp_path / "bad.obj"
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n")
    with pytest.raises(A.BrdfGpuError):
        A.read_obj(p)
    p.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2\n")
    with pytest.raises(A.BrdfGpuError):
        A.read_obj(p)


def _png_bytes(img, ctype, filters, palette=None, idat_split=1):
    """A PNG file written by hand with chosen per-row filter types (so every unfilter path is used)."""
    h, w = img.shape[:2]
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    rows = img.reshape(h, w * ch).astype(np.int32)
    raw = bytearray()
    prev = np.zeros(w * ch, dtype=np.int32)
    for y in range(h):
        ft = filters[y % len(filters)]
        cur = rows[y]
        left = np.concatenate([np.zeros(ch, dtype=np.int32), cur[:-ch]])
        ul = np.concatenate([np.zeros(ch, dtype=np.int32), prev[:-ch]])
        if ft == 0:
            pred = 0
        elif ft == 1:
            pred = left
        elif ft == 2:
            pred = prev
        elif ft == 3:
            pred = (left + prev) // 2
        else:
            pp = left + prev - ul
            pa, pb, pc = np.abs(pp - left), np.abs(pp - prev), np.abs(pp - ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
        raw.append(ft)
        raw += bytes(((cur - pred) & 255).astype(np.uint8))
        prev = cur

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))
    z = zlib.compress(bytes(raw), 6)
    cuts = [len(z) * k // idat_split for k in range(idat_split + 1)]
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0))
    if palette is not None:
        out += chunk(b"PLTE", bytes(palette.astype(np.uint8).ravel()))
    out += chunk(b"tEXt", b"Comment\x00hand made")
    for a, b in zip(cuts[:-1], cuts[1:]):
        out += chunk(b"IDAT", z[a:b])
    return out + chunk(b"IEND", b"")


@pytest.mark.parametrize("ctype", [2, 6, 0, 4, 3])
def test_read_png_every_filter_and_colour_type(tmp_path, ctype):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(10 + ctype)
    h, w = 37, 53
    ch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    img = rng.integers(0, 256, (h, w, ch), dtype=np.uint8)
    # smooth part so that Paeth / average predictors see every branch
    img[:, :, 0] = (np.add.outer(np.arange(h), np.arange(w)) * 3) % 256
    palette = rng.integers(0, 256, (256, 3)) if ctype == 3 else None
    p = tmp_path / ("t%d.png" % ctype)
    p.write_bytes(_png_bytes(img, ctype, filters=[0, 1, 2, 3, 4], palette=palette, idat_split=3))
    got = A.read_png(p)
    want = cv2.imread(str(p), cv2.IMREAD_COLOR)     # the decode the reference uses
    assert got.shape == want.shape == (h, w, 3)
    assert got.tobytes() == want.tobytes()


def test_read_png_written_by_cv2(tmp_path):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (120, 160, 3), dtype=np.uint8)
    p = tmp_path / "w.png"
    assert cv2.imwrite(str(p), img)
    assert A.read_png(p).tobytes() == img.tobytes()


def test_read_png_rejects_what_it_cannot_decode(tmp_path):
    p = tmp_path / "x.png"
    p.write_bytes(b"not a png at all, just text that is long enough to pass the size check......")
    with pytest.raises(A.BrdfGpuError):
        A.read_png(p)
    img = np.zeros((4, 4, 3), dtype=np.uint8)
    data = bytearray(_png_bytes(img, 2, filters=[0]))
    data[24] = 16       # bit depth 16
    p.write_bytes(bytes(data))
    with pytest.raises(A.BrdfGpuError):
        A.read_png(p)


def _rewrite_ihdr(data, w, h):
    """IHDR with new dimensions and a VALID chunk CRC (a crafted file, not a damaged one)."""
    out = bytearray(data)
    out[16:24] = struct.pack(">II", w, h)
    out[29:33] = struct.pack(">I", zlib.crc32(bytes(out[12:29])))
    return bytes(out)


def test_read_png_survives_corrupt_and_crafted_files(tmp_path):
    """A huge IHDR must not become a huge allocation (or an exception through the C boundary) and damaged chunks
    are reported: cv::imread returns an empty Mat there and the reference fails cleanly."""
    img = np.arange(4 * 4 * 3, dtype=np.uint8).reshape(4, 4, 3)
    good = _png_bytes(img, 2, filters=[0])
    p = tmp_path / "c.png"
    for w, h in ((0x7fffffff, 0x7fffffff), (65535, 65535), (70000, 4), (4, 4000)):
        p.write_bytes(_rewrite_ihdr(good, w, h))
        with pytest.raises(A.BrdfGpuError):
            A.read_png(p)
        msg = A.lib().brdfgpu_last_error(None).decode()
        assert "IDAT" in msg or "dimension" in msg or "inflate" in msg, msg
    flipped = bytearray(good)
    flipped[len(good) - 20] ^= 0x40      # inside the IDAT payload: its CRC no longer matches
    p.write_bytes(bytes(flipped))
    with pytest.raises(A.BrdfGpuError):
        A.read_png(p)
    assert "CRC" in A.lib().brdfgpu_last_error(None).decode()
    p.write_bytes(good)
    assert A.read_png(p).tobytes() == img[:, :, ::-1].tobytes()


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "img")), reason="reference tree not mounted")
@pytest.mark.parametrize("scene", ["cup", "bunny"])
def test_reference_files(scene):
    """The reference's own .obj / .cal / .png files: byte-identical to the Python readers that built the
    scene cache the gather parity tests use (tests/real_scenes.py)."""
    cv2 = pytest.importorskip("cv2")
    d = os.path.join(REF, "img", scene)
    V, F = A.read_obj(os.path.join(d, scene + ".obj"))
    wV, wF = RS.parse_obj(os.path.join(d, scene + ".obj"))
    assert V.tobytes() == wV.tobytes() and F.tobytes() == wF.tobytes()
    cals = [os.path.join(d, scene + ".cal")] + [os.path.join(REF, "Camera Calibrations", f)
                                                for f in sorted(os.listdir(os.path.join(REF, "Camera Calibrations")))]
    for c in cals:
        cam, mask = A.read_cal(c)
        assert mask == 0xFFFF and cam.tobytes() == RS.parse_cal(c).tobytes(), c
    for name in ("1.png", "9.png", "16.png", "dark.png"):
        got = A.read_png(os.path.join(d, name))
        assert got.tobytes() == cv2.imread(os.path.join(d, name), cv2.IMREAD_COLOR).tobytes(), name
