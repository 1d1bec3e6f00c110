"""CPU tests of the I/O neighbours of the draw path (include/srp_b200.h: srpB200LoadOBJ,
srpB200WritePNG): plain host code, checked against the behaviour of the reference's
examples/utility/objparser.c and tests/utils/save.c as restated in numpy / zlib."""
import struct
import zlib

import numpy as np

from srp_b200 import host, scenes as S


def decode_png_rgba(path):
    """minimal reader for what srpB200WritePNG writes: 8-bit RGBA, filter type 0"""
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    at, idat, ihdr = 8, b"", None
    while at < len(data):
        n, typ = struct.unpack(">I4s", data[at:at + 8])
        body = data[at + 8:at + 8 + n]
        (crc,) = struct.unpack(">I", data[at + 8 + n:at + 12 + n])
        assert crc == zlib.crc32(typ + body), typ
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", body)
        elif typ == b"IDAT":
            idat += body
        at += 12 + n
    w, h, depth, ctype, comp, flt, lace = ihdr
    assert (depth, ctype, comp, flt, lace) == (8, 6, 0, 0, 0)
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 4 * w)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 4)


def test_write_png_round_trip(tmp_path):
    lib = host.load_product()
    rng = np.random.RandomState(3)
    color = rng.randint(0, 2 ** 32, size=(37, 53), dtype=np.uint64).astype(np.uint32)
    lib.write_png(tmp_path / "a.png", color)
    rgba = decode_png_rgba(tmp_path / "a.png")
    assert np.array_equal(rgba[..., 0], (color >> 24) & 255)
    assert np.array_equal(rgba[..., 1], (color >> 16) & 255)
    assert np.array_equal(rgba[..., 2], (color >> 8) & 255)
    assert (rgba[..., 3] == 255).all()                      # tests/utils/save.c:26: alpha forced to 0xFF
    assert lib.dll.srpB200WritePNG(str(tmp_path / "nodir" / "b.png").encode(), 4, 4, color.ctypes.data) != 0


def test_png_written_here_loads_as_a_texture_source(tmp_path):
    """the writer's output goes through the library's own PNG reader (srp_png.c) unchanged in RGB"""
    import ctypes as C
    lib = host.load_product()
    color = (np.arange(16 * 9, dtype=np.uint32).reshape(9, 16) * 0x01030507) | 0xFF
    lib.write_png(tmp_path / "t.png", color)
    lib.dll.srpLoadPngRgb.restype = C.POINTER(C.c_uint8)
    lib.dll.srpLoadPngRgb.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_char_p)]
    w, h, why = C.c_int(), C.c_int(), C.c_char_p()
    p = lib.dll.srpLoadPngRgb(str(tmp_path / "t.png").encode(), C.byref(w), C.byref(h), C.byref(why))
    assert p and (w.value, h.value) == (16, 9), why.value
    rgb = np.ctypeslib.as_array(p, shape=(9, 16, 3))
    assert np.array_equal(rgb[..., 0], (color >> 24) & 255) and np.array_equal(rgb[..., 2], (color >> 8) & 255)


OBJ_TEXT = """# comment
o thing
v 1.5 -2.25e-1 3
v 0.1 0.2 0.3
v -7 8 9.000001
v 1e-3 +2 -0
vt 0.25 0.75
vt 1 0
vn 0 1 0
vn 0.57735 0.57735 -0.57735
s off
f 1/1/1 2/2/2 3/1/2
f 4/2/1 3/1/1 1/2/2 2/1/1
f 1//1 2//2 3//1
f 1/1/1 2/2/2 9/1/1
"""


def test_load_obj_matches_the_reference_loader_semantics(tmp_path):
    lib = host.load_product()
    path = tmp_path / "m.obj"
    path.write_text(OBJ_TEXT)
    v, i = lib.load_obj(path)
    # objparser.c: a quad contributes its first three corners; `1//1` and out-of-range faces are skipped
    want = np.array([
        [1.5, -0.225, 3, 0.25, 0.75, 0, 1, 0], [0.1, 0.2, 0.3, 1, 0, 0.57735, 0.57735, -0.57735],
        [-7, 8, 9.000001, 0.25, 0.75, 0.57735, 0.57735, -0.57735],
        [1e-3, 2, -0.0, 1, 0, 0, 1, 0], [-7, 8, 9.000001, 0.25, 0.75, 0, 1, 0], [1.5, -0.225, 3, 1, 0, 0.57735, 0.57735, -0.57735],
    ], np.float32)
    assert v.shape == (6, 8) and np.array_equal(v.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(i, np.arange(6, dtype=np.uint32))
    assert any("skipped" in m[3] for m in lib.messages)
    try:
        lib.load_obj(tmp_path / "missing.obj")
        assert False, "a missing file must fail"
    except OSError:
        pass


def _reference_load_obj(reference, path):
    """the reference's own loader (examples/utility/objparser.c: loadOBJMesh), linked into the oracle library"""
    import ctypes as C

    class OBJMesh(C.Structure):
        _fields_ = [("vertices", C.POINTER(C.c_float)), ("vertexCount", C.c_size_t),
                    ("indices", C.POINTER(C.c_uint32)), ("indexCount", C.c_size_t)]
    m = OBJMesh()
    reference.dll.loadOBJMesh.restype = C.c_bool
    reference.dll.loadOBJMesh.argtypes = [C.c_char_p, C.POINTER(OBJMesh)]
    reference.dll.freeOBJMesh.argtypes = [C.POINTER(OBJMesh)]
    assert reference.dll.loadOBJMesh(str(path).encode(), C.byref(m))
    v = np.ctypeslib.as_array(m.vertices, shape=(int(m.vertexCount), 8)).copy()
    i = np.ctypeslib.as_array(m.indices, shape=(int(m.indexCount),)).copy()
    reference.dll.freeOBJMesh(C.byref(m))
    return v, i


def test_load_obj_equals_the_reference_loader(reference, tmp_path):
    """srpB200LoadOBJ against the reference's loadOBJMesh itself: the teapot asset and the
    hand-written file above (quads, missing uv, out-of-range index) give the same arrays bit for bit"""
    import pytest
    lib = host.load_product()
    small = tmp_path / "m.obj"
    small.write_text(OBJ_TEXT.replace("f 1/1/1 2/2/2 9/1/1\n", ""))     # (the reference reads out of range there: undefined)
    paths = [small]
    try:
        paths.append(S.asset_path("objects/utah_teapot.obj"))
    except FileNotFoundError:
        pass
    for path in paths:
        v, i = lib.load_obj(path)
        wv, wi = _reference_load_obj(reference, path)
        assert v.shape == wv.shape and np.array_equal(v.view(np.uint32), wv.view(np.uint32)), path
        assert np.array_equal(i, wi), path
    if len(paths) == 2:
        assert lib.load_obj(paths[1])[0].shape == (3498, 8)
        # the Python description of the scenes uses the same semantics
        sv, si = S.load_obj(paths[1])
        assert np.array_equal(sv.view(np.uint32), wv.view(np.uint32)) and np.array_equal(si, wi)
