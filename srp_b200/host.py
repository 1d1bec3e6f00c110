"""Host-side mirror of the srp C API (ctypes), the way a C program would call it.

`SrpLibrary` wraps any shared object that exports the srp C ABI (include/srp/api.h).  The
product is srp_b200/lib/libsrp_b200.so (host C layer + sm_100a kernels); `load_product()` loads
it and raises if it is missing -- there is no CPU fallback on the product path and nothing in
this package knows where the oracle lives.  The tests drive the unmodified reference through
the same class (oracle/refhost.py, test infrastructure), so a parity test issues literally the
same calls against both and compares the framebuffer planes.

Function names, argument meaning and error behaviour are the reference's
(include/srp/{context,buffer,framebuffer,texture,shaders}.h); numpy only carries bytes.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
PRODUCT_SO = ROOT / "srp_b200" / "lib" / "libsrp_b200.so"

# ---- enums (include/srp/api.h) ------------------------------------------------------
SRP_FLOAT, SRP_DOUBLE, SRP_INT8, SRP_INT16, SRP_INT32, SRP_INT64, SRP_UINT8, SRP_UINT16, SRP_UINT32, SRP_UINT64 = range(10)
SRP_PROVOKING_VERTEX_FIRST, SRP_PROVOKING_VERTEX_LAST = 0, 1
SRP_WINDING_CCW, SRP_WINDING_CW = 0, 1
SRP_FACE_NONE, SRP_FACE_FRONT, SRP_FACE_BACK, SRP_FACE_FRONT_AND_BACK = range(4)
SRP_POLYGON_MODE_FILL, SRP_POLYGON_MODE_LINE, SRP_POLYGON_MODE_POINT = range(3)
(SRP_COMPARE_NEVER, SRP_COMPARE_ALWAYS, SRP_COMPARE_LESS, SRP_COMPARE_LEQUAL, SRP_COMPARE_GREATER,
 SRP_COMPARE_GEQUAL, SRP_COMPARE_EQUAL, SRP_COMPARE_NOTEQUAL) = range(8)
(SRP_STENCIL_KEEP, SRP_STENCIL_ZERO, SRP_STENCIL_REPLACE, SRP_STENCIL_INCR, SRP_STENCIL_INCR_WRAP,
 SRP_STENCIL_DECR, SRP_STENCIL_DECR_WRAP, SRP_STENCIL_INVERT) = range(8)
SRP_INTERPOLATION_MODE_PERSPECTIVE, SRP_INTERPOLATION_MODE_AFFINE, SRP_INTERPOLATION_MODE_FLAT = range(3)
(SRP_PRIM_POINTS, SRP_PRIM_LINES, SRP_PRIM_LINE_STRIP, SRP_PRIM_LINE_LOOP, SRP_PRIM_TRIANGLES,
 SRP_PRIM_TRIANGLE_STRIP, SRP_PRIM_TRIANGLE_FAN) = range(7)
TW_REPEAT, TW_CLAMP_TO_EDGE = 0, 1
SRP_MESSAGE_ERROR, SRP_MESSAGE_WARNING = 0, 1
SRP_B200_SYNC_DRAW, SRP_B200_SYNC_EXPLICIT = 0, 1

_INDEX_TYPES = {np.dtype("u1"): SRP_UINT8, np.dtype("u2"): SRP_UINT16, np.dtype("u4"): SRP_UINT32, np.dtype("u8"): SRP_UINT64}


# ---- structs ------------------------------------------------------------------------
class SRPVaryingInfo(C.Structure):
    _fields_ = [("nItems", C.c_size_t), ("type", C.c_int), ("interpolationMode", C.c_int)]


class SRPVertexShader(C.Structure):
    _fields_ = [("shader", C.c_void_p), ("nVaryings", C.c_size_t),
                ("varyingsInfo", C.POINTER(SRPVaryingInfo)), ("varyingsSize", C.c_size_t)]


class SRPFragmentShader(C.Structure):
    _fields_ = [("shader", C.c_void_p), ("mayOverwriteDepth", C.c_bool)]


class SRPShaderProgram(C.Structure):
    _fields_ = [("uniform", C.c_void_p), ("vs", C.POINTER(SRPVertexShader)), ("fs", C.POINTER(SRPFragmentShader))]


class SRPFramebuffer(C.Structure):
    _fields_ = [("width", C.c_size_t), ("height", C.c_size_t), ("size", C.c_size_t),
                ("color", C.POINTER(C.c_uint32)), ("depth", C.POINTER(C.c_float)), ("stencil", C.POINTER(C.c_uint8))]


class SRPMessageCallback(C.Structure):
    _fields_ = [("func", C.c_void_p), ("userParameter", C.c_void_p)]


class Mat4(C.Structure):
    _fields_ = [("data", C.c_float * 16)]

    def numpy(self):
        return np.array(self.data, dtype=np.float32).reshape(4, 4)


class SrpbProgramInfo(C.Structure):
    _fields_ = [("name", C.c_char_p), ("vs", C.c_void_p), ("fs", C.c_void_p),
                ("uniformSize", C.c_size_t), ("deviceId", C.c_int)]


class SRPB200Stats(C.Structure):
    _fields_ = [(n, C.c_ulonglong) for n in (
        "draws", "primsIn", "primsEmitted", "primsStored", "fragsEmitted", "fragsShaded",
        "kernelLaunches", "h2dBytes", "d2hBytes", "overflow", "subDraws")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class SRPB200Mesh(C.Structure):
    _fields_ = [("vertices", C.POINTER(C.c_float)), ("vertexCount", C.c_size_t), ("bytesPerVertex", C.c_size_t),
                ("indices", C.POINTER(C.c_uint32)), ("indexCount", C.c_size_t)]


MESSAGE_FUNC = C.CFUNCTYPE(None, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_void_p)


class Program:
    """An SRPShaderProgram built from one of the library's built-in shader pairs."""

    def __init__(self, lib, name, varyings, varyings_size, may_overwrite_depth=False):
        info = lib.dll.srpbFindProgram(name.encode())
        if not info:
            raise KeyError(f"no built-in program {name!r}")
        self.name = name
        self.uniform_size = int(info.contents.uniformSize)
        self._infos = (SRPVaryingInfo * max(1, len(varyings)))()
        for i, (n_items, typ, mode) in enumerate(varyings):
            self._infos[i] = SRPVaryingInfo(n_items, typ, mode)
        self.vs = SRPVertexShader(info.contents.vs, len(varyings),
                                  C.cast(self._infos, C.POINTER(SRPVaryingInfo)) if varyings else None, varyings_size)
        self.fs = SRPFragmentShader(info.contents.fs, may_overwrite_depth)
        self._uniform = None
        self.sp = SRPShaderProgram(None, C.pointer(self.vs), C.pointer(self.fs))

    def set_uniform(self, data: bytes | np.ndarray | None):
        """bind (a copy of) the uniform block; None binds a NULL uniform"""
        if data is None:
            self._uniform = None
            self.sp.uniform = None
            return
        raw = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8)).copy()
        self._uniform = raw
        self.sp.uniform = raw.ctypes.data

    def set_interpolation(self, index, mode):
        self._infos[index].interpolationMode = mode


class Framebuffer:
    def __init__(self, lib, ptr):
        self.lib = lib
        self.ptr = ptr
        self.width = int(ptr.contents.width)
        self.height = int(ptr.contents.height)

    def planes(self, download=True):
        """copies of (color u32, depth bits u32, stencil u8), each [H, W]"""
        if download and self.lib.is_product:
            self.lib.dll.srpB200FramebufferDownload(self.ptr)
        n = self.width * self.height
        f = self.ptr.contents
        color = np.ctypeslib.as_array(f.color, shape=(n,)).copy().reshape(self.height, self.width)
        depth = np.ctypeslib.as_array(f.depth, shape=(n,)).view(np.uint32).copy().reshape(self.height, self.width)
        stencil = np.ctypeslib.as_array(f.stencil, shape=(n,)).copy().reshape(self.height, self.width)
        return color, depth, stencil

    def clear(self):
        self.lib.dll.srpFramebufferClear(self.ptr)

    def free(self):
        if self.ptr:
            self.lib.dll.srpFreeFramebuffer(self.ptr)
            self.ptr = None


class SrpLibrary:
    """One loaded implementation of the srp C API."""

    def __init__(self, path: Path, is_product: bool):
        self.path = Path(path)
        self.is_product = is_product
        self.dll = C.CDLL(str(path), mode=os.RTLD_LOCAL | os.RTLD_NOW if hasattr(os, "RTLD_NOW") else C.DEFAULT_MODE)
        self._declare()
        self._callback = None
        self.messages: list[tuple[int, int, str, str]] = []
        self.new_context()

    # -- prototypes -------------------------------------------------------------------
    def _declare(self):
        d = self.dll
        vp, sz, i32, u8, f32, b = C.c_void_p, C.c_size_t, C.c_int, C.c_uint8, C.c_float, C.c_bool
        proto = {
            "srpNewContext": (None, [vp]),
            "srpSetMessageCallback": (None, [SRPMessageCallback]),
            "srpProvokingVertexMode": (None, [i32]), "srpRasterCullFace": (None, [i32]),
            "srpRasterFrontFace": (None, [i32]), "srpRasterPolygonMode": (None, [i32]),
            "srpRasterPointSize": (None, [f32]),
            "srpScissorTest": (None, [b]), "srpScissorOptions": (None, [sz, sz, sz, sz]),
            "srpStencilTest": (None, [b]), "srpStencilFunc": (None, [i32, u8, u8]),
            "srpStencilFuncSeparate": (None, [i32, i32, u8, u8]), "srpStencilOp": (None, [i32, i32, i32]),
            "srpStencilOpSeparate": (None, [i32, i32, i32, i32]), "srpStencilWriteMask": (None, [u8]),
            "srpStencilWriteMaskSeparate": (None, [i32, u8]),
            "srpDepthTest": (None, [b]), "srpDepthWrite": (None, [b]), "srpDepthCompareOp": (None, [i32]),
            "srpNewVertexBuffer": (vp, []), "srpFreeVertexBuffer": (None, [vp]),
            "srpVertexBufferCopyData": (None, [vp, sz, sz, vp]),
            "srpDrawVertexBuffer": (None, [vp, C.POINTER(SRPFramebuffer), C.POINTER(SRPShaderProgram), i32, sz, sz]),
            "srpNewIndexBuffer": (vp, []), "srpFreeIndexBuffer": (None, [vp]),
            "srpIndexBufferCopyData": (None, [vp, i32, sz, vp]),
            "srpDrawIndexBuffer": (None, [vp, vp, C.POINTER(SRPFramebuffer), C.POINTER(SRPShaderProgram), i32, sz, sz]),
            "srpNewFramebuffer": (C.POINTER(SRPFramebuffer), [sz, sz]),
            "srpFreeFramebuffer": (None, [C.POINTER(SRPFramebuffer)]),
            "srpFramebufferClear": (None, [C.POINTER(SRPFramebuffer)]),
            "srpNewTexture": (vp, [C.c_char_p, i32, i32]), "srpFreeTexture": (None, [vp]),
            "srpTextureGetFilteredColor": (None, [vp, f32, f32, C.POINTER(f32)]),
            "srpTextureGet": (i32, [vp, i32]), "srpTextureSet": (None, [vp, i32, i32]),
            "srpB200NewTextureFromMemory": (vp, [vp, i32, i32, i32, i32]),
            "srpbFindProgram": (C.POINTER(SrpbProgramInfo), [C.c_char_p]),
            "srpbIsReferenceBuild": (i32, []),
            "mat4ConstructIdentity": (Mat4, []),
            "mat4ConstructScale": (Mat4, [f32] * 3), "mat4ConstructTranslate": (Mat4, [f32] * 3),
            "mat4ConstructRotate": (Mat4, [f32] * 3), "mat4ConstructTRS": (Mat4, [f32] * 9),
            "mat4ConstructView": (Mat4, [f32] * 9),
            "mat4ConstructOrthogonalProjection": (Mat4, [f32] * 6),
            "mat4ConstructPerspectiveProjection": (Mat4, [f32] * 6),
        }
        product_only = {
            "srpB200RegisterProgram": (i32, [vp, vp, i32, sz]),
            "srpB200SetSyncMode": (None, [i32]), "srpB200GetSyncMode": (i32, []),
            "srpB200Finish": (None, []),
            "srpB200SetMirrorPlanes": (None, [i32]),
            "srpB200FramebufferDownload": (None, [C.POINTER(SRPFramebuffer)]),
            "srpB200FramebufferUpload": (None, [C.POINTER(SRPFramebuffer)]),
            "srpB200FramebufferDownloadAsync": (None, [C.POINTER(SRPFramebuffer)]),
            "srpB200FramebufferWait": (None, [C.POINTER(SRPFramebuffer)]),
            "srpB200FramebufferFence": (None, [C.POINTER(SRPFramebuffer)]),
            "srpB200NewFramebufferOnDevice": (C.POINTER(SRPFramebuffer), [sz, sz, vp, vp, vp]),
            "srpB200FramebufferDevicePlane": (vp, [C.POINTER(SRPFramebuffer), i32]),
            "srpB200LoadOBJ": (i32, [C.c_char_p, C.POINTER(SRPB200Mesh)]),
            "srpB200FreeMesh": (None, [C.POINTER(SRPB200Mesh)]),
            "srpB200WritePNG": (i32, [C.c_char_p, sz, sz, vp]),
            "srpB200SaveFramebufferPNG": (i32, [C.POINTER(SRPFramebuffer), C.c_char_p]),
            "srpB200DrawBatch": (None, [vp, vp, C.POINTER(C.POINTER(SRPFramebuffer)), sz,
                                        C.POINTER(SRPShaderProgram), vp, sz, i32, sz, sz, i32]),
            "srpB200SetRowRange": (None, [sz, sz]),
            "srpB200DeviceAlloc": (vp, [sz]), "srpB200DeviceFree": (None, [vp]),
            "srpB200IpcExport": (i32, [vp, vp]), "srpB200IpcOpen": (vp, [vp]), "srpB200IpcClose": (None, [vp]),
            "srpB200StreamSignal": (None, [vp, C.c_uint32]), "srpB200StreamWait": (None, [vp, C.c_uint32]),
            "srpB200TileWidth": (sz, []), "srpB200TileHeight": (sz, []),
            "srpB200GetStats": (None, [C.POINTER(SRPB200Stats)]), "srpB200ResetStats": (None, []),
            "srpB200SetProfiling": (None, [i32]),
            "srpB200CollectStageTimes": (C.c_ulonglong, [C.POINTER(C.c_double)]),
            "srpB200Version": (C.c_char_p, []), "srpB200SetDevice": (None, [i32]),
            "srpB200Stream": (vp, []),
            "srpB200StreamWaitAll": (None, [vp, C.c_uint32, C.c_uint32]),
            "srpB200SetLane": (i32, [i32]), "srpB200GetLane": (i32, []), "srpB200LaneCount": (i32, []),
        }
        if self.is_product:
            proto.update(product_only)
        for name, (res, args) in proto.items():
            fn = getattr(d, name)      # AttributeError here = the library does not export the symbol
            fn.restype = res
            fn.argtypes = args
        self.exported = sorted(proto)

    # -- context ------------------------------------------------------------------------
    def new_context(self):
        """srpNewContext(&srpContext) on the library's own context object + capture messages"""
        ctx = C.c_char.in_dll(self.dll, "srpContext")
        self.dll.srpNewContext(C.addressof(ctx))
        self.messages.clear()

        def on_message(mtype, severity, func, text, _user):
            self.messages.append((mtype, severity, (func or b"").decode(), (text or b"").decode()))

        self._callback = MESSAGE_FUNC(on_message)
        self.dll.srpSetMessageCallback(SRPMessageCallback(C.cast(self._callback, C.c_void_p), None))

    # -- objects ------------------------------------------------------------------------
    def program(self, name, varyings=(), varyings_size=0, may_overwrite_depth=False) -> Program:
        return Program(self, name, list(varyings), varyings_size, may_overwrite_depth)

    def vertex_buffer(self, data: np.ndarray, stride: int):
        vb = self.dll.srpNewVertexBuffer()
        self.vertex_buffer_copy(vb, data, stride)
        return vb

    def vertex_buffer_copy(self, vb, data: np.ndarray, stride: int):
        raw = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        self.dll.srpVertexBufferCopyData(vb, stride, raw.size, raw.ctypes.data)

    def index_buffer(self, indices: np.ndarray):
        ib = self.dll.srpNewIndexBuffer()
        self.index_buffer_copy(ib, indices)
        return ib

    def index_buffer_copy(self, ib, indices: np.ndarray):
        idx = np.ascontiguousarray(indices)
        self.dll.srpIndexBufferCopyData(ib, _INDEX_TYPES[idx.dtype], idx.nbytes, idx.ctypes.data)

    def framebuffer(self, width, height) -> Framebuffer:
        ptr = self.dll.srpNewFramebuffer(width, height)
        if not ptr:
            raise RuntimeError(f"srpNewFramebuffer({width}, {height}) failed: {self.messages[-1:] or 'no message'}")
        if not self.is_product:
            # The reference never initialises (or clears) the stencil plane and relies on
            # fresh zero pages (src/core/framebuffer.c:25,57-62).  The
            # harness gives it the zeros it assumes, so the oracle is deterministic.
            C.memset(ptr.contents.stencil, 0, width * height)
        return Framebuffer(self, ptr)

    def texture_from_memory(self, rgb: np.ndarray, wrap_x=TW_REPEAT, wrap_y=TW_REPEAT):
        rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
        h, w, c = rgb.shape
        assert c == 3
        return self.dll.srpB200NewTextureFromMemory(rgb.ctypes.data, w, h, wrap_x, wrap_y)

    def draw(self, fb: Framebuffer, prog: Program, primitive, start, count, vb, ib=None):
        if ib is None:
            self.dll.srpDrawVertexBuffer(vb, fb.ptr, C.byref(prog.sp), primitive, start, count)
        else:
            self.dll.srpDrawIndexBuffer(ib, vb, fb.ptr, C.byref(prog.sp), primitive, start, count)

    # -- math helpers (host-side constructors of the library itself) ----------------------
    def mat4(self, fn, *args) -> np.ndarray:
        return getattr(self.dll, fn)(*[float(a) for a in args]).numpy()

    # -- product-only ---------------------------------------------------------------------
    def load_obj(self, path):
        """srpB200LoadOBJ -> (vertices [n, 8] float32, indices uint32), copies"""
        m = SRPB200Mesh()
        if self.dll.srpB200LoadOBJ(str(path).encode(), C.byref(m)) != 0:
            raise OSError(f"srpB200LoadOBJ({path}) failed: {self.messages[-1:]}")
        try:
            n = int(m.vertexCount)
            v = np.ctypeslib.as_array(m.vertices, shape=(n, 8)).copy() if n else np.zeros((0, 8), np.float32)
            i = np.ctypeslib.as_array(m.indices, shape=(int(m.indexCount),)).copy() if n else np.zeros(0, np.uint32)
        finally:
            self.dll.srpB200FreeMesh(C.byref(m))
        return v, i

    def write_png(self, path, color: np.ndarray):
        """srpB200WritePNG of a [H, W] uint32 colour plane (R in the top byte)"""
        c = np.ascontiguousarray(color, dtype=np.uint32)
        if self.dll.srpB200WritePNG(str(path).encode(), c.shape[1], c.shape[0], c.ctypes.data) != 0:
            raise OSError(f"srpB200WritePNG({path}) failed")

    def stage_times(self) -> dict:
        ms = (C.c_double * 3)()
        n = int(self.dll.srpB200CollectStageTimes(ms))
        return {"draws": n, "geometry_ms": ms[0], "binning_ms": ms[1], "tiles_ms": ms[2]}

    def stats(self) -> dict:
        s = SRPB200Stats()
        self.dll.srpB200GetStats(C.byref(s))
        return s.as_dict()


_loaded: dict[str, SrpLibrary] = {}


def load_product() -> SrpLibrary:
    """The CUDA library.  Raises if it has not been built -- there is no fallback."""
    if "product" not in _loaded:
        so = Path(os.environ.get("SRP_B200_LIBRARY", PRODUCT_SO))   # tuning experiments may point at a variant build
        if so != PRODUCT_SO:
            _loaded["product"] = SrpLibrary(so, is_product=True)
            return _loaded["product"]
        if not PRODUCT_SO.exists():
            raise RuntimeError(
                f"{PRODUCT_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "srp_b200 has no CPU implementation to fall back to.")
        _loaded["product"] = SrpLibrary(PRODUCT_SO, is_product=True)
    return _loaded["product"]
