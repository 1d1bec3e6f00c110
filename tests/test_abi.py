"""CPU tests: the drop-in boundary.  The product library must load without a GPU (CUDA is
initialised lazily) and export every function include/srp/api.h and include/srp_b200.h
declare; public struct layouts are the reference's."""
import ctypes as C
import re
import subprocess

import pytest

from conftest import ROOT
from srp_b200 import host


def declared_functions(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(srp[A-Za-z0-9_]+)\s*\(", text)
    return sorted(set(n for n in names if not n.endswith("Func")))


def exported(so):
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(so)], text=True)
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


def test_library_is_built():
    assert host.PRODUCT_SO.exists(), "run __graft_entry__.build() first"


def test_exports_every_declared_entry_point():
    syms = exported(host.PRODUCT_SO)
    for header in ("srp/api.h", "srp_b200.h"):
        missing = [f for f in declared_functions(header) if f not in syms]
        assert not missing, f"{header}: not exported: {missing}"
    math = ["mat4MultiplyVec4", "mat4MultiplyMat4", "mat4ConstructIdentity", "mat4ConstructScale", "mat4ConstructTranslate",
            "mat4ConstructRotate", "mat4ConstructTRS", "mat4ConstructView", "mat4ConstructOrthogonalProjection",
            "mat4ConstructPerspectiveProjection"]
    math += [f"vec{n}{op}" for n in (2, 3, 4) for op in
             ("Add", "Subtract", "DotProduct", "MultiplyScalar", "Normalize", "Reflect", "Negate", f"MultiplyVec{n}")]
    assert not [m for m in math if m not in syms]


def test_loads_without_gpu_and_fails_loudly(capfd):
    lib = host.load_product()                      # no CUDA call yet
    assert b"sm_100a" in lib.dll.srpB200Version()
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-GPU failure mode cannot be observed")
    lib.new_context()
    fb = lib.dll.srpNewFramebuffer(64, 64)
    assert not fb, "without a GPU there is nothing to fall back to"
    assert any("no CUDA device" in m[3] or "CUDA" in m[3] for m in lib.messages), lib.messages
    assert "srp-b200" in capfd.readouterr().err


def test_context_defaults_and_quirks():
    lib = host.load_product()
    lib.new_context()

    class Raster(C.Structure):
        _fields_ = [("frontFace", C.c_int), ("cullFace", C.c_int), ("polygonMode", C.c_int), ("pointSize", C.c_float)]

    class Scissor(C.Structure):
        _fields_ = [("enabled", C.c_bool), ("x", C.c_size_t), ("y", C.c_size_t), ("w", C.c_size_t), ("h", C.c_size_t)]

    class Face(C.Structure):
        _fields_ = [("func", C.c_int), ("ref", C.c_uint8), ("mask", C.c_uint8), ("writeMask", C.c_uint8),
                    ("sfail", C.c_int), ("dfail", C.c_int), ("passOp", C.c_int)]

    class Stencil(C.Structure):
        _fields_ = [("enabled", C.c_bool), ("front", Face), ("back", Face)]

    class Depth(C.Structure):
        _fields_ = [("test", C.c_bool), ("write", C.c_bool), ("op", C.c_int)]

    class Ctx(C.Structure):        # include/srp/api.h SRPContext, 144 bytes
        _fields_ = [("cb", C.c_void_p * 2), ("provoking", C.c_int), ("raster", Raster), ("scissor", Scissor),
                    ("stencil", Stencil), ("depth", Depth), ("arena", C.c_void_p)]
    assert C.sizeof(Ctx) == 144
    ctx = Ctx.in_dll(lib.dll, "srpContext")
    assert (Ctx.raster.offset, Ctx.scissor.offset, Ctx.stencil.offset, Ctx.depth.offset, Ctx.arena.offset) == (20, 40, 80, 124, 136)
    r = ctx.raster
    assert (ctx.provoking, r.frontFace, r.cullFace, r.polygonMode, r.pointSize) == (host.SRP_PROVOKING_VERTEX_LAST, 0, 0, 0, 1.0)
    assert (ctx.depth.test, ctx.depth.write, ctx.depth.op) == (False, True, host.SRP_COMPARE_GREATER)
    assert not ctx.stencil.enabled and ctx.stencil.front.func == host.SRP_COMPARE_ALWAYS
    assert (ctx.stencil.back.mask, ctx.stencil.back.writeMask, ctx.stencil.back.passOp) == (0xFF, 0xFF, host.SRP_STENCIL_KEEP)
    lib.dll.srpStencilTest(False)                  # reference quirk: enables regardless of the argument
    assert ctx.stencil.enabled
    lib.dll.srpScissorOptions(1, 2, 3, 4)
    assert (ctx.scissor.x, ctx.scissor.y, ctx.scissor.w, ctx.scissor.h) == (1, 2, 3, 4)
    lib.dll.srpStencilFuncSeparate(host.SRP_FACE_FRONT, host.SRP_COMPARE_LESS, 7, 0x0F)
    assert ctx.stencil.front.func == host.SRP_COMPARE_LESS and ctx.stencil.back.func == host.SRP_COMPARE_ALWAYS
    lib.dll.srpStencilOpSeparate(host.SRP_FACE_FRONT_AND_BACK, host.SRP_STENCIL_INCR, host.SRP_STENCIL_KEEP, host.SRP_STENCIL_ZERO)
    assert ctx.stencil.front.sfail == ctx.stencil.back.sfail == host.SRP_STENCIL_INCR


def test_host_matrix_constructors_match_the_python_mirror():
    """scenes.py re-derives the matrices in numpy; the library's own C constructors (no FP
    contraction, reference src/math/mat.c order) must give the same bits."""
    import numpy as np
    from srp_b200 import scenes as S
    lib = host.load_product()
    assert np.array_equal(lib.mat4("mat4ConstructRotate", 0.7, 0.35, 0.14), S.rotate(0.7, 0.35, 0.14))
    assert np.array_equal(lib.mat4("mat4ConstructView", 0, 1.75, -7, 0, 0, 0, 1, 1, 1), S.view((0, 1.75, -7)))
    assert np.array_equal(lib.mat4("mat4ConstructPerspectiveProjection", -1, 1, -1, 1, 1, 10), S.perspective(-1, 1, -1, 1, 1, 10))
    assert np.array_equal(lib.mat4("mat4ConstructTRS", 0.2, -0.1, 0, 0, 0, 0.3, 0.6, 0.6, 1),
                          S.trs((0.2, -0.1, 0), (0, 0, 0.3), (0.6, 0.6, 1)))


def test_reference_and_product_host_math_agree(reference):
    import numpy as np
    lib = host.load_product()
    for fn, args in (("mat4ConstructRotate", (1.1, -0.4, 2.5)), ("mat4ConstructOrthogonalProjection", (-2, 3, -1, 1, 0.5, 9)),
                     ("mat4ConstructPerspectiveProjection", (-1, 1, -1, 1, 0.3, 6)), ("mat4ConstructTRS", (1, 2, 3, 0.1, 0.2, 0.3, 2, 2, 0.5))):
        assert np.array_equal(lib.mat4(fn, *args), reference.mat4(fn, *args)), fn


def test_every_kernel_waits_for_its_predecessor():
    """The draw's kernels are launched with the programmatic-stream-serialisation attribute
    (kernels.cuh: srpdLaunchKernel), which is only safe if EVERY kernel blocks in
    `griddepcontrol.wait` before touching memory.  Check the SASS of the shipped library: each
    srpd* kernel contains the wait (ACQBULK) and the early trigger (PREEXIT)."""
    sass = subprocess.run(["cuobjdump", "-sass", str(host.PRODUCT_SO)], capture_output=True, text=True)
    if sass.returncode != 0 or "Function :" not in sass.stdout:
        pytest.skip("cuobjdump not available")
    kernels, cur = {}, None
    for line in sass.stdout.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            kernels[cur] = set()
        elif cur and ("ACQBULK" in line or "PREEXIT" in line):
            kernels[cur].add("ACQBULK" if "ACQBULK" in line else "PREEXIT")
    ours = {k: v for k, v in kernels.items() if "srpd" in k and "Kernel" in k}
    assert len(ours) >= 9, sorted(kernels)
    bad = [k for k, v in ours.items() if v != {"ACQBULK", "PREEXIT"}]
    assert not bad, f"kernels without the dependency prologue: {bad}"
