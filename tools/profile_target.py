#!/usr/bin/env python3
"""A few frames of one config and nothing else, for running under ncu:
    ncu ... python tools/profile_target.py cfg4|cfg5|cfg3|cfg2 [frames]"""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from srp_b200 import host as H, scenes as S


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    lib = H.load_product()
    lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
    if which == "cfg5":
        n, size = 1024, 1024
        mesh = S.teapot_mesh()
        draws = [S.teapot_draw(f, mesh) for f in range(n)]
        lib.new_context()
        for fn, *args in draws[0].state:
            getattr(lib.dll, fn)(*args)
        vb = lib.vertex_buffer(mesh[0], 32); ib = lib.index_buffer(mesh[1])
        prog = lib.program("gouraud", S.GOURAUD_VARYINGS, 12)
        fbs = [lib.framebuffer(size, size) for _ in range(n)]
        arr = (C.POINTER(H.SRPFramebuffer) * n)(*[f.ptr for f in fbs])
        uni = np.frombuffer(b"".join(d.uniform for d in draws), dtype=np.uint8).copy()
        prog.set_uniform(draws[0].uniform)
        for _ in range(frames):
            lib.dll.srpB200DrawBatch(ib, vb, arr, n, C.byref(prog.sp), uni.ctypes.data, len(draws[0].uniform),
                                     H.SRP_PRIM_TRIANGLES, 0, len(mesh[1]), 1)
        lib.dll.srpB200Finish()
        print(lib.stats())
        return
    scene = {"cfg4": S.cfg4_subpixel, "cfg3": S.cfg3_shell, "cfg2": S.cfg2_teapot, "cfg1": S.cfg1_textured_cube,
             "cfg3_r12": lambda: S.cfg3_shell(radius=1.2)}[which]()
    p = S.Prepared(lib, scene)
    for _ in range(frames):
        p.draw_all()
    lib.dll.srpB200Finish()
    print(lib.stats())
    p.free()


if __name__ == "__main__":
    main()
