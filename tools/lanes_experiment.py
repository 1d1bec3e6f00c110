"""Experiment: do two independent instances of the library (two streams, two sets of scratch pools) in one
process overlap the low-occupancy front-end of frame i+1 with the tile kernel of frame i?  Wall clock over
many frames, one instance vs two instances fed alternately."""
import os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srp_b200 import host as H, scenes as S

def main():
    src = H.PRODUCT_SO
    libs = [H.SrpLibrary(src, is_product=True)]
    dst = src.with_name("libsrp_b200_lane1.so")
    shutil.copy(src, dst)
    libs.append(H.SrpLibrary(dst, is_product=True))
    scene = S.cfg3_shell(3840, 2160)
    preps = []
    for lib in libs:
        lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
        preps.append(S.Prepared(lib, scene))
    def finish():
        for lib in libs: lib.dll.srpB200Finish()
    for p in preps:
        for _ in range(3): p.draw_all()
    finish()
    N = 400
    for name, order in (("one", [0]), ("two", [0, 1]), ("one", [0]), ("two", [0, 1])):
        t = time.perf_counter()
        for k in range(N):
            preps[order[k % len(order)]].draw_all()
        finish()
        dt = time.perf_counter() - t
        print(name, "frames/s", round(N / dt, 1), "ms/frame", round(1e3 * dt / N, 4), flush=True)
    os.remove(dst)

main()
