#!/usr/bin/env python3
"""Development helper (run under gpurun): render scenes with the product and with the
oracle reference and print where they differ."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from srp_b200 import host as H, scenes as S
from oracle.refhost import load_oracle_reference


def diff(name, a, b):
    ok = True
    for plane, x, y in zip(("color", "depth", "stencil"), a, b):
        nd = int((x != y).sum())
        if nd:
            ok = False
            ys, xs = np.nonzero(x != y)
            print(f"  {name}.{plane}: {nd} px differ; first at (x={xs[0]}, y={ys[0]}): got {x[ys[0], xs[0]]:#x} want {y[ys[0], xs[0]]:#x}")
    print(f"{name}: {'OK' if ok else 'MISMATCH'}  covered={int((a[0] != 0).sum())}/{int((b[0] != 0).sum())}")
    return ok


def main():
    prod = H.load_product()
    print(prod.dll.srpB200Version().decode())
    ref = load_oracle_reference()
    which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3s", "cfg4s"]
    table = {
        "cfg1": lambda: S.cfg1_textured_cube(),
        "cfg1s": lambda: S.cfg1_textured_cube(256, 256),
        "cfg2": lambda: S.cfg2_teapot(),
        "cfg2s": lambda: S.cfg2_teapot(512, 512),
        "cfg3s": lambda: S.cfg3_shell(1024, 768, n=96),
        "cfg3": lambda: S.cfg3_shell(),
        "cfg4s": lambda: S.cfg4_subpixel(1024, 512, n=300, n_lines=5000, n_points=5000),
        "cfg4": lambda: S.cfg4_subpixel(),
    }
    allok = True
    for w in which:
        scene = table[w]()
        t0 = time.time(); a = S.render(prod, scene); t1 = time.time()
        b = S.render(ref, scene); t2 = time.time()
        allok &= diff(scene.name, a, b)
        print(f"  product {1e3*(t1-t0):.1f} ms (first call incl. setup), reference {1e3*(t2-t1):.1f} ms; messages: {prod.messages[:3]}")
        print("  stats:", prod.stats())
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()
