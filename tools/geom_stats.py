#!/usr/bin/env python3
"""Development helper (run under gpurun): geometry statistics of cfg3 -- how many triangles are
rejected / clipped / stored -- from the library's counters."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from srp_b200 import host as H, scenes as S
lib = H.load_product()
for sc in (S.cfg3_shell(), S.cfg3_shell(radius=1.2) if 'radius' in S.cfg3_shell.__code__.co_varnames else None):
    if sc is None: continue
    p = S.Prepared(lib, sc)
    lib.dll.srpB200ResetStats()
    p.draw_all()
    print(sc.name, lib.stats())
    p.free()
