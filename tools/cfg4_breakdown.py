import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from srp_b200 import host as H, scenes as S
lib = H.load_product()
scene = S.cfg4_subpixel()
lib.dll.srpB200SetSyncMode(H.SRP_B200_SYNC_EXPLICIT)
p = S.Prepared(lib, scene)
p.draw_all(); lib.dll.srpB200Finish()
lib.dll.srpB200SetProfiling(1)
p.fb.clear()
for d, prog, vb, ib, count in p.items:
    for fn, *args in d.state:
        getattr(lib.dll, fn)(*args)
    lib.dll.srpB200ResetStats(); lib.stage_times()
    lib.draw(p.fb, prog, d.primitive, d.start, count, vb, ib)
    st = lib.stage_times(); s = lib.stats()
    print(d.primitive, count, {k: round(v, 3) for k, v in st.items()}, {k: s[k] for k in ("primsIn", "primsEmitted", "primsStored", "fragsEmitted", "fragsShaded")})
