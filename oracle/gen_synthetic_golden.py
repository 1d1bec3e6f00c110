#!/usr/bin/env python3
"""TEST INFRASTRUCTURE -- golden fixtures for the synthetic scenes, rendered by the REAL
reference (oracle/_ref/libref_host.so = unmodified kitrofimov/srp + the C build of the
built-in shaders).  Writes tests/golden/synthetic/<scene>.npz with the three raw planes
(colour u32, depth bit pattern u32, stencil u8) plus the messages the reference emitted.

Run here (needs oracle/_ref, i.e. /root/reference):   python oracle/gen_synthetic_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from srp_b200 import host as H, scenes as S
from oracle.refhost import load_oracle_reference   # noqa: E402
import synthetic_scenes                        # noqa: E402


def main():
    out = ROOT / "tests" / "golden" / "synthetic"
    out.mkdir(parents=True, exist_ok=True)
    ref = load_oracle_reference()
    assert ref.dll.srpbIsReferenceBuild() == 1
    manifest = {}
    for name, scene in synthetic_scenes.all_scenes().items():
        color, depth, stencil = S.render(ref, scene)
        msgs = [list(m) for m in ref.messages]
        np.savez_compressed(out / f"{name}.npz", color=color, depth=depth, stencil=stencil)
        manifest[name] = {"covered": int((color != 0).sum()), "stencil_nonzero": int((stencil != 0).sum()),
                          "messages": msgs}
        print(f"{name}: covered={manifest[name]['covered']} stencil_nz={manifest[name]['stencil_nonzero']} msgs={len(msgs)}")
    (out / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
