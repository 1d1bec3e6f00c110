"""CPU tests: the oracle is pinned.

1. oracle/gen_golden.py checked, when it generated tests/golden/scenes, that the oracle
   build reproduces the reference's own committed PNG goldens within the reference's own
   tolerance (tests/compare.py: 1 LSB); here the scene executables of oracle/_ref are re-run
   and must reproduce the committed raw planes bit for bit (build determinism).
2. The reference driven through the shared C ABI (libref_host.so) must reproduce the
   committed synthetic fixtures.
Both are skipped where oracle/_ref does not exist (it is built from /root/reference)."""
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, assert_planes_equal, load_golden

REF = ROOT / "oracle" / "_ref"
SCENES = sorted(p.stem for p in (GOLDEN / "scenes").glob("*.npz"))
SYNTH = sorted(p.stem for p in (GOLDEN / "synthetic").glob("*.npz"))


def read_raw(path):
    with open(path, "rb") as f:
        assert f.read(8)[:7] == b"SRPRAW1"
        w, h = (int(v) for v in np.frombuffer(f.read(16), dtype="<u8"))
        n = w * h
        return (np.frombuffer(f.read(4 * n), dtype="<u4").reshape(h, w),
                np.frombuffer(f.read(4 * n), dtype="<u4").reshape(h, w),
                np.frombuffer(f.read(n), dtype="u1").reshape(h, w))


def test_fixture_inventory():
    assert len(SCENES) == 18, "one fixture per reference scene (tests/scenes of kitrofimov/srp)"
    assert len(SYNTH) >= 50


@pytest.mark.parametrize("name", SCENES)
def test_reference_scene_executable_matches_fixture(name):
    exe = REF / "scenes" / name
    if not exe.exists():
        pytest.skip("oracle/_ref not built")
    with tempfile.TemporaryDirectory() as td:
        out = Path(td) / "o.raw"
        subprocess.check_call([str(exe), str(out)], cwd=REF)
        assert_planes_equal(read_raw(out), load_golden("scenes", name), f"oracle scene {name}")


def test_synthetic_fixtures_reproduce(reference):
    import synthetic_scenes
    from srp_b200 import scenes as S
    scenes = synthetic_scenes.all_scenes()
    assert sorted(scenes) == SYNTH, "tests/golden/synthetic is stale: run oracle/gen_synthetic_golden.py"
    for name in SYNTH[::4]:          # every 4th: keeps the CPU suite short, all are covered on the GPU
        assert_planes_equal(S.render(reference, scenes[name]), load_golden("synthetic", name), f"oracle {name}")
