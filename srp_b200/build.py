"""Build recipe for srp-b200 (explicit nvcc / gcc command lines, no build system needed).

Artefacts (all git-ignored, all travel to the GPU box with the repo snapshot):

  srp_b200/_build/obj/...           host C objects (gcc -std=c2x, no FP contraction) and
                                    relocatable device objects holding LTO IR for sm_100a
  srp_b200/_build/libsrp.a          the library user programs link: host objects + device
                                    objects; the final executable is device-linked with
                                    `nvcc -dlto` together with the program's own shader twins
  srp_b200/lib/libsrp_b200.so       the same library device-linked against the built-in
                                    program table (csrc/programs): the C-ABI shared object
                                    that tests, smoke() and bench.py load through ctypes
  oracle/_ref/...                   (only where /root/reference exists) the unmodified
                                    reference, see oracle/Makefile
  tests/_build/scenes/<name>        (only where /root/reference exists) the reference's own
                                    tests/scenes/*.c compiled UNMODIFIED against this library
                                    + the device twins srp_b200/twingen.py generates from the
                                    same C sources (tests/_build/twins/)

Device code: -gencode arch=compute_100a,code=lto_100a -rdc=true -fmad=false -lineinfo;
device link: -arch=sm_100a -dlto -Xnvlink -Xnvvm=-fma=0 (no FMA contraction at LTO code
generation either -- bit-exact parity with the reference depends on it).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "srp_b200"
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
OBJ = BUILD / "obj"
LIBDIR = PKG / "lib"
REFERENCE = Path(os.environ.get("SRP_REFERENCE", "/root/reference"))

NVCC = os.environ.get("NVCC", "nvcc")
CC = os.environ.get("CC", "gcc")

INCLUDES = [f"-I{ROOT / 'include'}", f"-I{CSRC}", f"-I{CSRC / 'device'}"]
HOST_CFLAGS = ["-std=c2x", "-O2", "-fPIC", "-ffp-contract=off", "-Wall", "-Wextra", "-Wno-unused-parameter"]
DEVICE_FLAGS = [
    "-std=c++20", "-gencode", "arch=compute_100a,code=lto_100a", "-rdc=true", "-fmad=false",
    "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
]
DLINK_FLAGS = ["-arch=sm_100a", "-dlto", "-Xnvlink", "-Xnvvm=-fma=0", "-lineinfo"]

HOST_SOURCES = sorted((CSRC / "host").glob("*.c"))
DEVICE_SOURCES = sorted((CSRC / "device").glob("*.cu"))
HEADERS = (sorted((ROOT / "include").rglob("*.h")) + sorted((ROOT / "include").rglob("*.cuh"))
           + sorted(CSRC.rglob("*.h")) + sorted(CSRC.rglob("*.cuh")))


def _run(cmd, **kw):
    cmd = [str(c) for c in cmd]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if proc.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + proc.stdout + "\n")
        raise RuntimeError(f"build step failed: {cmd[0]} ... {cmd[-1]}")
    return proc.stdout


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _compile_host(src: Path, out: Path, extra=()):
    if _stale(out, [src, *HEADERS]):
        out.parent.mkdir(parents=True, exist_ok=True)
        _run([CC, *HOST_CFLAGS, *INCLUDES, *extra, "-c", src, "-o", out])
    return out


def _compile_device(src: Path, out: Path, extra=()):
    if _stale(out, [src, *HEADERS]):
        out.parent.mkdir(parents=True, exist_ok=True)
        _run([NVCC, *DEVICE_FLAGS, *INCLUDES, *extra, "-c", src, "-o", out])
    return out


def build_library(jobs: int = 8) -> Path:
    """host + device objects, libsrp.a, and the ctypes-loadable libsrp_b200.so"""
    tasks = []
    with ThreadPoolExecutor(jobs) as pool:
        for s in HOST_SOURCES:
            tasks.append(pool.submit(_compile_host, s, OBJ / "host" / (s.stem + ".o")))
        for s in DEVICE_SOURCES:
            tasks.append(pool.submit(_compile_device, s, OBJ / "device" / (s.stem + ".o")))
        prog_host = pool.submit(_compile_host, CSRC / "programs" / "builtin_host.c", OBJ / "programs" / "builtin_host.o")
        prog_dev = pool.submit(_compile_device, CSRC / "programs" / "builtin_device.cu", OBJ / "programs" / "builtin_device.o")
        lib_objs = [t.result() for t in tasks]
        pack_objs = [prog_host.result(), prog_dev.result()]

    archive = BUILD / "libsrp.a"
    if _stale(archive, lib_objs):
        archive.unlink(missing_ok=True)
        _run(["ar", "rcs", archive, *lib_objs])

    so = LIBDIR / "libsrp_b200.so"
    if _stale(so, lib_objs + pack_objs):
        LIBDIR.mkdir(parents=True, exist_ok=True)
        _run([NVCC, "-shared", *DLINK_FLAGS, "-Xlinker", "-Bsymbolic", *lib_objs, *pack_objs, "-o", so, "-lz"])
    return so


def build_variant(name: str, defines: list[str], jobs: int = 8) -> Path:
    """development aid: libsrp_b200_<name>.so built with extra -D flags (kernel tuning
    experiments; select it with SRP_B200_LIBRARY=<path> when loading)"""
    obj = BUILD / f"obj_{name}"
    with ThreadPoolExecutor(jobs) as pool:
        dev = [pool.submit(_compile_device, s, obj / "device" / (s.stem + ".o"), defines) for s in DEVICE_SOURCES]
        dev.append(pool.submit(_compile_device, CSRC / "programs" / "builtin_device.cu", obj / "programs" / "builtin_device.o", defines))
        dev = [t.result() for t in dev]
    host = [OBJ / "host" / (s.stem + ".o") for s in HOST_SOURCES] + [OBJ / "programs" / "builtin_host.o"]
    so = LIBDIR / f"libsrp_b200_{name}.so"
    _run([NVCC, "-shared", *DLINK_FLAGS, "-Xlinker", "-Bsymbolic", *dev, *host, "-o", so, "-lz"])
    return so


def build_oracle() -> bool:
    """the unmodified reference -> oracle/_ref (needs /root/reference; skipped elsewhere)"""
    if not (REFERENCE / "src").is_dir():
        return False
    _run(["make", "-C", ROOT / "oracle", "-j8", f"REF={REFERENCE}", "all"])
    return True


def build_scenes(jobs: int = 8) -> list[Path]:
    """The reference's own scene programs, relinked unchanged against this library."""
    from . import twingen
    scene_root = REFERENCE / "tests" / "scenes"
    if not scene_root.is_dir():
        return []
    out_dir = ROOT / "tests" / "_build" / "scenes"
    obj_dir = ROOT / "tests" / "_build" / "obj"
    twins = ROOT / "tests" / "_build" / "twins"      # generated from the scenes' own C sources
    for d in (out_dir, obj_dir, twins):
        d.mkdir(parents=True, exist_ok=True)
    archive = BUILD / "libsrp.a"
    save_o = _compile_host(ROOT / "oracle" / "ref_save_raw.c", obj_dir / "save_raw.o")
    objparser_o = obj_dir / "objparser.o"
    if _stale(objparser_o, [REFERENCE / "examples" / "utility" / "objparser.c"]):
        _run([CC, "-std=c2x", "-O2", "-w", f"-I{ROOT / 'include'}", "-c",
              REFERENCE / "examples" / "utility" / "objparser.c", "-o", objparser_o])

    def one(scene_c: Path):
        name = f"{scene_c.parent.name}_{scene_c.stem}"
        twin = twins / f"{name}.cu"
        exe = out_dir / name
        if not _stale(exe, [scene_c, archive, PKG / "twingen.py", *HEADERS]):
            return exe
        # the __device__ twins of the scene's shaders: its own function bodies, by the shader toolchain
        twin.write_text(twingen.generate(scene_c.read_text(), str(scene_c)))
        scene_o = obj_dir / f"{name}.o"
        twin_o = obj_dir / f"{name}_twin.o"
        # the reference's source file, compiled where it lies, with its own C dialect
        _run([CC, "-std=c2x", "-O2", "-w", "-ffp-contract=off", f"-I{ROOT / 'include'}",
              f"-I{REFERENCE / 'examples' / 'utility'}", f"-I{REFERENCE / 'tests' / 'utils'}",
              "-c", scene_c, "-o", scene_o])
        _run([NVCC, *DEVICE_FLAGS, *INCLUDES, f"-I{REFERENCE / 'examples' / 'utility'}", f"-I{REFERENCE / 'tests' / 'utils'}",
              "-c", twin, "-o", twin_o])
        _run([NVCC, *DLINK_FLAGS, scene_o, twin_o, save_o, objparser_o, archive, "-o", exe, "-lz", "-lm"])
        return exe

    with ThreadPoolExecutor(jobs) as pool:
        built = list(pool.map(one, sorted(scene_root.glob("*/*.c"))))
    res = ROOT / "tests" / "_build" / "res"
    if not res.exists() and (REFERENCE / "examples" / "res").is_dir():
        shutil.copytree(REFERENCE / "examples" / "res", res)
    return [b for b in built if b]


def build_all() -> dict:
    so = build_library()
    have_oracle = build_oracle()
    scenes = build_scenes()
    return {"library": str(so), "oracle": have_oracle, "scenes": len(scenes)}


if __name__ == "__main__":
    print(build_all())
