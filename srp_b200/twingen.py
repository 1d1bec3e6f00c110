"""Shader toolchain: the __device__ twins of a program's shaders, generated from its C source.

The srp API passes shaders as host function pointers; the GPU needs `__device__` functions with
the same bodies (include/srp_b200_device.cuh).  A program does not have to restate them by hand:

    python -m srp_b200.twingen app.c -o app_shaders.cu

reads `app.c` (the unmodified source of a program written against kitrofimov/srp), and writes the
one extra CUDA translation unit the program links with:

  * every `typedef struct ... NAME;`, object-like `#define` and `#include "..."` of the file,
    verbatim (vertex, varyings and uniform layouts; the program's own headers are compiled with
    the same include path as the C file),
  * every function with a shader signature -- `void f(SRPVertexShaderIn*, SRPVertexShaderOut*)`
    or `void f(SRPFragmentShaderIn*, SRPFragmentShaderOut*)` -- as `__device__ void srpTwin_f(...)`
    with the SAME body text (other top-level functions of the file, except main, come along as
    `__device__` helpers),
  * the shader tables (SRP_B200_DEFINE_SHADER_TABLES) and one registration per shader
    (srpB200RegisterVertexShader / srpB200RegisterFragmentShader), keyed by the host function it
    stands for; the uniform size of a shader is `sizeof` the struct it casts `in->uniform` to.

The same source therefore runs on both sides: gcc compiles the file for the host (the functions
stay the registry keys), nvcc compiles the generated unit for the device.  Plain C arithmetic in
the bodies stays un-fused through -fmad=false / -Xnvvm=-fma=0, the counterpart of the reference
being built in ISO C mode.  The reference's 18 tests/scenes programs are built this way
(srp_b200/build.py: build_scenes) and reproduce the reference's framebuffers bit for bit.

Limits (reported, not guessed around): the file's shaders must be top-level definitions with the
exact parameter types above; bodies must be valid in the common subset of C and CUDA C++ (explicit
casts from void*, no designated initialisers inside shader bodies).
"""
from __future__ import annotations

import argparse
import re
import sys
from pathlib import Path

VS_SIG = re.compile(r"^\s*(?:static\s+)?void\s+(\w+)\s*\(\s*SRPVertexShaderIn\s*\*\s*(\w+)\s*,\s*SRPVertexShaderOut\s*\*\s*(\w+)\s*\)\s*$")
FS_SIG = re.compile(r"^\s*(?:static\s+)?void\s+(\w+)\s*\(\s*SRPFragmentShaderIn\s*\*\s*(\w+)\s*,\s*SRPFragmentShaderOut\s*\*\s*(\w+)\s*\)\s*$")


def strip_comments(src: str) -> str:
    """comments -> spaces (newlines kept, string literals respected)"""
    out, i, n = [], 0, len(src)
    while i < n:
        c = src[i]
        if c == '"' or c == "'":
            j = i + 1
            while j < n and src[j] != c:
                j += 2 if src[j] == "\\" else 1
            out.append(src[i:j + 1]); i = j + 1
        elif src.startswith("//", i):
            j = src.find("\n", i)
            j = n if j < 0 else j
            out.append(" " * (j - i)); i = j
        elif src.startswith("/*", i):
            j = src.find("*/", i + 2)
            j = n if j < 0 else j + 2
            out.append("".join(ch if ch == "\n" else " " for ch in src[i:j])); i = j
        else:
            out.append(c); i += 1
    return "".join(out)


def match_brace(text: str, open_at: int) -> int:
    """index just past the brace that closes text[open_at] == '{'"""
    depth, i, n = 0, open_at, len(text)
    while i < n:
        c = text[i]
        if c == '"' or c == "'":
            i += 1
            while i < n and text[i] != c:
                i += 2 if text[i] == "\\" else 1
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced braces")


def top_level_items(clean: str):
    """yields (kind, head, body_start, body_end) for top-level `head { body }` constructs;
    kind = 'typedef' (typedef struct/union ... { } NAME;), 'function', or 'other'"""
    i, n, depth_start = 0, len(clean), 0
    stmt_start = 0
    while i < n:
        c = clean[i]
        if c == "#":                       # preprocessor line (with continuations)
            j = i
            while True:
                k = clean.find("\n", j)
                if k < 0:
                    k = n; break
                if clean[k - 1] != "\\":
                    break
                j = k + 1
            i = k + 1; stmt_start = i
            continue
        if c == ";":
            stmt_start = i + 1
        elif c == "{":
            head = clean[stmt_start:i]
            end = match_brace(clean, i)
            if re.search(r"\btypedef\s+(struct|union)\b", head):
                semi = clean.find(";", end)
                yield ("typedef", head, stmt_start, semi + 1)
                i = semi + 1; stmt_start = i
                continue
            if re.search(r"\)\s*$", head) and "=" not in head:
                yield ("function", head, stmt_start, end)
                i = end; stmt_start = i
                continue
            # an initialiser / struct definition: skip to its terminating ';'
            semi = clean.find(";", end)
            i = (semi + 1) if semi >= 0 else end
            stmt_start = i
            continue
        i += 1


def generate(source: str, origin: str = "<source>") -> str:
    clean = strip_comments(source)
    typedefs, helpers, vs, fs = [], [], [], []
    for kind, head, a, b in top_level_items(clean):
        text = clean[a:b].strip("\n")
        if kind == "typedef":
            typedefs.append(text)
            continue
        one_line = " ".join(head.split())
        mv, mf = VS_SIG.match(one_line), FS_SIG.match(one_line)
        body = clean[clean.index("{", a):b]
        if mv:
            vs.append((mv.group(1), mv.group(2), mv.group(3), body))
        elif mf:
            fs.append((mf.group(1), mf.group(2), mf.group(3), body))
        elif not re.search(r"\bmain\s*\($", re.sub(r"\(.*", "(", one_line)):
            helpers.append(text)
    if not vs or not fs:
        raise ValueError(f"{origin}: no vertex / fragment shader definitions found "
                         "(expected `void f(SRPVertexShaderIn* in, SRPVertexShaderOut* out)` at file scope)")
    defines = [ln.strip() for ln in clean.splitlines()
               if re.match(r"\s*#\s*define\s+\w+\s+\S", ln) and not re.match(r"\s*#\s*define\s+SRP_INCLUDE_", ln)]
    # the program's own headers (quoted includes) may define the vertex / uniform types
    includes = [ln.strip() for ln in clean.splitlines() if re.match(r'\s*#\s*include\s+"', ln)]

    def uniform_size(in_name, body):
        m = re.search(r"\(\s*(?:const\s+)?(\w+)\s*\*\s*\)\s*" + re.escape(in_name) + r"\s*->\s*uniform", body)
        return f"sizeof({m.group(1)})" if m else "0"

    out = [f"/* GENERATED by srp_b200/twingen.py from {origin} -- the __device__ twins of its shaders;",
           " * the function bodies are the source's own text.  Do not edit: regenerate. */",
           "#include <srp_b200_device.cuh>", ""]
    out += includes + defines + ([""] if includes or defines else [])
    for t in typedefs:
        out += [t, ""]
    for h in helpers:
        out += ["__device__ " + h, ""]
    for name, a_in, a_out, body in vs:
        out += [f"__device__ void srpTwin_{name}(SRPVertexShaderIn* {a_in}, SRPVertexShaderOut* {a_out})", body, ""]
    for name, a_in, a_out, body in fs:
        out += [f"__device__ void srpTwin_{name}(SRPFragmentShaderIn* {a_in}, SRPFragmentShaderOut* {a_out})", body, ""]
    out.append("#define SRP_TWIN_VS(X) " + " ".join(f"X({i}, srpTwin_{v[0]})" for i, v in enumerate(vs)))
    out.append("#define SRP_TWIN_FS(X) " + " ".join(f"X({i}, srpTwin_{f[0]})" for i, f in enumerate(fs)))
    out += ["SRP_B200_DEFINE_SHADER_TABLES(SRP_TWIN_VS, SRP_TWIN_FS)", ""]
    for i, (name, a_in, _, body) in enumerate(vs):
        out.append(f'extern "C" void {name}(SRPVertexShaderIn*, SRPVertexShaderOut*);')
        out.append(f"SRP_B200_REGISTER_VERTEX_SHADER({name}, {i}, {uniform_size(a_in, body)})")
    for i, (name, a_in, _, body) in enumerate(fs):
        out.append(f'extern "C" void {name}(SRPFragmentShaderIn*, SRPFragmentShaderOut*);')
        out.append(f"SRP_B200_REGISTER_FRAGMENT_SHADER({name}, {i}, {uniform_size(a_in, body)})")
    return "\n".join(out) + "\n"


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("source", type=Path)
    ap.add_argument("-o", "--output", type=Path, required=True)
    args = ap.parse_args(argv)
    args.output.write_text(generate(args.source.read_text(), str(args.source)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
