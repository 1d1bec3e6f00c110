"""CPU tests of the shader toolchain (srp_b200/twingen.py): one C source, both sides."""
from pathlib import Path

import pytest

from srp_b200 import twingen

SRC = r'''
#define SRP_INCLUDE_VEC
#include <srp/srp.h>
#include "objparser.h"
#define SCALE 2.5f   /* kept */
typedef struct Uniform { mat4 mvp; /* } not a brace */ float k; } Uniform;
typedef struct VSOutput
{
    vec3 color;
} VSOutput;
SRPContext srpContext;
void vertexShader(SRPVertexShaderIn* in, SRPVertexShaderOut* out);
static float helper(float x) { return x * SCALE; }
int main(void) { Uniform u = { .k = 1.f }; (void) u; return 0; }   // not copied
void vertexShader(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
    Uniform* u = (Uniform*) in->uniform;            // the uniform type of this shader
    ((VSOutput*) out->varyings)->color.x = helper(u->k);
    char c = '}'; (void) c;
}
void fsA(SRPFragmentShaderIn* fin, SRPFragmentShaderOut* fout) { fout->color[0] = 1.f; }
void fsB(SRPFragmentShaderIn* in,
         SRPFragmentShaderOut* out)
{
    const Uniform* u = (const Uniform*) in->uniform;
    out->color[1] = u->k;
}
'''


def test_generates_twins_tables_and_registrations():
    cu = twingen.generate(SRC, "app.c")
    assert '#include "objparser.h"' in cu and "#define SCALE 2.5f" in cu and "SRP_INCLUDE_VEC" not in cu
    assert "typedef struct Uniform { mat4 mvp;" in cu and "typedef struct VSOutput" in cu
    assert "__device__ static float helper(float x) { return x * SCALE; }" in cu
    assert "int main" not in cu and "SRPContext srpContext" not in cu
    assert "__device__ void srpTwin_vertexShader(SRPVertexShaderIn* in, SRPVertexShaderOut* out)" in cu
    assert "__device__ void srpTwin_fsA(SRPFragmentShaderIn* fin, SRPFragmentShaderOut* fout)" in cu
    assert "#define SRP_TWIN_VS(X) X(0, srpTwin_vertexShader)" in cu
    assert "#define SRP_TWIN_FS(X) X(0, srpTwin_fsA) X(1, srpTwin_fsB)" in cu
    assert "SRP_B200_REGISTER_VERTEX_SHADER(vertexShader, 0, sizeof(Uniform))" in cu
    assert "SRP_B200_REGISTER_FRAGMENT_SHADER(fsA, 0, 0)" in cu
    assert "SRP_B200_REGISTER_FRAGMENT_SHADER(fsB, 1, sizeof(Uniform))" in cu
    assert cu.count("helper(u->k)") == 1      # the body is the source's own text, once


def test_source_without_shaders_is_an_error():
    with pytest.raises(ValueError):
        twingen.generate("int main(void) { return 0; }", "x.c")


def test_every_reference_scene_generates():
    scenes = Path("/root/reference/tests/scenes")
    if not scenes.is_dir():
        pytest.skip("the reference is only present in the build container")
    files = sorted(scenes.glob("*/*.c"))
    assert len(files) == 18
    for f in files:
        cu = twingen.generate(f.read_text(), str(f))
        assert "SRP_B200_DEFINE_SHADER_TABLES" in cu and "srpTwin_vertexShader" in cu
