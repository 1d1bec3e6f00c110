#include "twin_common.cuh"
/* a FLAT varying declared as 3 floats over a 1-byte struct: only byte 0 is meaningful */
struct Vertex { vec3 position; uint8_t color; };
struct Uniform { mat4 model; };
__device__ void vs(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	const Vertex* v = (const Vertex*) in->vertex;
	const Uniform* u = (const Uniform*) in->uniform;
	*(vec4*) out->clipPosition = mat4MultiplyVec4(&u->model, VEC4_FROM_VEC3(v->position, 1.));
	*(uint8_t*) out->varyings = v->color;
}
__device__ void fs(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const uint8_t c = *(const uint8_t*) in->varyings;
	vec4* color = (vec4*) out->color;
	if (c == 0) *color = VEC4(1, 0, 0, 1);
	else if (c == 1) *color = VEC4(0, 1, 0, 1);
	else if (c == 2) *color = VEC4(0, 0, 1, 1);
}
#define PROGRAMS(X) X(0, vs, fs)
SRP_B200_DEFINE_PROGRAM_TABLE(PROGRAMS)
SRP_B200_REGISTER_PROGRAM(vertexShader, fragmentShader, 0, sizeof(Uniform))
