#include "twin_common.cuh"
struct Uniform { mat4 rotation; };
__device__ void vs(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	const Uniform* u = (const Uniform*) in->uniform;
	const vec3* p = (const vec3*) in->vertex;
	*(vec4*) out->clipPosition = mat4MultiplyVec4(&u->rotation, VEC4_FROM_VEC3(*p, 1.));
	twin::copyColor(in, out);
}
__device__ void fs(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out) { twin::fsVaryingColor(in, out); }
#define PROGRAMS(X) X(0, vs, fs)
SRP_B200_DEFINE_PROGRAM_TABLE(PROGRAMS)
SRP_B200_REGISTER_PROGRAM(vertexShader, fragmentShader, 0, sizeof(Uniform))
