#include "twin_common.cuh"
/* two programs share the vertex shader: sp1 = (vertexShader, fsPrimitive), sp2 = (vertexShader, fsWhite) */
extern "C" void fsPrimitive(SRPFragmentShaderIn*, SRPFragmentShaderOut*);
extern "C" void fsWhite(SRPFragmentShaderIn*, SRPFragmentShaderOut*);
__device__ void vs(SRPVertexShaderIn* in, SRPVertexShaderOut* out) { twin::transformMvp<twin::FrameMvp>(in, out); }
__device__ void fsId(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out) { twin::fsPrimitiveId<true>(in, out); }
__device__ void fsW(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out) { twin::fsWhite(in, out); }
#define PROGRAMS(X) X(0, vs, fsId) X(1, vs, fsW)
SRP_B200_DEFINE_PROGRAM_TABLE(PROGRAMS)
SRP_B200_REGISTER_PROGRAM(vertexShader, fsPrimitive, 0, sizeof(twin::FrameMvp))
SRP_B200_REGISTER_PROGRAM(vertexShader, fsWhite, 1, sizeof(twin::FrameMvp))
