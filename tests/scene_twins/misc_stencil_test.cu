#include "twin_common.cuh"
/* pass 1 = (vertexShader, fragmentShader) textured; pass 2 = (vertexShader, singleColor) outline */
extern "C" void singleColor(SRPFragmentShaderIn*, SRPFragmentShaderOut*);
__device__ void vs(SRPVertexShaderIn* in, SRPVertexShaderOut* out) { twin::transformMvp<twin::FrameMvpTex>(in, out); twin::copyUv(in, out); }
__device__ void fsTex(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out) { twin::fsTexture<twin::FrameMvpTex>(in, out); }
__device__ void fsRed(SRPFragmentShaderIn*, SRPFragmentShaderOut* out) { *(vec4*) out->color = VEC4(1, 0, 0, 1); }
#define PROGRAMS(X) X(0, vs, fsTex) X(1, vs, fsRed)
SRP_B200_DEFINE_PROGRAM_TABLE(PROGRAMS)
SRP_B200_REGISTER_PROGRAM(vertexShader, fragmentShader, 0, sizeof(twin::FrameMvpTex))
SRP_B200_REGISTER_PROGRAM(vertexShader, singleColor, 1, sizeof(twin::FrameMvpTex))
