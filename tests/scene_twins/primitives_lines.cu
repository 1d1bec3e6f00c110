#include "twin_common.cuh"
__device__ void vs(SRPVertexShaderIn* in, SRPVertexShaderOut* out) { twin::transformMvp<twin::Mvp>(in, out); }
__device__ void fs(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out) { twin::fsWhite(in, out); }
#define PROGRAMS(X) X(0, vs, fs)
SRP_B200_DEFINE_PROGRAM_TABLE(PROGRAMS)
SRP_B200_REGISTER_PROGRAM(vertexShader, fragmentShader, 0, sizeof(twin::Mvp))
