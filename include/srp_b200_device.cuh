/* srp-b200 -- what a program's CUDA translation unit includes to provide the __device__
 * twins of its shaders.
 *
 *   #include <srp_b200_device.cuh>
 *   __device__ void myVS(SRPVertexShaderIn* in, SRPVertexShaderOut* out) { ... }   // same body as the C shader
 *   __device__ void myFS(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out) { ... }
 *   #define MY_PROGRAMS(X)  X(0, myVS, myFS)
 *   SRP_B200_DEFINE_PROGRAM_TABLE(MY_PROGRAMS)
 *   extern "C" void vertexShader(SRPVertexShaderIn*, SRPVertexShaderOut*);        // the host originals
 *   extern "C" void fragmentShader(SRPFragmentShaderIn*, SRPFragmentShaderOut*);
 *   SRP_B200_REGISTER_PROGRAM(vertexShader, fragmentShader, 0, sizeof(Uniform))
 *
 * The unit is compiled with
 *   nvcc -std=c++20 -gencode arch=compute_100a,code=lto_100a -rdc=true -fmad=false
 * and device-linked (-dlto -Xnvlink -Xnvvm=-fma=0) against libsrp's relocatable device
 * code, whose kernels call the two dispatchers defined by the table macro; with LTO the
 * shader bodies are inlined into the geometry and tile kernels.
 *
 * Shader bodies may use the vec / mat4 helpers (same names as on the host; under nvcc
 * they are inline and FMA-free) and srpTextureGetFilteredColor().  Plain `a*b+c` in a
 * shader relies on -fmad=false / -fma=0 to stay un-fused, exactly like the C original
 * relies on ISO C mode. */
#pragma once
#ifndef SRP_INCLUDE_VEC
	#define SRP_INCLUDE_VEC
#endif
#ifndef SRP_INCLUDE_MAT
	#define SRP_INCLUDE_MAT
#endif
#include "srp/srp.h"
#include "srp_b200.h"

/* Dispatchers the library kernels call (defined once per executable by the macro below). */
extern "C" __device__ void srpB200DeviceVS(int programId, SRPVertexShaderIn* in, SRPVertexShaderOut* out);
extern "C" __device__ void srpB200DeviceFS(int programId, SRPFragmentShaderIn* in, SRPFragmentShaderOut* out);

#define SRP_B200_VS_CASE_(id, vs, fs) case (id): vs(in, out); return;
#define SRP_B200_FS_CASE_(id, vs, fs) case (id): fs(in, out); return;
#define SRP_B200_DEFINE_PROGRAM_TABLE(TABLE) \
	extern "C" __device__ void srpB200DeviceVS(int programId, SRPVertexShaderIn* in, SRPVertexShaderOut* out) \
	{ switch (programId) { TABLE(SRP_B200_VS_CASE_) default: return; } } \
	extern "C" __device__ void srpB200DeviceFS(int programId, SRPFragmentShaderIn* in, SRPFragmentShaderOut* out) \
	{ switch (programId) { TABLE(SRP_B200_FS_CASE_) default: return; } }

/* Separate lists of vertex and fragment shaders (what srp_b200/twingen.py emits):
 *   #define MY_VS(X) X(0, vsA) X(1, vsB)
 *   #define MY_FS(X) X(0, fsA)
 *   SRP_B200_DEFINE_SHADER_TABLES(MY_VS, MY_FS) */
#define SRP_B200_SHADER_CASE_(id, fn) case (id): fn(in, out); return;
#define SRP_B200_DEFINE_SHADER_TABLES(VS_TABLE, FS_TABLE) \
	extern "C" __device__ void srpB200DeviceVS(int programId, SRPVertexShaderIn* in, SRPVertexShaderOut* out) \
	{ switch (programId) { VS_TABLE(SRP_B200_SHADER_CASE_) default: return; } } \
	extern "C" __device__ void srpB200DeviceFS(int programId, SRPFragmentShaderIn* in, SRPFragmentShaderOut* out) \
	{ switch (programId) { FS_TABLE(SRP_B200_SHADER_CASE_) default: return; } }

#define SRP_B200_CAT2_(a, b) a##b
#define SRP_B200_CAT_(a, b) SRP_B200_CAT2_(a, b)
#define SRP_B200_REGISTER_PROGRAM(hostVS, hostFS, id, uniformSize) \
	static const int SRP_B200_CAT_(srpB200Registered_, __LINE__) = \
		srpB200RegisterProgram((hostVS), (hostFS), (id), (uniformSize));
#define SRP_B200_REGISTER_VERTEX_SHADER(hostVS, id, uniformSize) \
	static const int SRP_B200_CAT_(srpB200RegisteredVS_, __LINE__) = \
		srpB200RegisterVertexShader((hostVS), (id), (uniformSize));
#define SRP_B200_REGISTER_FRAGMENT_SHADER(hostFS, id, uniformSize) \
	static const int SRP_B200_CAT_(srpB200RegisteredFS_, __LINE__) = \
		srpB200RegisterFragmentShader((hostFS), (id), (uniformSize));
