/* srp-b200 built-in programs -- device twins and their registration.
 * Compiles builtin_shaders.h a second time, as __device__ code, defines the program table
 * the library kernels dispatch through, and registers every (host VS, host FS) pair of
 * builtin_host.c with its table index and uniform size. */
#include <srp_b200_device.cuh>
#include "builtin_uniforms.h"

#define SRPB_FN __device__
#define SRPB_NAME(n) n##_dev
#include "builtin_shaders.h"
#include "builtin_table.h"

#define SRPB_VS_CASE(id, name, vs, fs, U) case id: vs##_dev(in, out); return;
#define SRPB_FS_CASE(id, name, vs, fs, U) case id: fs##_dev(in, out); return;
extern "C" __device__ void srpB200DeviceVS(int programId, SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	switch (programId) { SRPB_PROGRAM_TABLE(SRPB_VS_CASE) default: return; }
}
extern "C" __device__ void srpB200DeviceFS(int programId, SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	switch (programId) { SRPB_PROGRAM_TABLE(SRPB_FS_CASE) default: return; }
}

/* host originals (builtin_host.c) */
#define SRPB_DECLARE_HOST(id, name, vs, fs, U) \
	extern "C" void vs(SRPVertexShaderIn*, SRPVertexShaderOut*); \
	extern "C" void fs(SRPFragmentShaderIn*, SRPFragmentShaderOut*);
SRPB_PROGRAM_TABLE(SRPB_DECLARE_HOST)

namespace {
struct SrpbRegistrar
{
	SrpbRegistrar()
	{
		#define SRPB_REGISTER_CALL(id, name, vs, fs, U) srpB200RegisterProgram(vs, fs, id, sizeof(U));
		SRPB_PROGRAM_TABLE(SRPB_REGISTER_CALL)
	}
};
SrpbRegistrar gSrpbRegistrar;
}
