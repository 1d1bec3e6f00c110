/* srp-b200 host layer -- shader program registry.
 * The API hands the library host function pointers (include/srp/api.h, SRPVertexShader
 * / SRPFragmentShader); the device cannot call those, so each executable registers,
 * per (vertex shader, fragment shader) pair, the index of the matching __device__ twins
 * in its program table and the size of its uniform struct (include/srp_b200.h). */
#include <stdlib.h>
#include "srp_internal.h"

static SRPProgramEntry* gPrograms = NULL;
static size_t gProgramCount = 0, gProgramCapacity = 0;

int srpB200RegisterProgram(SRPVertexShaderFunc hostVS, SRPFragmentShaderFunc hostFS,
                           int deviceProgramId, size_t uniformSize)
{
	for (size_t i = 0; i < gProgramCount; i++)
		if (gPrograms[i].vs == hostVS && gPrograms[i].fs == hostFS)
		{
			gPrograms[i].deviceId = deviceProgramId;
			gPrograms[i].uniformSize = uniformSize;
			return 0;
		}
	if (gProgramCount == gProgramCapacity)
	{
		gProgramCapacity = gProgramCapacity ? 2 * gProgramCapacity : 16;
		gPrograms = realloc(gPrograms, gProgramCapacity * sizeof *gPrograms);
		if (!gPrograms) abort();
	}
	gPrograms[gProgramCount++] = (SRPProgramEntry) { hostVS, hostFS, deviceProgramId, uniformSize };
	return 0;
}

const SRPProgramEntry* srpLookupProgram(SRPVertexShaderFunc vs, SRPFragmentShaderFunc fs)
{
	for (size_t i = 0; i < gProgramCount; i++)
		if (gPrograms[i].vs == vs && gPrograms[i].fs == fs)
			return &gPrograms[i];
	return NULL;
}
