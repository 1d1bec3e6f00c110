/* srp-b200 host layer -- shader registry.
 * The API hands the library host function pointers (include/srp/api.h, SRPVertexShader
 * / SRPFragmentShader); the device cannot call those, so each executable registers, per host
 * shader function, the index of the matching __device__ twin in its shader tables and the size
 * of the uniform struct that shader reads (include/srp_b200.h).  Vertex and fragment shaders are
 * registered independently -- programs recombine them freely at run time (reference
 * tests/scenes/clipping/point.c:62-66 swaps the fragment shader of a copied program) -- and
 * srpB200RegisterProgram is the pairwise shorthand. */
#include <stdlib.h>
#include "srp_internal.h"

typedef struct ShaderEntry
{
	void (*fn)(void);
	int deviceId;
	size_t uniformSize;
} ShaderEntry;

typedef struct ShaderTable
{
	ShaderEntry* e;
	size_t count, capacity;
} ShaderTable;

static ShaderTable gVS, gFS;

/* pairs registered as pairs keep their own uniform size (a shared vertex shader may serve
 * programs with different uniform structs) */
static SRPProgramEntry* gPairs = NULL;
static size_t gPairCount = 0, gPairCapacity = 0;

static int registerShader(ShaderTable* t, void (*fn)(void), int deviceId, size_t uniformSize)
{
	for (size_t i = 0; i < t->count; i++)
		if (t->e[i].fn == fn)
		{
			t->e[i].deviceId = deviceId;
			t->e[i].uniformSize = uniformSize;
			return 0;
		}
	if (t->count == t->capacity)
	{
		t->capacity = t->capacity ? 2 * t->capacity : 16;
		t->e = realloc(t->e, t->capacity * sizeof *t->e);
		if (!t->e) abort();
	}
	t->e[t->count++] = (ShaderEntry) { fn, deviceId, uniformSize };
	return 0;
}

static const ShaderEntry* findShader(const ShaderTable* t, void (*fn)(void))
{
	for (size_t i = 0; i < t->count; i++)
		if (t->e[i].fn == fn)
			return &t->e[i];
	return NULL;
}

int srpB200RegisterVertexShader(SRPVertexShaderFunc hostVS, int deviceShaderId, size_t uniformSize)
{
	return registerShader(&gVS, (void (*)(void)) hostVS, deviceShaderId, uniformSize);
}

int srpB200RegisterFragmentShader(SRPFragmentShaderFunc hostFS, int deviceShaderId, size_t uniformSize)
{
	return registerShader(&gFS, (void (*)(void)) hostFS, deviceShaderId, uniformSize);
}

int srpB200RegisterProgram(SRPVertexShaderFunc hostVS, SRPFragmentShaderFunc hostFS,
                           int deviceProgramId, size_t uniformSize)
{
	for (size_t i = 0; i < gPairCount; i++)
		if (gPairs[i].vs == hostVS && gPairs[i].fs == hostFS)
		{
			gPairs[i].vsDeviceId = gPairs[i].fsDeviceId = deviceProgramId;
			gPairs[i].uniformSize = uniformSize;
			return 0;
		}
	if (gPairCount == gPairCapacity)
	{
		gPairCapacity = gPairCapacity ? 2 * gPairCapacity : 16;
		gPairs = realloc(gPairs, gPairCapacity * sizeof *gPairs);
		if (!gPairs) abort();
	}
	gPairs[gPairCount++] = (SRPProgramEntry) { hostVS, hostFS, deviceProgramId, deviceProgramId, uniformSize };
	return 0;
}

bool srpLookupProgram(SRPVertexShaderFunc vs, SRPFragmentShaderFunc fs, SRPProgramEntry* out)
{
	for (size_t i = 0; i < gPairCount; i++)
		if (gPairs[i].vs == vs && gPairs[i].fs == fs)
		{
			*out = gPairs[i];
			return true;
		}
	const ShaderEntry* v = findShader(&gVS, (void (*)(void)) vs);
	const ShaderEntry* f = findShader(&gFS, (void (*)(void)) fs);
	if (v == NULL || f == NULL)
		return false;
	out->vs = vs; out->fs = fs;
	out->vsDeviceId = v->deviceId;
	out->fsDeviceId = f->deviceId;
	out->uniformSize = v->uniformSize > f->uniformSize ? v->uniformSize : f->uniformSize;
	return true;
}
