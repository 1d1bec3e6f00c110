/* srp-b200 built-in programs -- host side (plain C).
 *
 * Built twice from this one source:
 *   - into libsrp_b200.so, where the functions are the registry keys of the device twins
 *     (builtin_device.cu) and never run;
 *   - with -DSRPB_REFERENCE_BUILD into oracle/_ref/libref_host.so together with the
 *     unmodified reference library, where they ARE the shaders the reference executes
 *     (oracle and CPU baseline).
 * Both libraries therefore export the identical C ABI: the srp API, `srpContext`, and
 * srpbFindProgram() -- the Python host mirror drives either one with the same code. */
#include <math.h>
#include <string.h>
#include "builtin_uniforms.h"

#define SRPB_FN
#define SRPB_NAME(n) n
#include "builtin_shaders.h"
#include "builtin_table.h"

/* the context object a user program would define (include/srp/api.h) */
SRPContext srpContext;

typedef struct SrpbProgramInfo
{
	const char* name;
	void (*vs)(SRPVertexShaderIn*, SRPVertexShaderOut*);
	void (*fs)(SRPFragmentShaderIn*, SRPFragmentShaderOut*);
	size_t uniformSize;
	int deviceId;
} SrpbProgramInfo;

#define SRPB_INFO_ROW(id, name, vs, fs, U) { #name, vs, fs, sizeof(U), id },
static const SrpbProgramInfo gPrograms[] = { SRPB_PROGRAM_TABLE(SRPB_INFO_ROW) };

const SrpbProgramInfo* srpbFindProgram(const char* name)
{
	for (size_t i = 0; i < sizeof gPrograms / sizeof gPrograms[0]; i++)
		if (strcmp(gPrograms[i].name, name) == 0)
			return &gPrograms[i];
	return NULL;
}

/* 1 when this library is the reference build (oracle / CPU baseline), 0 for the product */
int srpbIsReferenceBuild(void)
{
#ifdef SRPB_REFERENCE_BUILD
	return 1;
#else
	return 0;
#endif
}

/* fragment-shader invocation counter for the reference arm (the product counts on the
 * device, include/srp_b200.h SRPB200Stats).  Not used by the shaders above; the bench
 * measures shaded fragments of the reference through a counting wrapper instead. */
