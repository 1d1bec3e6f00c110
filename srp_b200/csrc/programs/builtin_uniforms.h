/* srp-b200 built-in programs -- vertex, varyings and uniform layouts (plain C, shared by
 * the host C build, the CUDA build and mirrored with ctypes in srp_b200/programs.py).
 * These are the shader programs of the BASELINE configs plus a few
 * that exercise state the reference's scenes do not pin. */
#ifndef SRPB_BUILTIN_UNIFORMS_H_
#define SRPB_BUILTIN_UNIFORMS_H_
#ifndef SRP_INCLUDE_VEC
	#define SRP_INCLUDE_VEC
#endif
#ifndef SRP_INCLUDE_MAT
	#define SRP_INCLUDE_MAT
#endif
#include "srp/srp.h"

/* ---- vertex formats ---- */
typedef struct SrpbTexVertex { vec3 position; vec2 uv; } SrpbTexVertex;                   /* 20 B, cfg1 cube   */
typedef struct SrpbMeshVertex { vec3 position; vec2 uv; vec3 normal; } SrpbMeshVertex;    /* 32 B, OBJ meshes  */
typedef struct SrpbTagVertex { vec3 position; uint32_t tag; } SrpbTagVertex;              /* 16 B, cfg4        */
typedef struct SrpbColorVertex { vec3 position; vec3 color; } SrpbColorVertex;            /* 24 B              */

/* ---- uniforms ---- */
typedef struct SrpbTransform { mat4 model, view, projection; } SrpbTransform;

typedef struct SrpbTexCubeUniform
{
	SrpbTransform xf;
	SRPTexture* texture;
} SrpbTexCubeUniform;

typedef struct SrpbGouraudUniform
{
	SrpbTransform xf;
	vec3 materialAmbient, materialDiffuse;
	vec3 lightAmbient, lightDiffuse, lightDirection;
} SrpbGouraudUniform;

typedef struct SrpbSolidUniform
{
	SrpbTransform xf;
	vec4 color;
} SrpbSolidUniform;

/* ---- varyings ---- */
typedef struct SrpbUvVaryings { vec2 uv; } SrpbUvVaryings;
typedef struct SrpbColorVaryings { vec3 color; } SrpbColorVaryings;
typedef struct SrpbTagVaryings { uint8_t tag; } SrpbTagVaryings;
typedef struct SrpbMixedVaryings
{
	double height;        /* SRP_DOUBLE x1 */
	vec3 color;           /* SRP_FLOAT  x3 */
	vec2 uv;              /* SRP_FLOAT  x2 */
	int32_t tag;          /* SRP_INT32  x1 */
	uint16_t pair[2];     /* SRP_UINT16 x2 */
} SrpbMixedVaryings;

#endif
