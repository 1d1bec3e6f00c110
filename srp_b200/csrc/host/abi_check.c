/* srp-b200 host layer -- compile-time check that the public structs keep the x86-64
 * layouts programs were compiled against (measured on the reference). */
#include <stddef.h>
#include "srp/api.h"
#define SRP_INCLUDE_VEC
#include "srp/vec.h"
#include "srp/mat.h"

_Static_assert(sizeof(SRPVertexShaderIn) == 24, "SRPVertexShaderIn");
_Static_assert(sizeof(SRPVertexShaderOut) == 24, "SRPVertexShaderOut");
_Static_assert(sizeof(SRPFragmentShaderIn) == 48, "SRPFragmentShaderIn");
_Static_assert(offsetof(SRPFragmentShaderIn, fragCoord) == 16, "fragCoord");
_Static_assert(offsetof(SRPFragmentShaderIn, frontFacing) == 32, "frontFacing");
_Static_assert(offsetof(SRPFragmentShaderIn, primitiveID) == 40, "primitiveID");
_Static_assert(sizeof(SRPFragmentShaderOut) == 20, "SRPFragmentShaderOut");
_Static_assert(sizeof(SRPVertexShader) == 32, "SRPVertexShader");
_Static_assert(sizeof(SRPFragmentShader) == 16, "SRPFragmentShader");
_Static_assert(sizeof(SRPShaderProgram) == 24, "SRPShaderProgram");
_Static_assert(sizeof(SRPVaryingInfo) == 16, "SRPVaryingInfo");
_Static_assert(sizeof(SRPFramebuffer) == 48, "SRPFramebuffer");
_Static_assert(sizeof(SRPContext) == 144, "SRPContext");
_Static_assert(offsetof(SRPContext, raster) == 20, "raster");
_Static_assert(offsetof(SRPContext, scissor) == 40, "scissor");
_Static_assert(offsetof(SRPContext, stencil) == 80, "stencil");
_Static_assert(offsetof(SRPContext, depth) == 124, "depth");
_Static_assert(offsetof(SRPContext, arena) == 136, "arena");
_Static_assert(sizeof(vec2) == 8 && sizeof(vec3) == 12 && sizeof(vec4) == 16, "vec sizes");
_Static_assert(_Alignof(vec3) == 1, "vec types are packed to alignment 1");
_Static_assert(sizeof(mat4) == 64, "mat4");

int srpAbiCheck(void) { return 1; }
