/* Device twins for the reference's own scene programs (tests/scenes of kitrofimov/srp).
 *
 * The scene .c files are compiled UNMODIFIED from /root/reference and relinked against
 * libsrp.a ("existing programs relink unchanged", BASELINE.json north_star); the only new
 * translation unit per program is one of the small .cu files in this directory, which
 * supplies __device__ twins of that scene's shaders and registers them against the host
 * function symbols the scene defines (vertexShader, fragmentShader, ...).
 *
 * The 18 scenes use a handful of shader shapes; the building blocks are here.  Arithmetic
 * mirrors the C originals: plain float/double expressions (kept un-fused by -fmad=false /
 * -Xnvvm=-fma=0, like ISO C keeps the originals un-fused) and the vec/mat helpers. */
#pragma once
#include <srp_b200_device.cuh>

namespace twin {

struct Mvp { mat4 model, view, projection; };
struct FrameMvp { size_t frameCount; mat4 model, view, projection; };        /* most teapot / cube scenes */
struct FrameMvpTex { size_t frameCount; mat4 model, view, projection; SRPTexture* texture; };

struct ColorVertex { vec3 position; vec3 color; };
struct UvVertex { vec3 position; vec2 uv; };
struct ObjVertex { vec3 position; vec2 uv; vec3 normal; };                  /* examples/utility/objparser.h:6-10 */

/* out = (position, 1) */
__device__ inline void positionOnly(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	const vec3* p = (const vec3*) in->vertex;
	*(vec4*) out->clipPosition = VEC4_FROM_VEC3(*p, 1.);
}

/* out = projection * (view * (model * (position, 1))) for a uniform that holds the three
 * matrices at member pointers model/view/projection */
template <typename U>
__device__ inline void transformMvp(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	const U* u = (const U*) in->uniform;
	const vec3* p = (const vec3*) in->vertex;
	vec4 v = VEC4_FROM_VEC3(*p, 1.);
	v = mat4MultiplyVec4(&u->model, v);
	v = mat4MultiplyVec4(&u->view, v);
	v = mat4MultiplyVec4(&u->projection, v);
	*(vec4*) out->clipPosition = v;
}

__device__ inline void copyColor(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	*(vec3*) out->varyings = ((const ColorVertex*) in->vertex)->color;
}
__device__ inline void copyUv(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	*(vec2*) out->varyings = ((const UvVertex*) in->vertex)->uv;
}

/* colour = varying vec3, alpha 1 */
__device__ inline void fsVaryingColor(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const vec3* c = (const vec3*) in->varyings;
	out->color[0] = c->x; out->color[1] = c->y; out->color[2] = c->z; out->color[3] = 1.;
}
__device__ inline void fsWhite(SRPFragmentShaderIn*, SRPFragmentShaderOut* out)
{
	out->color[0] = 1.; out->color[1] = 1.; out->color[2] = 1.; out->color[3] = 1.;
}
/* colour from the primitive id: ((id*k) % 255) / 255. [/ 2.] evaluated in double */
template <bool HALVED>
__device__ inline void fsPrimitiveId(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const int id = (int) in->primitiveID;
	const int r = (id * 97) % 255, g = (id * 57) % 255, b = (id * 23) % 255;
	if (HALVED)
	{
		out->color[0] = r / 255. / 2.; out->color[1] = g / 255. / 2.; out->color[2] = b / 255. / 2.;
	}
	else
	{
		out->color[0] = r / 255.; out->color[1] = g / 255.; out->color[2] = b / 255.;
	}
	out->color[3] = 1.;
}
template <typename U>
__device__ inline void fsTexture(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const vec2 uv = *(const vec2*) in->varyings;
	srpTextureGetFilteredColor(((const U*) in->uniform)->texture, uv.x, uv.y, out->color);
}

} // namespace twin

/* the host originals every scene defines */
extern "C" void vertexShader(SRPVertexShaderIn*, SRPVertexShaderOut*);
extern "C" void fragmentShader(SRPFragmentShaderIn*, SRPFragmentShaderOut*);
