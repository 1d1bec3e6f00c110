/* srp-b200 built-in programs -- shader bodies, written ONCE in the common subset of C and
 * CUDA C++ and compiled twice: by gcc as the host functions (which the reference library
 * calls directly when it serves as oracle / CPU baseline, and which serve as registry keys
 * for this library), and by nvcc as the __device__ twins the kernels run.  Same source,
 * same operation order, no FMA contraction on either side => identical results.
 *
 * SRPB_FN / SRPB_NAME are provided by the including file. */

/* The vec types are packed to alignment 1 (they have to match the host layouts), which
 * makes the device compiler fetch them byte by byte.  All built-in vertex / varyings
 * layouts are 4-byte aligned, so the shaders go through float pointers instead. */
SRPB_FN static vec3 SRPB_NAME(srpbLoad3)(const void* p)
{
	const float* f = (const float*) p;
	return VEC3(f[0], f[1], f[2]);
}
SRPB_FN static void SRPB_NAME(srpbStore3)(void* p, vec3 v)
{
	float* f = (float*) p;
	f[0] = v.x; f[1] = v.y; f[2] = v.z;
}
SRPB_FN static void SRPB_NAME(srpbStoreClip)(SRPVertexShaderOut* out, vec4 p)
{
	out->clipPosition[0] = p.x; out->clipPosition[1] = p.y; out->clipPosition[2] = p.z; out->clipPosition[3] = p.w;
}

/* transform a model-space position to clip space: projection * (view * (model * p)) */
SRPB_FN static vec4 SRPB_NAME(srpbToClip)(const SrpbTransform* xf, vec3 position)
{
	vec4 p = VEC4_FROM_VEC3(position, 1.f);
	p = mat4MultiplyVec4(&xf->model, p);
	p = mat4MultiplyVec4(&xf->view, p);
	p = mat4MultiplyVec4(&xf->projection, p);
	return p;
}

/* ---- texcube: cfg1 (geometry/state of reference examples/03_textured_cube.c) ---- */
SRPB_FN void SRPB_NAME(srpb_texcube_vs)(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	const float* v = (const float*) in->vertex;                 /* SrpbTexVertex  */
	const SrpbTexCubeUniform* u = (const SrpbTexCubeUniform*) in->uniform;
	float* o = (float*) out->varyings;                          /* SrpbUvVaryings */
	SRPB_NAME(srpbStoreClip)(out, SRPB_NAME(srpbToClip)(&u->xf, SRPB_NAME(srpbLoad3)(v)));
	o[0] = v[3];
	o[1] = v[4];
}
SRPB_FN void SRPB_NAME(srpb_texcube_fs)(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const float* uv = (const float*) in->varyings;
	const SrpbTexCubeUniform* u = (const SrpbTexCubeUniform*) in->uniform;
	srpTextureGetFilteredColor(u->texture, uv[0], uv[1], out->color);
}

/* ---- gouraud: cfg2 / cfg3 / cfg5 -- ambient + diffuse lighting in the vertex shader,
 * one perspective-correct vec3 colour varying ---- */
SRPB_FN void SRPB_NAME(srpb_gouraud_vs)(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	const float* v = (const float*) in->vertex;                 /* SrpbMeshVertex: position 0..2, uv 3..4, normal 5..7 */
	const SrpbGouraudUniform* u = (const SrpbGouraudUniform*) in->uniform;
	SRPB_NAME(srpbStoreClip)(out, SRPB_NAME(srpbToClip)(&u->xf, SRPB_NAME(srpbLoad3)(v)));

	vec4 n4 = mat4MultiplyVec4(&u->xf.model, VEC4_FROM_VEC3(SRPB_NAME(srpbLoad3)(v + 5), 0.f));
	vec3 n = vec3Normalize(VEC3(n4.x, n4.y, n4.z));
	float diff = fmaxf(vec3DotProduct(n, vec3Negate(u->lightDirection)), 0.f);
	vec3 ambient = vec3MultiplyVec3(u->lightAmbient, u->materialAmbient);
	vec3 diffuse = vec3MultiplyVec3(u->lightDiffuse, vec3MultiplyScalar(u->materialDiffuse, diff));
	SRPB_NAME(srpbStore3)(out->varyings, vec3Add(ambient, diffuse));
}
SRPB_FN void SRPB_NAME(srpb_gouraud_fs)(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const float* color = (const float*) in->varyings;           /* SrpbColorVaryings */
	out->color[0] = color[0];
	out->color[1] = color[1];
	out->color[2] = color[2];
	out->color[3] = 1.f;
}

/* ---- vcolor: per-vertex colour passed through (interpolation-mode tests) ---- */
SRPB_FN void SRPB_NAME(srpb_vcolor_vs)(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	const float* v = (const float*) in->vertex;                 /* SrpbColorVertex */
	const SrpbTransform* u = (const SrpbTransform*) in->uniform;
	SRPB_NAME(srpbStoreClip)(out, SRPB_NAME(srpbToClip)(u, SRPB_NAME(srpbLoad3)(v)));
	SRPB_NAME(srpbStore3)(out->varyings, SRPB_NAME(srpbLoad3)(v + 3));
}

/* ---- primid: no varyings, colour derived from the primitive id so that any ordering or
 * id-assignment error is visible (used by points / lines / polygon-mode tests) ---- */
SRPB_FN void SRPB_NAME(srpb_primid_vs)(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	/* any 4-byte aligned vertex format that starts with a vec3 */
	const SrpbTransform* u = (const SrpbTransform*) in->uniform;
	SRPB_NAME(srpbStoreClip)(out, SRPB_NAME(srpbToClip)(u, SRPB_NAME(srpbLoad3)(in->vertex)));
}
SRPB_FN void SRPB_NAME(srpb_primid_fs)(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const unsigned id = (unsigned) in->primitiveID;
	out->color[0] = (float) (1u + (id * 37u) % 254u) / 255.f;
	out->color[1] = (float) (1u + (id * 101u) % 254u) / 255.f;
	out->color[2] = (float) (1u + (id / 254u) % 254u) / 255.f;
	out->color[3] = in->frontFacing ? 1.f : 0.5f;
}

/* ---- tagged: cfg4 -- positions are given directly in NDC (w = 1), a FLAT uint8 varying
 * carries the vertex tag; the colour mixes tag and primitive id ---- */
SRPB_FN void SRPB_NAME(srpb_tagged_vs)(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	const float* v = (const float*) in->vertex;                 /* SrpbTagVertex */
	SrpbTagVaryings* o = (SrpbTagVaryings*) out->varyings;
	out->clipPosition[0] = v[0];
	out->clipPosition[1] = v[1];
	out->clipPosition[2] = v[2];
	out->clipPosition[3] = 1.f;
	o->tag = (uint8_t) (((const uint32_t*) in->vertex)[3] & 0xFFu);
}
SRPB_FN void SRPB_NAME(srpb_tagged_fs)(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const SrpbTagVaryings* v = (const SrpbTagVaryings*) in->varyings;
	const unsigned id = (unsigned) in->primitiveID;
	out->color[0] = (float) v->tag / 255.f;
	out->color[1] = (float) (1u + (id * 101u) % 254u) / 255.f;
	out->color[2] = (float) (1u + (id / 254u) % 254u) / 255.f;
	out->color[3] = 1.f;
}

/* ---- solid: constant colour from the uniform ---- */
SRPB_FN void SRPB_NAME(srpb_solid_vs)(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	const SrpbSolidUniform* u = (const SrpbSolidUniform*) in->uniform;
	SRPB_NAME(srpbStoreClip)(out, SRPB_NAME(srpbToClip)(&u->xf, SRPB_NAME(srpbLoad3)(in->vertex)));
}
SRPB_FN void SRPB_NAME(srpb_solid_fs)(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const SrpbSolidUniform* u = (const SrpbSolidUniform*) in->uniform;
	const float* c = (const float*) &u->color;
	out->color[0] = c[0]; out->color[1] = c[1]; out->color[2] = c[2]; out->color[3] = c[3];
}

/* ---- depthout: fragment shader that replaces the depth on every other pixel column
 * (mayOverwriteDepth = true => the late depth-test path, reference fragment.c:100-111) ---- */
SRPB_FN void SRPB_NAME(srpb_depthout_fs)(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const float* color = (const float*) in->varyings;
	out->color[0] = color[0]; out->color[1] = color[1]; out->color[2] = color[2]; out->color[3] = 1.f;
	if (((int) in->fragCoord[0]) % 2 == 0)
		out->fragDepth = in->fragCoord[2] * 0.5f;
}

/* ---- mixed: every varying type class in one blob (double, float, int32, uint16) ---- */
SRPB_FN void SRPB_NAME(srpb_mixed_vs)(SRPVertexShaderIn* in, SRPVertexShaderOut* out)
{
	const float* v = (const float*) in->vertex;                 /* SrpbColorVertex */
	const SrpbTransform* u = (const SrpbTransform*) in->uniform;
	SrpbMixedVaryings* o = (SrpbMixedVaryings*) out->varyings;
	SRPB_NAME(srpbStoreClip)(out, SRPB_NAME(srpbToClip)(u, SRPB_NAME(srpbLoad3)(v)));
	o->height = (double) v[1] * 0.333333333333;
	o->color = SRPB_NAME(srpbLoad3)(v + 3);
	o->uv = VEC2(v[0], v[2]);
	o->tag = (int32_t) in->vertexID * 7 - 3;
	o->pair[0] = (uint16_t) (in->vertexID & 0xFFFFu);
	o->pair[1] = (uint16_t) (1000u + in->vertexID);
}
SRPB_FN void SRPB_NAME(srpb_mixed_fs)(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out)
{
	const SrpbMixedVaryings* v = (const SrpbMixedVaryings*) in->varyings;
	out->color[0] = v->color.x * 0.5f + (float) (v->height * 0.25 + 0.25);
	out->color[1] = v->color.y * 0.5f + v->uv.x * 0.125f + 0.25f;
	out->color[2] = (float) ((unsigned) (v->tag + 3) % 251u) / 255.f;
	out->color[3] = (float) ((v->pair[0] + v->pair[1]) % 256u) / 255.f;
}
