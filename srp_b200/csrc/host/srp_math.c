/* srp-b200 host layer -- vec2/3/4 and mat4 library functions for C programs.
 * Compiled as ISO C (no FP contraction), operation order as in reference
 * src/math/vec.c:16-189 and src/math/mat.c:16-215: sums of products associate to the
 * left with every product rounded, Normalize multiplies by 1/length (zero stays zero),
 * Reflect is i - n*(2*dot(n,i)), rotation entries are evaluated in double from
 * double sin/cos and rounded once. */
#include <math.h>
#define SRP_INCLUDE_VEC
#define SRP_INCLUDE_MAT
#include "srp/srp.h"

#define LANES2(OP) OP(x) OP(y)
#define LANES3(OP) OP(x) OP(y) OP(z)
#define LANES4(OP) OP(x) OP(y) OP(z) OP(w)
#define L_ADD(l) r.l = a.l + b.l;
#define L_SUB(l) r.l = a.l - b.l;
#define L_MUL(l) r.l = a.l * b.l;
#define L_SCALE(l) r.l = a.l * b;
#define L_NEG(l) r.l = -a.l;
#define L_ZERO(l) r.l = 0;

#define DOT2(a, b) ((a).x * (b).x + (a).y * (b).y)
#define DOT3(a, b) ((a).x * (b).x + (a).y * (b).y + (a).z * (b).z)
#define DOT4(a, b) ((a).x * (b).x + (a).y * (b).y + (a).z * (b).z + (a).w * (b).w)

#define DEFINE_VEC(T, LANES, DOT) \
	T T##Add(T a, T b) { T r; LANES(L_ADD) return r; } \
	T T##Subtract(T a, T b) { T r; LANES(L_SUB) return r; } \
	float T##DotProduct(T a, T b) { return DOT(a, b); } \
	T T##MultiplyScalar(T a, float b) { T r; LANES(L_SCALE) return r; } \
	T T##Negate(T a) { T r; LANES(L_NEG) return r; } \
	T T##Normalize(T a) \
	{ \
		T r; \
		float length = sqrtf(DOT(a, a)); \
		if (length > 0) { float b = 1.0f / length; LANES(L_SCALE) return r; } \
		LANES(L_ZERO) return r; \
	} \
	T T##Reflect(T i, T n) \
	{ \
		float d = DOT(n, i); \
		return T##Subtract(i, T##MultiplyScalar(n, 2.f * d)); \
	}

DEFINE_VEC(vec2, LANES2, DOT2)
DEFINE_VEC(vec3, LANES3, DOT3)
DEFINE_VEC(vec4, LANES4, DOT4)
vec2 vec2MultiplyVec2(vec2 a, vec2 b) { vec2 r; LANES2(L_MUL) return r; }
vec3 vec3MultiplyVec3(vec3 a, vec3 b) { vec3 r; LANES3(L_MUL) return r; }
vec4 vec4MultiplyVec4(vec4 a, vec4 b) { vec4 r; LANES4(L_MUL) return r; }

/* ---- mat4 ---- */
#define ROWCOL(A, i, b0, b1, b2, b3) ((A)[i][0] * (b0) + (A)[i][1] * (b1) + (A)[i][2] * (b2) + (A)[i][3] * (b3))

vec4 mat4MultiplyVec4(const mat4* restrict m, vec4 v)
{
	vec4 r;
	r.x = ROWCOL(m->data, 0, v.x, v.y, v.z, v.w);
	r.y = ROWCOL(m->data, 1, v.x, v.y, v.z, v.w);
	r.z = ROWCOL(m->data, 2, v.x, v.y, v.z, v.w);
	r.w = ROWCOL(m->data, 3, v.x, v.y, v.z, v.w);
	return r;
}

mat4 mat4MultiplyMat4(const mat4* restrict a, const mat4* restrict b)
{
	mat4 r;
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++)
			r.data[i][j] = ROWCOL(a->data, i, b->data[0][j], b->data[1][j], b->data[2][j], b->data[3][j]);
	return r;
}

mat4 mat4ConstructScale(float x, float y, float z)
{
	mat4 r = {{{x, 0, 0, 0}, {0, y, 0, 0}, {0, 0, z, 0}, {0, 0, 0, 1}}};
	return r;
}
mat4 mat4ConstructIdentity(void) { return mat4ConstructScale(1, 1, 1); }
mat4 mat4ConstructTranslate(float x, float y, float z)
{
	mat4 r = {{{1, 0, 0, x}, {0, 1, 0, y}, {0, 0, 1, z}, {0, 0, 0, 1}}};
	return r;
}

mat4 mat4ConstructRotate(float x, float y, float z)
{
	const double sx = sin(x), cx = cos(x), sy = sin(y), cy = cos(y), sz = sin(z), cz = cos(z);
	mat4 r = {{{0}}};
	r.data[0][0] = cy * cz;
	r.data[0][1] = sx * sy * cz - cx * sz;
	r.data[0][2] = cx * sy * cz + sx * sz;
	r.data[1][0] = cy * sz;
	r.data[1][1] = sx * sy * sz + cx * cz;
	r.data[1][2] = cx * sy * sz - sx * cz;
	r.data[2][0] = -sy;
	r.data[2][1] = sx * cy;
	r.data[2][2] = cx * cy;
	r.data[3][3] = 1;
	return r;
}

mat4 mat4ConstructTRS(float tx, float ty, float tz, float rx, float ry, float rz, float sx, float sy, float sz)
{
	mat4 T = mat4ConstructTranslate(tx, ty, tz);
	mat4 R = mat4ConstructRotate(rx, ry, rz);
	mat4 S = mat4ConstructScale(sx, sy, sz);
	mat4 RS = mat4MultiplyMat4(&R, &S);
	return mat4MultiplyMat4(&T, &RS);
}

mat4 mat4ConstructView(float cx, float cy, float cz, float rx, float ry, float rz, float sx, float sy, float sz)
{
	return mat4ConstructTRS(-cx, -cy, -cz, -rx, -ry, -rz, sx, sy, sz);
}

mat4 mat4ConstructOrthogonalProjection(float x_min, float x_max, float y_min, float y_max, float z_min, float z_max)
{
	mat4 r = {{{0}}};
	r.data[0][0] = 2 / (x_max - x_min);
	r.data[0][3] = -(x_max + x_min) / (x_max - x_min);
	r.data[1][1] = 2 / (y_max - y_min);
	r.data[1][3] = -(y_max + y_min) / (y_max - y_min);
	r.data[2][2] = 2 / (z_max - z_min);
	r.data[2][3] = -(z_max + z_min) / (z_max - z_min);
	r.data[3][3] = 1;
	return r;
}

mat4 mat4ConstructPerspectiveProjection(float x_min_near, float x_max_near, float y_min_near, float y_max_near,
                                        float z_near, float z_far)
{
	mat4 p = {{{0}}};
	p.data[0][0] = z_near;
	p.data[1][1] = z_near;
	p.data[2][2] = z_near + z_far;
	p.data[2][3] = -z_near * z_far;
	p.data[3][2] = 1;
	mat4 o = mat4ConstructOrthogonalProjection(x_min_near, x_max_near, y_min_near, y_max_near, z_near, z_far);
	return mat4MultiplyMat4(&o, &p);
}
