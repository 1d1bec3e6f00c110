/* srp-b200 -- inline C++/CUDA twins of the vec2/3/4 functions.
 * Operation order follows reference src/math/vec.c:16-189 (sums associate to the
 * left; Normalize = sqrtf of the squared length, then multiply by 1/len; Reflect =
 * i - n*(2*dot(n,i))).  Included from srp/vec.h in C++ translation units only. */
#pragma once
#include "srp/detail/fpops.h"

SRP_HD vec4 srpVec4FromVec3(vec3 v, float a) { vec4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = a; return r; }

#define SRP_VEC_FOR2(OP) OP(x) OP(y)
#define SRP_VEC_FOR3(OP) OP(x) OP(y) OP(z)
#define SRP_VEC_FOR4(OP) OP(x) OP(y) OP(z) OP(w)

#define SRP_DEFINE_VEC_TWINS(T, FOR, DOT_EXPR, ZERO) \
	SRP_HD T T##Add(T a, T b) { T r; FOR(SRP_VEC_ADD_LANE) return r; } \
	SRP_HD T T##Subtract(T a, T b) { T r; FOR(SRP_VEC_SUB_LANE) return r; } \
	SRP_HD float T##DotProduct(T a, T b) { return DOT_EXPR; } \
	SRP_HD T T##MultiplyScalar(T a, float b) { T r; FOR(SRP_VEC_SCALE_LANE) return r; } \
	SRP_HD T T##Negate(T a) { T r; FOR(SRP_VEC_NEG_LANE) return r; } \
	SRP_HD T T##Normalize(T a) { \
		T b = a; float length = SRP_FSQRT(DOT_EXPR); \
		if (length > 0) { float inv = SRP_FDIV(1.0f, length); b = T##MultiplyScalar(a, inv); return b; } \
		return ZERO; } \
	SRP_HD T T##Reflect(T i, T n) { \
		float d = T##DotProduct(n, i); \
		return T##Subtract(i, T##MultiplyScalar(n, SRP_FMUL(2.f, d))); }

#define SRP_VEC_ADD_LANE(l)   r.l = SRP_FADD(a.l, b.l);
#define SRP_VEC_SUB_LANE(l)   r.l = SRP_FSUB(a.l, b.l);
#define SRP_VEC_SCALE_LANE(l) r.l = SRP_FMUL(a.l, b);
#define SRP_VEC_NEG_LANE(l)   r.l = -a.l;
#define SRP_VEC_HAD_LANE(l)   r.l = SRP_FMUL(a.l, b.l);

#define SRP_DOT2 SRP_FADD(SRP_FMUL(a.x, b.x), SRP_FMUL(a.y, b.y))
#define SRP_DOT3 SRP_FADD(SRP_DOT2, SRP_FMUL(a.z, b.z))
#define SRP_DOT4 SRP_FADD(SRP_DOT3, SRP_FMUL(a.w, b.w))

SRP_DEFINE_VEC_TWINS(vec2, SRP_VEC_FOR2, SRP_DOT2, (VEC2(0, 0)))
SRP_DEFINE_VEC_TWINS(vec3, SRP_VEC_FOR3, SRP_DOT3, (VEC3(0, 0, 0)))
SRP_DEFINE_VEC_TWINS(vec4, SRP_VEC_FOR4, SRP_DOT4, (VEC4(0, 0, 0, 0)))

SRP_HD vec2 vec2MultiplyVec2(vec2 a, vec2 b) { vec2 r; SRP_VEC_FOR2(SRP_VEC_HAD_LANE) return r; }
SRP_HD vec3 vec3MultiplyVec3(vec3 a, vec3 b) { vec3 r; SRP_VEC_FOR3(SRP_VEC_HAD_LANE) return r; }
SRP_HD vec4 vec4MultiplyVec4(vec4 a, vec4 b) { vec4 r; SRP_VEC_FOR4(SRP_VEC_HAD_LANE) return r; }
