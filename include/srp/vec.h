/* srp-b200 -- vec2 / vec3 / vec4 helpers (API of the reference's include/srp/vec.h:15-104).
 *
 * Types: unions of named lanes, sub-vector views and a flat array, packed to
 * alignment 1 exactly like the reference (so user vertex structs such as
 * `{ vec3 position; uint8_t color; }` keep their 13-byte layout on both sides).
 *
 * Host C programs get the 24 functions as ordinary library symbols.  When the header
 * is compiled by nvcc the same names are defined inline as __host__ __device__
 * functions (srp/detail/math_inline.h) whose device bodies use the _rn intrinsics, so
 * a shader twin computes bit-identical results to the C original built without FMA
 * contraction (reference src/math/vec.c:16-189, left-associative sums of products). */
#pragma once
#include <stdint.h>

#pragma pack(push, 1)
typedef union vec2 {
	struct { float x, y; };
	float v[2];
} vec2;

typedef union vec3 {
	struct { float x, y, z; };
	struct { vec2 xy; float _z; };
	struct { float _x; vec2 yz; };
	float v[3];
} vec3;

typedef union vec4 {
	struct { float x, y, z, w; };
	struct { vec2 xy; float _z, _w; };
	struct { float _x; vec2 yz; float __w; };
	struct { float __x, _y; vec2 zw; };
	struct { vec3 xyz; float ___w; };
	struct { float ___x; vec3 yzw; };
	float v[4];
} vec4;
#pragma pack(pop)

/* constructors and swizzles (compound literals; accepted by gcc, g++ and nvcc) */
#define VEC2(x, y)           ((vec2) {{x, y}})
#define VEC3(x, y, z)        ((vec3) {{x, y, z}})
#define VEC4(x, y, z, w)     ((vec4) {{x, y, z, w}})
#define SWZ2(v, a, b)        ((vec2) {{(v).a, (v).b}})
#define SWZ3(v, a, b, c)     ((vec3) {{(v).a, (v).b, (v).c}})
#define SWZ4(v, a, b, c, d)  ((vec4) {{(v).a, (v).b, (v).c, (v).d}})

#if defined(__cplusplus)
	/* C++ has no multi-member designated initialisers for unions */
	#include "srp/detail/math_inline.h"
	#define VEC4_FROM_VEC3(v, a) (srpVec4FromVec3((v), (a)))
#else
	#define VEC4_FROM_VEC3(v, a) ((vec4) {.xyz = (v), .___w = (a)})

	#define SRP_DECLARE_VEC_API(T, N) \
		T N##Add(T a, T b); \
		T N##Subtract(T a, T b); \
		float N##DotProduct(T a, T b); \
		T N##MultiplyScalar(T a, float b); \
		T N##Normalize(T v);               /* zero vector stays zero */ \
		T N##Reflect(T i, T n);            /* i - 2*dot(n,i)*n */ \
		T N##Multiply##T(T a, T b);        /* component-wise */ \
		T N##Negate(T v);
	#define vec2Multiplyvec2 vec2MultiplyVec2
	#define vec3Multiplyvec3 vec3MultiplyVec3
	#define vec4Multiplyvec4 vec4MultiplyVec4
	SRP_DECLARE_VEC_API(vec2, vec2)
	SRP_DECLARE_VEC_API(vec3, vec3)
	SRP_DECLARE_VEC_API(vec4, vec4)
	#undef vec2Multiplyvec2
	#undef vec3Multiplyvec3
	#undef vec4Multiplyvec4
	#undef SRP_DECLARE_VEC_API
#endif
