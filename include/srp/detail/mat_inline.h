/* srp-b200 -- inline C++/CUDA twins of the mat4 functions.
 * mat4MultiplyVec4 / mat4MultiplyMat4 follow reference src/math/mat.c:16-75: each
 * output element is ((p0 + p1) + p2) + p3 with individually rounded products.  The
 * constructors (src/math/mat.c:77-215) run sin/cos in double and are meant for host
 * use; they are provided here so that C++ host code can build uniforms too. */
#pragma once
#include <math.h>
#include "srp/detail/fpops.h"

#define SRP_DOT4_ROWCOL(A, i, B0, B1, B2, B3) \
	SRP_FADD(SRP_FADD(SRP_FADD(SRP_FMUL((A)[i][0], (B0)), SRP_FMUL((A)[i][1], (B1))), \
	                  SRP_FMUL((A)[i][2], (B2))), SRP_FMUL((A)[i][3], (B3)))

SRP_HD vec4 mat4MultiplyVec4(const mat4* m, vec4 v)
{
	vec4 r;
	r.x = SRP_DOT4_ROWCOL(m->data, 0, v.x, v.y, v.z, v.w);
	r.y = SRP_DOT4_ROWCOL(m->data, 1, v.x, v.y, v.z, v.w);
	r.z = SRP_DOT4_ROWCOL(m->data, 2, v.x, v.y, v.z, v.w);
	r.w = SRP_DOT4_ROWCOL(m->data, 3, v.x, v.y, v.z, v.w);
	return r;
}

SRP_HD mat4 mat4MultiplyMat4(const mat4* a, const mat4* b)
{
	mat4 r;
	for (int i = 0; i < 4; i++)
		for (int j = 0; j < 4; j++)
			r.data[i][j] = SRP_DOT4_ROWCOL(a->data, i, b->data[0][j], b->data[1][j], b->data[2][j], b->data[3][j]);
	return r;
}

SRP_HD mat4 mat4ConstructScale(float x, float y, float z)
{
	mat4 r = {{{x, 0, 0, 0}, {0, y, 0, 0}, {0, 0, z, 0}, {0, 0, 0, 1}}};
	return r;
}
SRP_HD mat4 mat4ConstructIdentity() { return mat4ConstructScale(1, 1, 1); }
SRP_HD mat4 mat4ConstructTranslate(float x, float y, float z)
{
	mat4 r = {{{1, 0, 0, x}, {0, 1, 0, y}, {0, 0, 1, z}, {0, 0, 0, 1}}};
	return r;
}
SRP_HD mat4 mat4ConstructRotate(float x, float y, float z)
{
	/* products and sums are evaluated in double and rounded once per element */
	const double sx = sin((double) x), cx = cos((double) x);
	const double sy = sin((double) y), cy = cos((double) y);
	const double sz = sin((double) z), cz = cos((double) z);
	mat4 r = {{{0}}};
	r.data[0][0] = (float) SRP_DMUL(cy, cz);
	r.data[0][1] = (float) SRP_DSUB(SRP_DMUL(SRP_DMUL(sx, sy), cz), SRP_DMUL(cx, sz));
	r.data[0][2] = (float) SRP_DADD(SRP_DMUL(SRP_DMUL(cx, sy), cz), SRP_DMUL(sx, sz));
	r.data[1][0] = (float) SRP_DMUL(cy, sz);
	r.data[1][1] = (float) SRP_DADD(SRP_DMUL(SRP_DMUL(sx, sy), sz), SRP_DMUL(cx, cz));
	r.data[1][2] = (float) SRP_DSUB(SRP_DMUL(SRP_DMUL(cx, sy), sz), SRP_DMUL(sx, cz));
	r.data[2][0] = (float) -sy;
	r.data[2][1] = (float) SRP_DMUL(sx, cy);
	r.data[2][2] = (float) SRP_DMUL(cx, cy);
	r.data[3][3] = 1;
	return r;
}
SRP_HD mat4 mat4ConstructTRS(float tx, float ty, float tz, float rx, float ry, float rz,
                             float sx, float sy, float sz)
{
	mat4 T = mat4ConstructTranslate(tx, ty, tz);
	mat4 R = mat4ConstructRotate(rx, ry, rz);
	mat4 S = mat4ConstructScale(sx, sy, sz);
	mat4 RS = mat4MultiplyMat4(&R, &S);
	return mat4MultiplyMat4(&T, &RS);
}
SRP_HD mat4 mat4ConstructView(float cx, float cy, float cz, float rx, float ry, float rz,
                              float sx, float sy, float sz)
{
	return mat4ConstructTRS(-cx, -cy, -cz, -rx, -ry, -rz, sx, sy, sz);
}
SRP_HD mat4 mat4ConstructOrthogonalProjection(float x_min, float x_max, float y_min, float y_max,
                                              float z_min, float z_max)
{
	mat4 r = {{{0}}};
	r.data[0][0] = SRP_FDIV(2.f, SRP_FSUB(x_max, x_min));
	r.data[0][3] = SRP_FDIV(-SRP_FADD(x_max, x_min), SRP_FSUB(x_max, x_min));
	r.data[1][1] = SRP_FDIV(2.f, SRP_FSUB(y_max, y_min));
	r.data[1][3] = SRP_FDIV(-SRP_FADD(y_max, y_min), SRP_FSUB(y_max, y_min));
	r.data[2][2] = SRP_FDIV(2.f, SRP_FSUB(z_max, z_min));
	r.data[2][3] = SRP_FDIV(-SRP_FADD(z_max, z_min), SRP_FSUB(z_max, z_min));
	r.data[3][3] = 1;
	return r;
}
SRP_HD mat4 mat4ConstructPerspectiveProjection(float x_min_near, float x_max_near,
                                               float y_min_near, float y_max_near,
                                               float z_near, float z_far)
{
	mat4 p = {{{0}}};
	p.data[0][0] = z_near;
	p.data[1][1] = z_near;
	p.data[2][2] = SRP_FADD(z_near, z_far);
	p.data[2][3] = SRP_FMUL(-z_near, z_far);
	p.data[3][2] = 1;
	mat4 o = mat4ConstructOrthogonalProjection(x_min_near, x_max_near, y_min_near, y_max_near, z_near, z_far);
	return mat4MultiplyMat4(&o, &p);
}
