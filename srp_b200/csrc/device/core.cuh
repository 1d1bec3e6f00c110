/* srp-b200 device core -- the exact arithmetic of the draw path.
 *
 * Every function here is __host__ __device__ and written only in the rounding-exact
 * operations of srp/detail/fpops.h (no operator that the compiler could contract into
 * an FMA), in the evaluation order of the reference, so that results are bit-identical
 * to kitrofimov/srp built with its own flags (ISO C => no FP contraction; SURVEY.md
 * App. A).  The kernels in geom.cu / raster.cu call these; nothing here touches
 * memory layout decisions other than the primitive record defined below.
 *
 * Reference citations are given per function as file:line under /root/reference. */
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>
#include "srp/api.h"
#include "srp/detail/fpops.h"
#include "draw_types.h"

#define SRPD_EPS 1e-9   /* reference src/math/utils.h:38 (double literal) */

/* ---------------------------------------------------------------------------------
 * Primitive record (HBM, written by the geometry kernel, read by the tile kernel).
 * 80-byte header of 20 words followed by nVerts varyings blobs of st.slotSize bytes;
 * records are 16-byte aligned so a warp can fetch the header with five broadcast
 * 16-byte loads.
 *
 *   word   TRIANGLE                         LINE                    POINT
 *   0-2    lambda at (minX+.5, minY+.5)     x, y (segment), xInc    minBP.x, minBP.y, maxBP.x
 *   3      minX | maxX<<16                  yInc                    maxBP.y
 *   4-6    dlambda/dx                       tInc, count, z0*iw0     minX, maxX, minY  (inclusive ints)
 *   7      minY | maxY<<16                  z1*iw1                  maxY
 *   8-10   dlambda/dy                       iw0, iw1, t (segment)   ndc z, ndc w (=1), -
 *   11     flags: bit0-2 edge TL, bit3 frontFacing
 *   12-14  z_i * iw_i                       -                       -
 *   15     primitive id (all kinds)
 *   16-18  iw_i (1 / clip w)                -                       -
 *   19     1 + offset into the barycentric checkpoint table (large triangles), 0 = none
 * For PERSPECTIVE float/double attributes the blobs hold a_i * iw_i (the first product
 * of the reference's `a_i * invW[i] * weights[i]`, interpolation.c:72), everything else
 * is stored verbatim. */
#define SRPD_REC_HEADER_WORDS 20
#define SRPD_REC_HEADER_BYTES 80

struct SrpdPos { float x, y, z, w; };

SRP_HD uint32_t srpdF2U(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
SRP_HD float srpdU2F(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

SRP_HD uint32_t srpdRecordStride(const SrpdState& st, int nVerts)
{
	return (uint32_t) ((SRPD_REC_HEADER_BYTES + nVerts * st.slotSize + 15) & ~15);
}
SRP_HD int srpdVertsOfKind(uint32_t kind) { return kind == SRPD_KIND_TRIANGLE ? 3 : (kind == SRPD_KIND_LINE ? 2 : 1); }

/* reference ROUGHLY_ZERO (src/math/utils.h:40-42): fabs((double) x) <= 1e-9.  For a float
 * argument that is the same predicate as |x| <= F with F the largest float not above 1e-9,
 * F = 0x3089705F = 9.999999717180685e-10 (the next float, 1.0000000605e-09, is above). */
SRP_HD bool srpdRoughlyZero(float x) { return fabsf(x) <= srpdU2F(0x3089705Fu); }

/* ---------------------------------------------------------------------------------
 * applyPerspectiveDivide, reference src/pipeline/vertex_processing.c:76-88.
 * invW = (float)(1.0 / (double) w); a correctly rounded double quotient rounded again
 * to float equals the float quotient (53 >= 2*24+2), so the f32 divide is used. */
SRP_HD float srpdPerspectiveDivide(SrpdPos& p)
{
	float invW = SRP_FDIV(1.0f, p.w);
	p.x = SRP_FMUL(p.x, invW);
	p.y = SRP_FMUL(p.y, invW);
	p.z = SRP_FMUL(p.z, invW);
	p.w = 1.0f;
	return invW;
}

/* framebufferNDCToScreenSpace, reference src/core/framebuffer.c:48-55: evaluated in
 * double, rounded to float once per component. */
SRP_HD void srpdNdcToScreen(const SrpdState& st, const SrpdPos& ndc, float& sx, float& sy)
{
	double halfW = SRP_DDIV((double) (float) st.width, 2.0);
	double halfH = SRP_DDIV((double) (float) st.height, 2.0);
	sx = (float) SRP_DMUL(halfW, SRP_DADD((double) ndc.x, 1.0));
	sy = (float) SRP_DMUL(halfH, SRP_DSUB(1.0, (double) ndc.y));
}

/* a.x*b.y - a.y*b.x, reference src/raster/triangle.c:228-233 */
SRP_HD float srpdCross2(float ax, float ay, float bx, float by)
{
	return SRP_FSUB(SRP_FMUL(ax, by), SRP_FMUL(ay, bx));
}

/* clip-space outcode, reference src/pipeline/clipping.c:121-137 */
SRP_HD uint32_t srpdClipCode(const SrpdPos& p)
{
	uint32_t c = 0;
	c |= (uint32_t) (SRP_FADD(p.x, p.w) < 0) << 0;
	c |= (uint32_t) (SRP_FSUB(p.w, p.x) < 0) << 1;
	c |= (uint32_t) (SRP_FADD(p.y, p.w) < 0) << 2;
	c |= (uint32_t) (SRP_FSUB(p.w, p.y) < 0) << 3;
	c |= (uint32_t) (SRP_FADD(p.z, p.w) < 0) << 4;
	c |= (uint32_t) (SRP_FSUB(p.w, p.z) < 0) << 5;
	return c;
}

/* distance to clip plane `plane` (0..5 = L,R,B,T,N,F), reference clipping.c:260-277 */
SRP_HD float srpdPlaneDistance(const SrpdPos& p, int plane)
{
	switch (plane)
	{
		case 0:  return SRP_FADD(p.x, p.w);
		case 1:  return SRP_FSUB(p.w, p.x);
		case 2:  return SRP_FADD(p.y, p.w);
		case 3:  return SRP_FSUB(p.w, p.y);
		case 4:  return SRP_FADD(p.z, p.w);
		default: return SRP_FSUB(p.w, p.z);
	}
}

/* ---------------------------------------------------------------------------------
 * Varyings blobs.  Reading/writing through memcpy keeps unaligned user layouts legal. */
template <typename T> SRP_HD T srpdLoad(const unsigned char* p) { T v; memcpy(&v, p, sizeof(T)); return v; }
template <typename T> SRP_HD void srpdStore(unsigned char* p, T v) { memcpy(p, &v, sizeof(T)); }

SRP_HD bool srpdTypeIsFloat(uint8_t type) { return type == SRP_FLOAT; }
SRP_HD bool srpdTypeIsDouble(uint8_t type) { return type == SRP_DOUBLE; }

/* Two-vertex affine blend used for clip-generated vertices: interpolateVertex ->
 * interpolateAttributes(nVertices = 2, invW = NULL), reference clipping.c:244-258 and
 * interpolation.c:93-163.  Floating attributes: value = 0; value += a*w0; value += b*w1
 * (perspective and FLAT are both forced to affine, App. B-15); integer attributes copy
 * the provoking endpoint (FIRST -> a, LAST -> b). */
SRP_HD void srpdBlendVaryings(const SrpdState& st, const unsigned char* a, const unsigned char* b,
                              float w0, float w1, unsigned char* out)
{
	if (st.allFloat)
	{
		/* all-float, 4-byte aligned layout: every float is blended (FLAT ones too, App. B-15) */
		const float* fa = (const float*) a;
		const float* fb = (const float*) b;
		float* fo = (float*) out;
		for (int e = 0; e < st.nFloats; e++)
		{
			float v = 0.f;
			v = SRP_FADD(v, SRP_FMUL(fa[e], w0));
			v = SRP_FADD(v, SRP_FMUL(fb[e], w1));
			fo[e] = v;
		}
		return;
	}
	for (int ai = 0; ai < st.nVaryings; ai++)
	{
		const SrpdVarying& at = st.varyings[ai];
		if (srpdTypeIsFloat(at.type))
			for (int e = 0; e < at.nItems; e++)
			{
				int off = at.offset + 4 * e;
				float v = 0.f;
				v = SRP_FADD(v, SRP_FMUL(srpdLoad<float>(a + off), w0));
				v = SRP_FADD(v, SRP_FMUL(srpdLoad<float>(b + off), w1));
				srpdStore<float>(out + off, v);
			}
		else if (srpdTypeIsDouble(at.type))
			for (int e = 0; e < at.nItems; e++)
			{
				int off = at.offset + 8 * e;
				double v = 0.;
				v = SRP_DADD(v, SRP_DMUL(srpdLoad<double>(a + off), (double) w0));
				v = SRP_DADD(v, SRP_DMUL(srpdLoad<double>(b + off), (double) w1));
				srpdStore<double>(out + off, v);
			}
		else
		{
			const unsigned char* src = st.provokingFirst ? a : b;
			for (int k = 0; k < at.nItems * at.elemSize; k++)
				out[at.offset + k] = src[at.offset + k];
		}
	}
}

/* position part of interpolateVertex: a*(1-t) + b*t per component, clipping.c:249-250 */
SRP_HD SrpdPos srpdBlendPos(const SrpdPos& a, const SrpdPos& b, float t)
{
	float omt = SRP_FSUB(1.0f, t);
	SrpdPos r;
	r.x = SRP_FADD(SRP_FMUL(a.x, omt), SRP_FMUL(b.x, t));
	r.y = SRP_FADD(SRP_FMUL(a.y, omt), SRP_FMUL(b.y, t));
	r.z = SRP_FADD(SRP_FMUL(a.z, omt), SRP_FMUL(b.z, t));
	r.w = SRP_FADD(SRP_FMUL(a.w, omt), SRP_FMUL(b.w, t));
	return r;
}

/* Blob as it is stored in a primitive record: PERSPECTIVE floating attributes are
 * pre-multiplied by the vertex' 1/w (first product of interpolation.c:72); `persp`
 * is false for POINT records (varyings pass through, point.c:66). */
SRP_HD void srpdStoreBlob(const SrpdState& st, const unsigned char* src, float invW, bool persp, unsigned char* dst)
{
	for (int k = 0; k < st.slotSize; k++)
		dst[k] = src[k];
	if (!persp)
		return;
	for (int ai = 0; ai < st.nVaryings; ai++)
	{
		const SrpdVarying& at = st.varyings[ai];
		if (at.mode != SRP_INTERPOLATION_MODE_PERSPECTIVE)
			continue;
		if (srpdTypeIsFloat(at.type))
			for (int e = 0; e < at.nItems; e++)
				srpdStore<float>(dst + at.offset + 4 * e, SRP_FMUL(srpdLoad<float>(src + at.offset + 4 * e), invW));
		else if (srpdTypeIsDouble(at.type))
			for (int e = 0; e < at.nItems; e++)
				srpdStore<double>(dst + at.offset + 8 * e, SRP_DMUL(srpdLoad<double>(src + at.offset + 8 * e), (double) invW));
	}
}

/* ---------------------------------------------------------------------------------
 * Triangle setup, reference src/raster/triangle.c:113-238.
 * Input: three clip-space positions in API order.  Output: header words + the vertex
 * order the rasteriser sees (`order[i]` = which input vertex became v[i] after the
 * winding normalisation, needed to place the varyings blobs).
 * Returns 0 if the triangle is culled / degenerate (no primitive id is consumed),
 * 1 if it survives; *stored tells whether its clamped bounding box is non-empty. */
struct SrpdTriSetup
{
	uint32_t w[SRPD_REC_HEADER_WORDS];
	float invW[3];
	int order[3];
	uint16_t minX, minY, maxX, maxY;
};

SRP_HD int srpdSetupTriangle(const SrpdState& st, const SrpdPos clip[3], SrpdTriSetup& out, bool& stored)
{
	SrpdPos v[3] = { clip[0], clip[1], clip[2] };
	float iw[3];
	for (int i = 0; i < 3; i++)
		iw[i] = srpdPerspectiveDivide(v[i]);

	/* shouldCullTriangle, triangle.c:162-182 (signed area in NDC, y up) */
	float e0x = SRP_FSUB(v[1].x, v[0].x), e0y = SRP_FSUB(v[1].y, v[0].y);
	float e1x = SRP_FSUB(v[2].x, v[0].x), e1y = SRP_FSUB(v[2].y, v[0].y);
	float signedArea = srpdCross2(e0x, e0y, e1x, e1y);
	bool isCCW = signedArea > 0;
	if (st.cullFace == SRP_FACE_FRONT_AND_BACK)
		return 0;
	bool frontFacing = ((signedArea > 0) & (st.frontFaceCW == 0)) || ((signedArea < 0) & (st.frontFaceCW != 0));
	bool cull = (frontFacing && st.cullFace == SRP_FACE_FRONT) || (!frontFacing && st.cullFace == SRP_FACE_BACK);
	if (cull)
		return 0;

	/* triangleChangeWinding, triangle.c:184-195: raster always sees CCW-in-NDC */
	int order[3] = { 0, 1, 2 };
	if (!isCCW)
	{
		SrpdPos tp = v[1]; v[1] = v[2]; v[2] = tp;
		float ti = iw[1]; iw[1] = iw[2]; iw[2] = ti;
		order[1] = 2; order[2] = 1;
	}

	float sx[3], sy[3];
	for (int i = 0; i < 3; i++)
		srpdNdcToScreen(st, v[i], sx[i], sy[i]);

	float ex[3], ey[3];   /* edge i = ss[i+1] - ss[i] */
	for (int i = 0; i < 3; i++)
	{
		ex[i] = SRP_FSUB(sx[(i + 1) % 3], sx[i]);
		ey[i] = SRP_FSUB(sy[(i + 1) % 3], sy[i]);
	}

	float areaX2 = (float) fabs((double) srpdCross2(ex[0], ey[0], ex[2], ey[2]));
	if (srpdRoughlyZero(areaX2))
		return 0;

	/* bounding box, triangle.c:139-146: MIN(a,b) = a > b ? b : a, MAX(a,b) = a > b ? a : b */
	#define SRPD_MIN(a, b) ((a) > (b) ? (b) : (a))
	#define SRPD_MAX(a, b) ((a) > (b) ? (a) : (b))
	float mnx = SRPD_MIN(sx[0], SRPD_MIN(sx[1], sx[2]));
	float mny = SRPD_MIN(sy[0], SRPD_MIN(sy[1], sy[2]));
	float mxx = SRPD_MAX(sx[0], SRPD_MAX(sx[1], sx[2]));
	float mxy = SRPD_MAX(sy[0], SRPD_MAX(sy[1], sy[2]));
	double fminx = floor((double) mnx), fminy = floor((double) mny);
	double cmaxx = ceil((double) mxx), cmaxy = ceil((double) mxy);
	float minBPx = (float) SRPD_MAX(fminx, 0.0);
	float minBPy = (float) SRPD_MAX(fminy, 0.0);
	float maxBPx = (float) SRPD_MIN(cmaxx, (double) st.width);
	float maxBPy = (float) SRPD_MIN(cmaxy, (double) st.height);
	#undef SRPD_MIN
	#undef SRPD_MAX

	/* calculateBarycentrics at (minBP + 0.5), triangle.c:197-226 */
	float px = (float) ((double) minBPx + 0.5), py = (float) ((double) minBPy + 0.5);
	float apx = SRP_FSUB(px, sx[0]), apy = SRP_FSUB(py, sy[0]);
	float bpx = SRP_FSUB(px, sx[1]), bpy = SRP_FSUB(py, sy[1]);
	float cpx = SRP_FSUB(px, sx[2]), cpy = SRP_FSUB(py, sy[2]);
	float l0 = SRP_FDIV(srpdCross2(bpx, bpy, ex[1], ey[1]), areaX2);
	float l1 = SRP_FDIV(srpdCross2(cpx, cpy, ex[2], ey[2]), areaX2);
	float l2 = SRP_FDIV(srpdCross2(apx, apy, ex[0], ey[0]), areaX2);
	float dx0 = SRP_FDIV(ey[1], areaX2), dx1 = SRP_FDIV(ey[2], areaX2), dx2 = SRP_FDIV(ey[0], areaX2);
	float dy0 = SRP_FDIV(-ex[1], areaX2), dy1 = SRP_FDIV(-ex[2], areaX2), dy2 = SRP_FDIV(-ex[0], areaX2);

	/* isEdgeFlatTopOrLeft, triangle.c:235-238; edgeTL[i] belongs to edge i and is the
	 * flag tested together with lambda[i] (triangle.c:84,156) */
	uint32_t flags = 0;
	for (int i = 0; i < 3; i++)
	{
		bool tl = ((ex[i] > 0) && srpdRoughlyZero(ey[i])) || (ey[i] < 0);
		flags |= (uint32_t) tl << i;
	}
	flags |= (uint32_t) frontFacing << 3;

	/* loop bounds of rasterizeTriangle (triangle.c:78-80): size_t y = minBP.y; y < maxBP.y */
	long long iMinX = (long long) minBPx, iMinY = (long long) minBPy;
	long long iMaxX = (long long) ceil((double) maxBPx), iMaxY = (long long) ceil((double) maxBPy);
	stored = (iMinX < iMaxX) && (iMinY < iMaxY) && iMaxX > 0 && iMaxY > 0;
	out.minX = (uint16_t) (stored ? iMinX : 0); out.maxX = (uint16_t) (stored ? iMaxX : 0);
	out.minY = (uint16_t) (stored ? iMinY : 0); out.maxY = (uint16_t) (stored ? iMaxY : 0);

	out.w[0] = srpdF2U(l0);  out.w[1] = srpdF2U(l1);  out.w[2] = srpdF2U(l2);
	out.w[3] = (uint32_t) out.minX | ((uint32_t) out.maxX << 16);
	out.w[4] = srpdF2U(dx0); out.w[5] = srpdF2U(dx1); out.w[6] = srpdF2U(dx2);
	out.w[7] = (uint32_t) out.minY | ((uint32_t) out.maxY << 16);
	out.w[8] = srpdF2U(dy0); out.w[9] = srpdF2U(dy1); out.w[10] = srpdF2U(dy2);
	out.w[11] = flags;
	for (int i = 0; i < 3; i++)
	{
		out.w[12 + i] = srpdF2U(SRP_FMUL(v[i].z, iw[i]));   /* z_i * iw_i, interpolation.c:44 */
		out.w[16 + i] = srpdF2U(iw[i]);
		out.invW[i] = iw[i];
		out.order[i] = order[i];
	}
	out.w[15] = 0; out.w[19] = 0;
	return 1;
}

/* Exact coverage pre-test for small triangles: replays rasterizeTriangle's loop
 * (triangle.c:78-110) over the bounding box with the same incremental barycentrics and the
 * same top-left inside test, and reports whether ANY pixel would emit a fragment. */
SRP_HD bool srpdTriangleIsSmall(const SrpdTriSetup& s)
{
	return (int) (s.maxX - s.minX) <= 4 && (int) (s.maxY - s.minY) <= 4;
}
SRP_HD bool srpdSmallTriangleCoversAnyPixel(const SrpdTriSetup& s)
{
	float row0 = srpdU2F(s.w[0]), row1 = srpdU2F(s.w[1]), row2 = srpdU2F(s.w[2]);
	const float dx0 = srpdU2F(s.w[4]), dx1 = srpdU2F(s.w[5]), dx2 = srpdU2F(s.w[6]);
	const float dy0 = srpdU2F(s.w[8]), dy1 = srpdU2F(s.w[9]), dy2 = srpdU2F(s.w[10]);
	const uint32_t flags = s.w[11];
	for (int y = s.minY; y < s.maxY; y++)
	{
		float l0 = row0, l1 = row1, l2 = row2;
		for (int x = s.minX; x < s.maxX; x++)
		{
			const bool in0 = (l0 > 0.f) || (srpdRoughlyZero(l0) && (flags & 1u));
			const bool in1 = (l1 > 0.f) || (srpdRoughlyZero(l1) && (flags & 2u));
			const bool in2 = (l2 > 0.f) || (srpdRoughlyZero(l2) && (flags & 4u));
			if (in0 && in1 && in2)
				return true;
			l0 = SRP_FADD(l0, dx0); l1 = SRP_FADD(l1, dx1); l2 = SRP_FADD(l2, dx2);
		}
		row0 = SRP_FADD(row0, dy0); row1 = SRP_FADD(row1, dy1); row2 = SRP_FADD(row2, dy2);
	}
	return false;
}

/* ---------------------------------------------------------------------------------
 * Line setup, reference src/raster/line.c:34-55,79-86 (the loop-invariant part of
 * rasterizeLine is hoisted here).  Both endpoints are clip-space positions. */
struct SrpdLineSetup
{
	float x0, y0, xInc, yInc, tInc;
	int steps;                         /* the DDA emits steps + 1 fragments */
	float zw[2], invW[2];
};

SRP_HD void srpdSetupLine(const SrpdState& st, const SrpdPos clip[2], SrpdLineSetup& out)
{
	SrpdPos v[2] = { clip[0], clip[1] };
	out.invW[0] = srpdPerspectiveDivide(v[0]);
	out.invW[1] = srpdPerspectiveDivide(v[1]);
	float x1, y1;
	srpdNdcToScreen(st, v[0], out.x0, out.y0);
	srpdNdcToScreen(st, v[1], x1, y1);

	const float dx = SRP_FSUB(x1, out.x0), dy = SRP_FSUB(y1, out.y0);
	int steps = (int) ceil(fmax(fabs((double) dx), fabs((double) dy)));
	if (steps == 0)
		steps = 1;
	out.steps = steps;
	out.xInc = SRP_FDIV(dx, (float) steps);
	out.yInc = SRP_FDIV(dy, (float) steps);
	out.tInc = (float) SRP_DDIV(1.0, (double) steps);
	out.zw[0] = SRP_FMUL(v[0].z, out.invW[0]);
	out.zw[1] = SRP_FMUL(v[1].z, out.invW[1]);
}

/* round half away from zero of a float (C `round` on the promoted double, line.c:58-59);
 * exact because d + 0.5 is exact for a float-valued double of this magnitude */
SRP_HD int srpdRoundToInt(float v)
{
	const double d = (double) v;
	return (int) trunc(d + copysign(0.5, d));
}

/* A line is stored as SEGMENTS of up to SRPD_LINE_SEG consecutive DDA fragments so that a
 * tile never has to replay more than a segment of the chain: the geometry kernel walks the
 * whole chain once (serial float additions, line.c:72-74) and records, per segment, the
 * chain state (x, y, t) at its first fragment and the exact bounding box of the pixels its
 * fragments land on -- including fragments whose x == width wraps them onto the next row
 * through the reference's unchecked y*W + x index; fragments whose index
 * falls outside the planes are not representable and are dropped.
 *
 * LINE record words: 0 x, 1 y at the segment's first fragment, 2 xInc, 3 yInc, 4 tInc,
 * 5 fragment count, 6 z0*iw0, 7 z1*iw1, 8 iw0, 9 iw1, 10 t at the first fragment, 15 id. */
#define SRPD_LINE_SEG 16

struct SrpdLineSegment
{
	uint32_t w[SRPD_REC_HEADER_WORDS];
	uint16_t minX, minY, maxX, maxY;   /* half-open pixel bbox of the fragments inside the planes */
	bool any;
};

/* advances (x, y, t) over `n` fragments and fills `seg` */
SRP_HD void srpdLineSegment(const SrpdState& st, const SrpdLineSetup& ln, float& x, float& y, float& t, int n, SrpdLineSegment& seg)
{
	for (int k = 0; k < SRPD_REC_HEADER_WORDS; k++)
		seg.w[k] = 0;
	seg.w[0] = srpdF2U(x); seg.w[1] = srpdF2U(y); seg.w[2] = srpdF2U(ln.xInc); seg.w[3] = srpdF2U(ln.yInc);
	seg.w[4] = srpdF2U(ln.tInc); seg.w[5] = (uint32_t) n;
	seg.w[6] = srpdF2U(ln.zw[0]); seg.w[7] = srpdF2U(ln.zw[1]);
	seg.w[8] = srpdF2U(ln.invW[0]); seg.w[9] = srpdF2U(ln.invW[1]);
	seg.w[10] = srpdF2U(t);
	/* The chain: n - 1 float additions per axis (line.c:72-74).  Adding one constant over and over
	 * is monotone, and so is rounding: every fragment's pixel lies between the pixels of the
	 * segment's first and last fragment, so the exact box needs two roundings per axis instead of
	 * one per fragment.  A fixed trip count with the body under `k < n`: the lanes of a warp (one
	 * line each, of different lengths) stay together. */
	const float xs = x, ys = y;
	float xl = x, yl = y;
	for (int k = 0; k < SRPD_LINE_SEG; k++)
		if (k < n)
		{
			xl = x; yl = y;
			x = SRP_FADD(x, ln.xInc);
			y = SRP_FADD(y, ln.yInc);
			t = SRP_FADD(t, ln.tInc);
		}
	int x0, y0, x1, y1;
	{
		const int ax = srpdRoundToInt(xs), ay = srpdRoundToInt(ys), bx = srpdRoundToInt(xl), by = srpdRoundToInt(yl);
		x0 = ax < bx ? ax : bx; x1 = ax < bx ? bx : ax;
		y0 = ay < by ? ay : by; y1 = ay < by ? by : ay;
	}
	if (x0 < 0 || y0 < 0 || x1 >= st.width || y1 >= st.height)
	{
		/* the segment touches the framebuffer's border: fragment by fragment, with the reference's
		 * unchecked y*W + x index (a fragment with x == width lands on the next row, App. B-1) */
		x0 = 1 << 30; y0 = 1 << 30; x1 = -1; y1 = -1;
		const long long W = st.width, total = (long long) st.width * st.height;
		float fx = xs, fy = ys;
		for (int k = 0; k < n; k++)
		{
			const int ix = srpdRoundToInt(fx), iy = srpdRoundToInt(fy);
			const long long idx = (long long) iy * W + ix;
			if (idx >= 0 && idx < total)
			{
				/* a fragment inside the row needs no 64-bit division: idx = iy*W + ix with 0 <= ix < W */
				const bool inRow = ix >= 0 && ix < W;
				const int px = inRow ? ix : (int) (idx % W), py = inRow ? iy : (int) (idx / W);
				x0 = px < x0 ? px : x0; x1 = px > x1 ? px : x1;
				y0 = py < y0 ? py : y0; y1 = py > y1 ? py : y1;
			}
			fx = SRP_FADD(fx, ln.xInc);
			fy = SRP_FADD(fy, ln.yInc);
		}
	}
	seg.any = x1 >= 0;
	seg.minX = (uint16_t) (seg.any ? x0 : 0); seg.maxX = (uint16_t) (seg.any ? x1 + 1 : 0);
	seg.minY = (uint16_t) (seg.any ? y0 : 0); seg.maxY = (uint16_t) (seg.any ? y1 + 1 : 0);
}

/* ---------------------------------------------------------------------------------
 * Point setup: setupPoint + computeMathAndRasterBoundaries, reference
 * src/raster/point.c:32-48,76-118.  Returns false if the point is entirely outside
 * the framebuffer (it still consumed its primitive id). */
struct SrpdPointSetup
{
	uint32_t w[SRPD_REC_HEADER_WORDS];
	uint16_t minX, minY, maxX, maxY;   /* half-open pixel bbox */
};

SRP_HD bool srpdSetupPoint(const SrpdState& st, const SrpdPos& clip, SrpdPointSetup& out)
{
	SrpdPos v = clip;
	srpdPerspectiveDivide(v);
	float sx, sy;
	srpdNdcToScreen(st, v, sx, sy);
	float half = (float) SRP_DMUL((double) st.pointSize, 0.5);
	float minBPx = SRP_FSUB(sx, half), minBPy = SRP_FSUB(sy, half);
	float maxBPx = SRP_FADD(sx, half), maxBPy = SRP_FADD(sy, half);
	int minX = (int) floor((double) minBPx), maxX = (int) floor((double) maxBPx);
	int minY = (int) floor((double) minBPy), maxY = (int) floor((double) maxBPy);
	if (maxX < 0 || maxY < 0 || minX >= st.width || minY >= st.height)
		return false;
	if (minX < 0) minX = 0;
	if (minY < 0) minY = 0;
	if (maxX >= st.width) maxX = st.width - 1;
	if (maxY >= st.height) maxY = st.height - 1;
	memset(out.w, 0, sizeof(out.w));
	out.w[0] = srpdF2U(minBPx); out.w[1] = srpdF2U(minBPy); out.w[2] = srpdF2U(maxBPx); out.w[3] = srpdF2U(maxBPy);
	out.w[4] = (uint32_t) minX; out.w[5] = (uint32_t) maxX; out.w[6] = (uint32_t) minY; out.w[7] = (uint32_t) maxY;
	out.w[8] = srpdF2U(v.z); out.w[9] = srpdF2U(v.w);
	out.minX = (uint16_t) minX; out.maxX = (uint16_t) (maxX + 1);
	out.minY = (uint16_t) minY; out.maxY = (uint16_t) (maxY + 1);
	return true;
}

/* clipPoint, reference clipping.c:185-197: true = rejected */
SRP_HD bool srpdClipPoint(const SrpdPos& p)
{
	if (p.x < -p.w || p.x > p.w) return true;
	if (p.y < -p.w || p.y > p.w) return true;
	if (p.z < -p.w || p.z > p.w) return true;
	return false;
}

/* ---------------------------------------------------------------------------------
 * Fragment stage helpers, reference src/raster/fragment.c:127-245, src/core/color.c:14-23 */
SRP_HD bool srpdCompare(uint8_t op, float a, float b)
{
	switch (op)
	{
		case SRP_COMPARE_NEVER:    return false;
		case SRP_COMPARE_ALWAYS:   return true;
		case SRP_COMPARE_LESS:     return a <  b;
		case SRP_COMPARE_LEQUAL:   return a <= b;
		case SRP_COMPARE_GREATER:  return a >  b;
		case SRP_COMPARE_GEQUAL:   return a >= b;
		case SRP_COMPARE_EQUAL:    return a == b;
		case SRP_COMPARE_NOTEQUAL: return a != b;
	}
	return false;
}
/* The same comparison without a switch per fragment: the operator becomes a 4-bit set of
 * accepted relations {a < b, a == b, a > b, unordered} (NOTEQUAL accepts unordered: `a != b`
 * is true for NaN; an unknown operator accepts nothing, like the reference's default). */
SRP_HD uint32_t srpdCompareMask(uint8_t op)
{
	return op < 8 ? (0xD26431F0u >> (4 * op)) & 0xFu : 0u;
}
SRP_HD bool srpdComparePass(uint32_t mask, float a, float b)
{
	const uint32_t rel = (a < b) ? 1u : ((a == b) ? 2u : ((a > b) ? 4u : 8u));
	return (mask & rel) != 0u;
}
SRP_HD bool srpdCompareU8(uint8_t op, uint8_t a, uint8_t b)
{
	switch (op)
	{
		case SRP_COMPARE_NEVER:    return false;
		case SRP_COMPARE_ALWAYS:   return true;
		case SRP_COMPARE_LESS:     return a <  b;
		case SRP_COMPARE_LEQUAL:   return a <= b;
		case SRP_COMPARE_GREATER:  return a >  b;
		case SRP_COMPARE_GEQUAL:   return a >= b;
		case SRP_COMPARE_EQUAL:    return a == b;
		case SRP_COMPARE_NOTEQUAL: return a != b;
	}
	return false;   /* unknown op: the reference falls back to NEVER (fragment.c:171-176) */
}
SRP_HD uint8_t srpdStencilOp(uint8_t op, uint8_t stored, uint8_t ref)
{
	switch (op)
	{
		case SRP_STENCIL_KEEP:      return stored;
		case SRP_STENCIL_ZERO:      return 0;
		case SRP_STENCIL_REPLACE:   return ref;
		case SRP_STENCIL_INCR:      return (uint8_t) ((stored < 255) ? stored + 1 : 255);
		case SRP_STENCIL_INCR_WRAP: return (uint8_t) (stored + 1);
		case SRP_STENCIL_DECR:      return (uint8_t) ((stored > 0) ? stored - 1 : 0);
		case SRP_STENCIL_DECR_WRAP: return (uint8_t) (stored - 1);
		case SRP_STENCIL_INVERT:    return (uint8_t) ~stored;
	}
	return stored;   /* unknown op: KEEP (fragment.c:216-221) */
}
SRP_HD uint8_t srpdStencilWrite(uint8_t current, uint8_t val, uint8_t writeMask)
{
	return (uint8_t) ((current & ~writeMask) | (val & writeMask));
}

/* colorPack: v = c*255 (float); <0 -> 0, >255 -> 255, else truncate; R<<24|G<<16|B<<8|A */
SRP_HD uint32_t srpdPackChannel(float c)
{
	float v = SRP_FMUL(c, 255.0f);
	/* v < 0 -> 0, v > 255 -> 255, else truncate; NaN is UB in the reference (unpinned): 0 here */
#ifdef __CUDA_ARCH__
	/* a float -> u8 conversion does exactly that: truncation, clamped to [0, 255], NaN -> 0 */
	uint32_t r;
	asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
	return r;
#else
	return (uint32_t) (int) fminf(fmaxf(v, 0.0f), 255.0f);   /* fmaxf returns its non-NaN argument */
#endif
}
SRP_HD uint32_t srpdColorPack(const float c[4])
{
	return (srpdPackChannel(c[0]) << 24) | (srpdPackChannel(c[1]) << 16) | (srpdPackChannel(c[2]) << 8) | srpdPackChannel(c[3]);
}

/* scissorTest, fragment.c:127-140: x, y arrive as int and are compared as size_t */
SRP_HD bool srpdScissor(const SrpdState& st, int x, int y)
{
	if (!st.scissorEnabled)
		return true;
	uint64_t ux = (uint64_t) (int64_t) x, uy = (uint64_t) (int64_t) y;
	return !(ux < st.scissorX0 || ux >= st.scissorX1 || uy < st.scissorY0 || uy >= st.scissorY1);
}

/* ---------------------------------------------------------------------------------
 * Varyings interpolation for one fragment, reference interpolation.c:93-163.
 * `blob[i]` are the record's blobs (PERSPECTIVE attributes already multiplied by iw_i),
 * `wgt[i]` the barycentric / line weights, `rec` the reciprocal of the interpolated 1/w.
 * Provoking vertex: FIRST -> 0, LAST -> n-1 in the order the rasteriser sees. */
template <int NV>
SRP_HD void srpdInterpolate(const SrpdState& st, const unsigned char* const blob[NV], const float wgt[NV],
                            float rec, unsigned char* out)
{
	const int prov = st.provokingFirst ? 0 : NV - 1;
	for (int ai = 0; ai < st.nVaryings; ai++)
	{
		const SrpdVarying& at = st.varyings[ai];
		const bool persp = at.mode == SRP_INTERPOLATION_MODE_PERSPECTIVE;
		const bool affine = at.mode == SRP_INTERPOLATION_MODE_AFFINE;
		if (srpdTypeIsFloat(at.type))
		{
			for (int e = 0; e < at.nItems; e++)
			{
				const int off = at.offset + 4 * e;
				float v = 0.f;
				if (persp || affine)
				{
					for (int i = 0; i < NV; i++)
						v = SRP_FADD(v, SRP_FMUL(srpdLoad<float>(blob[i] + off), wgt[i]));
					if (persp)
						v = SRP_FMUL(v, rec);
				}
				else
					v = srpdLoad<float>(blob[prov] + off);
				srpdStore<float>(out + off, v);
			}
		}
		else if (srpdTypeIsDouble(at.type))
		{
			for (int e = 0; e < at.nItems; e++)
			{
				const int off = at.offset + 8 * e;
				double v = 0.;
				if (persp || affine)
				{
					for (int i = 0; i < NV; i++)
						v = SRP_DADD(v, SRP_DMUL(srpdLoad<double>(blob[i] + off), (double) wgt[i]));
					if (persp)
						v = SRP_DMUL(v, (double) rec);
				}
				else
					v = srpdLoad<double>(blob[prov] + off);
				srpdStore<double>(out + off, v);
			}
		}
		else
		{
			const int n = at.nItems * at.elemSize;
			for (int k = 0; k < n; k++)
				out[at.offset + k] = blob[prov][at.offset + k];
		}
	}
}
