/* srp-b200 -- tile kernel: coverage, interpolation, fragment shading and the whole
 * per-fragment test sequence, one CTA per 32x16-pixel framebuffer tile (sm_100a).
 *
 * Replaces the reference's immediate-mode inner loops
 *   src/raster/triangle.c:73-111 (rasterizeTriangle), line.c:34-77 (rasterizeLine),
 *   point.c:32-74 (rasterizePoint), fragment.c:63-125 (emitFragment),
 *   pipeline/interpolation.c:34-163, core/color.c:14-23 (colorPack)
 * and srpFramebufferClear (core/framebuffer.c:57-62), which is fused in here.
 *
 * Ownership model: every thread owns TWO pixels of the tile (same column, rows ly and ly + 4
 * of its warp's 8x8 block) for the whole draw and keeps their colour / depth / stencil in
 * registers; the tile's primitives are visited in primitive-id order, so the reference's
 * "later primitive wins" semantics (no blending, depth EQUAL/ALWAYS, stencil counters) hold
 * with no atomics.  Work distribution (CTA = 8 warps = one 32x16 tile):
 *   - the CTA scans its candidate list (all records of the frame, or the coarse bin of
 *     its supertile) one record per thread at a time and compacts the ones whose bounding
 *     box touches the tile into shared memory with a ballot + warp reductions (order kept);
 *   - each warp walks that list 32 entries per step and keeps those touching its block;
 *   - triangles: "row lanes" -- one lane per (triangle, block row) -- walk the barycentric
 *     chain to their row and decide the coverage of its 8 pixels; the pixel threads then
 *     gather the bits of their own pixels and shade their covering triangles in order.
 * Exact arithmetic: a pixel's barycentrics are NOT evaluated in closed form; the reference's
 * incremental chain is replayed -- (y - minY) float additions of dlambda/dy from the value at
 * the bounding-box corner, then (x - minX) additions of dlambda/dx (triangle.c:102-109) --
 * which is what makes depth bit-exact.  Lines replay the DDA chain of
 * line.c:56-76 the same way.
 *
 * Framebuffer traffic per tile: at most one read and one write of 9 B/px; with a pending
 * clear no read at all.  Stores leave straight from the registers, every warp store as four
 * full 32-byte sectors (the eight lanes of a block row hold eight neighbouring pixels). */
#include "kernels.cuh"

namespace {

struct Pixel
{
	uint32_t color;
	float depth;
	uint32_t stencil;
	uint32_t dirty;      /* bit0 colour, bit1 depth, bit2 stencil */
};

struct FragCounters { uint32_t emitted, shaded; };

/* float varyings of one fragment, two per step; PAIRS > 0: compile-time trip count */
template <int NV, int PAIRS>
__device__ __forceinline__ void interpolateFloatPairs(
	const SrpdState& st, const float2* b, int slotPairs, int pairs, const float* wgt, float rec, float2* out)
{
	uint32_t modes = st.floatModes;
	const int n = PAIRS > 0 ? PAIRS : pairs;
	#pragma unroll
	for (int e = 0; e < n; e++, modes >>= 4)
	{
		float2 in[NV];
		#pragma unroll
		for (int i = 0; i < NV; i++)
			in[i] = __ldg(b + i * slotPairs + e);
		float vx = 0.f, vy = 0.f;
		#pragma unroll
		for (int i = 0; i < NV; i++)
		{
			vx = __fadd_rn(vx, __fmul_rn(in[i].x, wgt[i]));
			vy = __fadd_rn(vy, __fmul_rn(in[i].y, wgt[i]));
		}
		const uint32_t mx = modes & 3u, my = (modes >> 2) & 3u;
		if (mx == SRP_INTERPOLATION_MODE_PERSPECTIVE) vx = __fmul_rn(vx, rec);
		if (my == SRP_INTERPOLATION_MODE_PERSPECTIVE) vy = __fmul_rn(vy, rec);
		const float2 pv = st.provokingFirst ? in[0] : in[NV - 1];
		if (mx == SRP_INTERPOLATION_MODE_FLAT) vx = pv.x;
		if (my == SRP_INTERPOLATION_MODE_FLAT) vy = pv.y;
		out[e] = make_float2(vx, vy);
	}
}

/* emitFragment, reference src/raster/fragment.c:63-125.  `sx, sy` are the (unwrapped)
 * integer fragment coordinates the scissor test sees; interpolation of the varyings is
 * deferred until the early tests have passed (it is pure). */
/* SIMPLE (compile-time) = 1: no scissor, no stencil, the shader does not write depth and all
 * varyings are floats -- the state almost every draw has; the tests on the draw state then
 * disappear from the fragment stage instead of being evaluated per fragment. */
/* SIMPLE = 2: additionally every varying is PERSPECTIVE (what Gouraud colours and texture
 * coordinates are): the two mode bits per float are not decoded per fragment at all. */
template <int NV, int PAIRS>
__device__ __forceinline__ void interpolatePerspectivePairs(const float2* b, int slotPairs, const float* wgt, float rec, float2* out)
{
	#pragma unroll
	for (int e = 0; e < PAIRS; e++)
	{
		float2 in[NV];
		#pragma unroll
		for (int i = 0; i < NV; i++)
			in[i] = __ldg(b + i * slotPairs + e);
		float vx = 0.f, vy = 0.f;
		#pragma unroll
		for (int i = 0; i < NV; i++)
		{
			vx = __fadd_rn(vx, __fmul_rn(in[i].x, wgt[i]));
			vy = __fadd_rn(vy, __fmul_rn(in[i].y, wgt[i]));
		}
		out[e] = make_float2(__fmul_rn(vx, rec), __fmul_rn(vy, rec));
	}
}

template <int NV, int SIMPLE>
__device__ __forceinline__ void emitFragment(
	const SrpdState& st, const SrpdFrame& fr, Pixel& px, FragCounters& cnt,
	int sx, int sy, float fragX, float fragY, float depth, float rec, float fragW,
	bool frontFacing, uint32_t primitiveID,
	const unsigned char* blobs, const float* wgt)
{
	const bool scissorEnabled = SIMPLE ? false : (bool) st.scissorEnabled;
	const bool stencilEnabled = SIMPLE ? false : (bool) st.stencilEnabled;
	const bool earlyDepth = SIMPLE ? true : (bool) st.earlyDepth;
	const bool allFloat = SIMPLE ? true : (bool) st.allFloat;
	cnt.emitted++;
	if (scissorEnabled && !srpdScissor(st, sx, sy))
		return;

	const float storedDepth = px.depth;
	if (stencilEnabled)
	{
		const SrpdStencilFace& sf = frontFacing ? st.stencilFront : st.stencilBack;
		const uint8_t storedStencil = (uint8_t) px.stencil;
		if (!srpdCompareU8(sf.func, (uint8_t) (sf.ref & sf.mask), (uint8_t) (storedStencil & sf.mask)))
		{
			px.stencil = srpdStencilWrite(storedStencil, srpdStencilOp(sf.sfailOp, storedStencil, sf.ref), sf.writeMask);
			px.dirty |= 4u;
			return;
		}
	}
	if (earlyDepth && st.depthTest && !srpdComparePass(srpdCompareMask(st.depthOp), depth, storedDepth))
	{
		if (stencilEnabled)
		{
			const SrpdStencilFace& sf = frontFacing ? st.stencilFront : st.stencilBack;
			const uint8_t storedStencil = (uint8_t) px.stencil;
			px.stencil = srpdStencilWrite(storedStencil, srpdStencilOp(sf.dfailOp, storedStencil, sf.ref), sf.writeMask);
			px.dirty |= 4u;
		}
		return;
	}

	alignas(8) unsigned char interpolated[SRPD_MAX_VARYING_BYTES];
	if (NV == 1)
	{
		for (int k = 0; k < st.slotSize / 4; k++)
			((uint32_t*) interpolated)[k] = __ldg((const uint32_t*) blobs + k);
	}
	else if (allFloat)
	{
		/* all attributes are floats: the blob is an array of st.nFloats floats, two bits of
		 * interpolation mode each; same operation order as srpdInterpolate (interpolation.c:63-83).
		 * Blobs are 8-byte aligned and slotSize is a multiple of 8: two floats per load; the
		 * usual sizes are unrolled so that the mode bits are decoded once per draw, not per fragment. */
		const float2* b = (const float2*) blobs;
		float2* o = (float2*) interpolated;
		const int pairs = (st.nFloats + 1) >> 1, slotPairs = st.slotSize / 8;
		if (SIMPLE == 2)
			switch (pairs)      /* 1..4 (launchTileKernel) */
			{
				case 1:  interpolatePerspectivePairs<NV, 1>(b, slotPairs, wgt, rec, o); break;
				case 2:  interpolatePerspectivePairs<NV, 2>(b, slotPairs, wgt, rec, o); break;
				case 3:  interpolatePerspectivePairs<NV, 3>(b, slotPairs, wgt, rec, o); break;
				default: interpolatePerspectivePairs<NV, 4>(b, slotPairs, wgt, rec, o); break;
			}
		else
			switch (pairs)
			{
				case 1:  interpolateFloatPairs<NV, 1>(st, b, slotPairs, 1, wgt, rec, o); break;
				case 2:  interpolateFloatPairs<NV, 2>(st, b, slotPairs, 2, wgt, rec, o); break;
				case 3:  interpolateFloatPairs<NV, 3>(st, b, slotPairs, 3, wgt, rec, o); break;
				case 4:  interpolateFloatPairs<NV, 4>(st, b, slotPairs, 4, wgt, rec, o); break;
				default: interpolateFloatPairs<NV, 0>(st, b, slotPairs, pairs, wgt, rec, o); break;
			}
	}
	else
	{
		const unsigned char* b[NV];
		for (int i = 0; i < NV; i++)
			b[i] = blobs + i * st.slotSize;
		srpdInterpolate<NV>(st, b, wgt, rec, interpolated);
	}

	SRPFragmentShaderIn in;
	in.uniform = (SRPUniform*) fr.uniform;
	in.varyings = (SRPInterpolated*) interpolated;
	in.fragCoord[0] = fragX; in.fragCoord[1] = fragY; in.fragCoord[2] = depth; in.fragCoord[3] = fragW;
	in.frontFacing = frontFacing;
	in.primitiveID = primitiveID;
	SRPFragmentShaderOut out;
	out.color[0] = 0.f; out.color[1] = 0.f; out.color[2] = 0.f; out.color[3] = 0.f;
	out.fragDepth = __int_as_float(0x7FC00000);   /* NAN */
	srpB200DeviceFS(st.programId, &in, &out);
	cnt.shaded++;

	if (!earlyDepth)
	{
		if (!isnan(out.fragDepth))
			depth = out.fragDepth;
		if (st.depthTest && !srpdComparePass(srpdCompareMask(st.depthOp), depth, storedDepth))
		{
			if (stencilEnabled)
			{
				const SrpdStencilFace& sf = frontFacing ? st.stencilFront : st.stencilBack;
				const uint8_t storedStencil = (uint8_t) px.stencil;
				px.stencil = srpdStencilWrite(storedStencil, srpdStencilOp(sf.dfailOp, storedStencil, sf.ref), sf.writeMask);
				px.dirty |= 4u;
			}
			return;
		}
	}
	if (stencilEnabled)
	{
		const SrpdStencilFace& sf = frontFacing ? st.stencilFront : st.stencilBack;
		const uint8_t storedStencil = (uint8_t) px.stencil;
		px.stencil = srpdStencilWrite(storedStencil, srpdStencilOp(sf.passOp, storedStencil, sf.ref), sf.writeMask);
		px.dirty |= 4u;
	}
	px.color = srpdColorPack(out.color);
	px.dirty |= 1u;
	if (st.depthTest && st.depthWrite)
	{
		px.depth = depth;
		px.dirty |= 2u;
	}
}

/* Triangle coverage, one LANE per (TRIANGLE, BLOCK ROW).
 *
 * The reference reaches pixel (x, y) of a triangle by (y - minY) float additions of
 * dlambda/dy from the value at the bounding-box corner, then (x - minX) additions of
 * dlambda/dx (triangle.c:102-109).  The 8 pixels of a block row share that chain up to the
 * row's first pixel, and only a handful of the (up to 32) triangles of a list step touch a
 * warp's 8x4 block, so coverage is not decided by the pixel threads (most of which would
 * only find out that they are outside) but by "row lanes": lane 4*k + r takes row r of the
 * k-th touching triangle, walks the chain down to its row and right to the first column
 * (from the box corner, or from the (row, tile column) checkpoint of a large triangle),
 * tests the row's <= 8 pixels with the top-left rule and leaves
 *   - the row's 8 coverage bits (combined by shuffles into the triangle's 32-bit block mask),
 *   - lambda at the first column, for the pixel threads to resume from.
 * The pixel threads then transpose the masks (bit t of `cov` = triangle t covers my pixel)
 * and shade their own covered triangles in primitive order, resuming the chain with their
 * <= 7 remaining x steps: every value goes through exactly the reference's sequence of
 * additions => bit-exact. */
__device__ __forceinline__ uint32_t warpInclusiveScan(uint32_t v, int lane)
{
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, o);
		if (lane >= o) v += n;
	}
	return v;
}

/* A thread owns SRPD_PX pixels of its column (rows ly and ly + 4 of the warp's block).  Their
 * state lives in small arrays that are only ever indexed by compile-time constants, so they
 * stay in registers; a run-time choice between them goes through selects (getSel / setSel),
 * which lets the long fragment stage exist once in the code instead of once per pixel. */
template <typename T>
__device__ __forceinline__ T getSel(const T (&v)[SRPD_PX], int h)
{
	T r = v[0];
	#pragma unroll
	for (int i = 1; i < SRPD_PX; i++)
		if (h == i) r = v[i];
	return r;
}
template <typename T>
__device__ __forceinline__ void setSel(T (&v)[SRPD_PX], int h, const T& x)
{
	#pragma unroll
	for (int i = 0; i < SRPD_PX; i++)
		if (h == i) v[i] = x;
}
__device__ __forceinline__ int pickPending(const uint32_t (&cov)[SRPD_PX])
{
	int h = 0;
	#pragma unroll
	for (int i = SRPD_PX - 1; i > 0; i--)
		if (cov[i] != 0u && cov[0] == 0u) h = i;
	return h;
}

struct RowStart { float l0, l1, l2; int xs; };          /* lambda at column xs of the row          */
struct TriStep  { float dx0, dx1, dx2; uint32_t rec; };  /* dlambda/dx and the record slot          */

/* per-warp scratch of one list step (shared memory) */
struct WarpStep
{
	RowStart row[32 * SRPD_BLK_H];       /* [compact triangle][block row]                          */
	TriStep  tri[32];                    /* [compact triangle]                                     */
	uint32_t idBase[32];                 /* [compact triangle] id prefix of the triangle's batch   */
	alignas(16) uint8_t bits[SRPD_BLK_H * 32];   /* [block row][compact triangle]: the row's 8 coverage bits */
	uint8_t  pair[SRPD_BLK_H * 32];      /* work list of the row lanes: triangle * SRPD_BLK_H + row   */
};

/* top-left rule as ONE comparison per edge: the reference accepts lambda when
 * lambda > 0 || (|lambda| <= 1e-9 && edgeTL) (triangle.c:82-87).  With F = the largest float
 * <= 1e-9 (srpdRoughlyZero) that is lambda > 0 for a non-TL edge and lambda >= -F, i.e.
 * lambda > nextbelow(-F), for a TL edge; NaN fails both forms. */
__device__ __forceinline__ float coverageThreshold(bool topLeft)
{
	return topLeft ? __uint_as_float(0xB0897060u) : 0.0f;      /* 0xB089705F = -F; one ulp further from zero */
}

__device__ __forceinline__ uint32_t coverTriangleRow(
	const unsigned char* rec, const float* ckptTable, int bx0, int y, RowStart* rowOut, TriStep* triOut)
{
	const uint4* h = (const uint4*) rec;
	const uint4 q0 = __ldg(h + 0), q1 = __ldg(h + 1), q2 = __ldg(h + 2);
	const int minX = (int) (q0.w & 0xFFFFu), maxX = (int) (q0.w >> 16);
	const int minY = (int) (q1.w & 0xFFFFu), maxY = (int) (q1.w >> 16);
	const float dx0 = __uint_as_float(q1.x), dx1 = __uint_as_float(q1.y), dx2 = __uint_as_float(q1.z);
	triOut->dx0 = dx0; triOut->dx1 = dx1; triOut->dx2 = dx2;      /* (every row lane of the triangle writes the same values) */
	const int xs = max(bx0, minX);
	const int n = min(bx0 + SRPD_BLK_W, maxX) - xs;
	if (y < minY || y >= maxY || n <= 0)
		return 0u;
	float l0, l1, l2;
	int nx;
	const uint32_t ckpt = __ldg((const uint32_t*) rec + 19);
	if (ckpt)
	{
		/* large triangle: resume from the checkpoint of (row, this tile's column), written by
		 * srpdCheckpointKernel with the reference's own sequence of additions */
		const int col0 = minX / SRPD_TILE_W;
		const int cols = (maxX - 1) / SRPD_TILE_W - col0 + 1;
		const int tileX0 = (bx0 / SRPD_TILE_W) * SRPD_TILE_W;
		const float* e = ckptTable + 3 * ((size_t) (ckpt - 1) + (size_t) (y - minY) * cols + (bx0 / SRPD_TILE_W - col0));
		l0 = __ldg(e + 0); l1 = __ldg(e + 1); l2 = __ldg(e + 2);
		nx = xs - max(tileX0, minX);
	}
	else
	{
		l0 = __uint_as_float(q0.x); l1 = __uint_as_float(q0.y); l2 = __uint_as_float(q0.z);
		const float dy0 = __uint_as_float(q2.x), dy1 = __uint_as_float(q2.y), dy2 = __uint_as_float(q2.z);
		int ny = y - minY;
		for (; ny >= 4; ny -= 4)
		{
			#pragma unroll
			for (int u = 0; u < 4; u++)
			{
				l0 = __fadd_rn(l0, dy0); l1 = __fadd_rn(l1, dy1); l2 = __fadd_rn(l2, dy2);
			}
		}
		#pragma unroll
		for (int u = 0; u < 3; u++)
			if (u < ny)
			{
				l0 = __fadd_rn(l0, dy0); l1 = __fadd_rn(l1, dy1); l2 = __fadd_rn(l2, dy2);
			}
		nx = xs - minX;
	}
	for (; nx >= 4; nx -= 4)
	{
		#pragma unroll
		for (int u = 0; u < 4; u++)
		{
			l0 = __fadd_rn(l0, dx0); l1 = __fadd_rn(l1, dx1); l2 = __fadd_rn(l2, dx2);
		}
	}
	#pragma unroll
	for (int u = 0; u < 3; u++)
		if (u < nx)
		{
			l0 = __fadd_rn(l0, dx0); l1 = __fadd_rn(l1, dx1); l2 = __fadd_rn(l2, dx2);
		}
	rowOut->l0 = l0; rowOut->l1 = l1; rowOut->l2 = l2; rowOut->xs = xs;
	/* the row's pixels; values past the row's last pixel are computed but masked off */
	const uint32_t flags = q2.w;
	const float t0 = coverageThreshold(flags & 1u), t1 = coverageThreshold(flags & 2u), t2 = coverageThreshold(flags & 4u);
	uint32_t bits = 0u;
	#pragma unroll
	for (int i = 0; i < SRPD_BLK_W; i++)
	{
#ifdef SRPD_COVER_CHAINED_SETP
		/* experiment (leads for round 2): the three edge tests as one chain of
		 * predicated compares instead of three compares combined through selects */
		uint32_t bit;
		asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\tsetp.gt.and.f32 p, %3, %4, p;\n\tsetp.gt.and.f32 p, %5, %6, p;\n\t"
		    "selp.u32 %0, %7, 0, p;\n\t}"
		    : "=r"(bit) : "f"(l0), "f"(t0), "f"(l1), "f"(t1), "f"(l2), "f"(t2), "r"(1u << i));
		bits |= bit;
#else
		if (l0 > t0 && l1 > t1 && l2 > t2)
			bits |= 1u << i;
#endif
		if (i + 1 < SRPD_BLK_W)
		{
			l0 = __fadd_rn(l0, dx0); l1 = __fadd_rn(l1, dx1); l2 = __fadd_rn(l2, dx2);
		}
	}
	return (bits & ((1u << n) - 1u)) << (xs - bx0);
}

/* fragment stage of one covered pixel of a triangle: the pixel's remaining x steps, depth /
 * 1/w interpolation (interpolateDepthAndWTriangle, interpolation.c:34-47) and emitFragment */
template <int SIMPLE>
__device__ __forceinline__ void shadeTriangleFragment(
	const SrpdTileArgs& a, const SrpdFrame& fr, const unsigned char* records, const RowStart& rs, const TriStep& ts, uint32_t idBase,
	Pixel& px, FragCounters& cnt, int x, int y)
{
	float l0 = rs.l0, l1 = rs.l1, l2 = rs.l2;
	const int nx = x - rs.xs;      /* 0..7 */
#ifndef SRPD_NO_STEP_IN_REGS
	/* ONE 16-byte shared load of the triangle's step, pinned by `volatile`: left to itself the
	 * compiler, at the 48-register cap, re-loads it under every predicated step below (7 LDS.128
	 * per fragment) and spills around the loop; measured on cfg3: tiles 0.194 -> 0.186 ms */
	float dx0, dx1, dx2;
	uint32_t recSlot;
	asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=f"(dx0), "=f"(dx1), "=f"(dx2), "=r"(recSlot)
	             : "r"((uint32_t) __cvta_generic_to_shared(&ts)));
#else
	const float dx0 = ts.dx0, dx1 = ts.dx1, dx2 = ts.dx2;
	const uint32_t recSlot = ts.rec;
#endif
	#pragma unroll
	for (int i = 0; i < SRPD_BLK_W - 1; i++)
		if (i < nx)
		{
			l0 = __fadd_rn(l0, dx0); l1 = __fadd_rn(l1, dx1); l2 = __fadd_rn(l2, dx2);
		}
	const unsigned char* rec = records + (size_t) recSlot * a.recStride;
	const uint4* h = (const uint4*) rec;
	const uint4 q3 = __ldg(h + 3), q4 = __ldg(h + 4);
	const uint32_t flags = __ldg((const uint32_t*) rec + 11);
	const float wgt[3] = { l0, l1, l2 };
	const float iwSum = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(q4.x), l0), __fmul_rn(__uint_as_float(q4.y), l1)),
	                              __fmul_rn(__uint_as_float(q4.z), l2));
	const float recW = __fdiv_rn(1.0f, iwSum);
	const float depth = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(q3.x), l0), __fmul_rn(__uint_as_float(q3.y), l1)),
	                              __fmul_rn(__uint_as_float(q3.z), l2));
	/* pixel centre: (float) ((double) x + 0.5) is exact, and so is the float sum for these magnitudes */
	emitFragment<3, SIMPLE>(a.d.st, fr, px, cnt, x, y, __fadd_rn((float) x, 0.5f), __fadd_rn((float) y, 0.5f),
	                depth, recW, recW, (flags & 8u) != 0, q3.w + idBase, rec + SRPD_REC_HEADER_BYTES, wgt);
}

/* One list step of a warp.  `mine` = this lane's list entry (record slot `recSlot`) touches the
 * warp's block.  The touching entries are compacted in order (t = 0 .. n-1); row lanes decide
 * coverage, 8 triangles x 4 rows per round; pixel threads gather the bits of their pixel --
 * bit t of `cov` <=> triangle t covers my pixel -- and shade them in order. */
template <int SIMPLE>
__device__ __forceinline__ void visitTriangles(
	const SrpdTileArgs& a, const SrpdFrame& fr, const unsigned char* records, bool mine, uint32_t recSlot, uint32_t idBase, int rowLo, int rowCnt,
	WarpStep& ws, Pixel (&px)[SRPD_PX], FragCounters& cnt, int x, int y0, int bx0, int by0, int lane)
{
	const uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
	if (m == 0u)
		return;
	const int n = __popc(m);
	/* work list of (triangle, row) pairs: only the block rows inside a triangle's bounding box,
	 * so the row lanes are (nearly) all busy whatever the triangles' heights */
	const uint32_t inc = warpInclusiveScan(mine ? (uint32_t) rowCnt : 0u, lane);
	const int nPairs = (int) __shfl_sync(0xFFFFFFFFu, inc, 31);
	#pragma unroll
	for (int i = 0; i < SRPD_BLK_H * 32 / 4 / 32; i++)
		((uint32_t*) ws.bits)[lane + 32 * i] = 0u;
	if (mine)
	{
		const int t = __popc(m & ((1u << lane) - 1u));
		ws.tri[t].rec = recSlot;
		ws.idBase[t] = idBase;
		uint8_t* out = ws.pair + (inc - (uint32_t) rowCnt);
		for (int r = 0; r < rowCnt; r++)
			out[r] = (uint8_t) (t * SRPD_BLK_H + rowLo + r);
	}
	__syncwarp();
	for (int q = lane; q < nPairs; q += 32)
	{
		const int pr = ws.pair[q];
		const int t = pr / SRPD_BLK_H, row = pr % SRPD_BLK_H;
		const uint32_t bits = coverTriangleRow(records + (size_t) ws.tri[t].rec * a.recStride, a.ckptTable, bx0, by0 + row,
		                                       &ws.row[pr], &ws.tri[t]);
		ws.bits[row * 32 + t] = (uint8_t) bits;
	}
	__syncwarp();
	/* gather: byte t of a row's 32 bytes holds the row's coverage of triangle t; bit (lane % 8)
	 * of it is my pixel.  Four triangles per 32-bit word: isolate the bit in each byte, then
	 * one multiply moves the four bits next to each other (no carries: all partial products
	 * land on distinct bit positions). */
	const int ly = lane / SRPD_BLK_W, lx = lane % SRPD_BLK_W;
	uint32_t cov[SRPD_PX];
	#pragma unroll
	for (int h = 0; h < SRPD_PX; h++)
	{
		cov[h] = 0u;
		const uint32_t* rowBits = (const uint32_t*) (ws.bits + (ly + 4 * h) * 32);
		for (int w = 0; w * 4 < n; w++)
		{
			const uint32_t four = (rowBits[w] >> lx) & 0x01010101u;
			cov[h] |= (((four * 0x00204081u) >> 21) & 0xFu) << (4 * w);
		}
		if (n < 32)
			cov[h] &= (1u << n) - 1u;      /* bytes of triangles beyond n are stale */
	}
	/* shade: every lane works through the covering triangles of its own pixel(s), in order */
	for (;;)
	{
		uint32_t any = cov[0];
		#pragma unroll
		for (int h = 1; h < SRPD_PX; h++)
			any |= cov[h];
		if (!__any_sync(0xFFFFFFFFu, any != 0u))
			break;
		if (any)
		{
			const int h = pickPending(cov);
			const uint32_t c = getSel(cov, h);
			const int t = __ffs(c) - 1;
			setSel(cov, h, c & (c - 1u));
			Pixel cur = getSel(px, h);
			shadeTriangleFragment<SIMPLE>(a, fr, records, ws.row[t * SRPD_BLK_H + ly + 4 * h], ws.tri[t], ws.idBase[t], cur, cnt, x, y0 + 4 * h);
			setSel(px, h, cur);
		}
	}
	__syncwarp();      /* the next step overwrites this warp's scratch */
}

/* rasterizeLine for the warp's pixel block, reference line.c:34-77.  A record is a segment of
 * <= SRPD_LINE_SEG (16) consecutive DDA fragments with the chain state at its first one.  Lane k
 * walks the chain to fragment k -- k float additions per coordinate, exactly the reference's
 * sequence, all fragments side by side instead of every lane replaying all of them -- and rounds
 * it to its pixel once.  The fragments that land in this block are then matched to the lanes
 * that own their pixels, and every lane shades its own pixels' fragments in DDA order, all lanes
 * at once: a fragment is matched through its linear index
 * y*W + x, which is also how the reference's unchecked indexing wraps x == width onto the next
 * row (App. B-1).  Must be called by all 32 lanes. */
static_assert(SRPD_LINE_SEG <= 32, "one lane per fragment of a segment");
__device__ __forceinline__ void visitLine(
	const SrpdTileArgs& a, const SrpdFrame& fr, const unsigned char* rec, uint32_t idBase, Pixel (&px)[SRPD_PX], FragCounters& cnt,
	int x, int y0, const bool (&valid)[SRPD_PX])
{
	const int lane = threadIdx.x & 31;
	const uint4* h = (const uint4*) rec;
	const uint4 q0 = __ldg(h + 0), q1 = __ldg(h + 1), q2 = __ldg(h + 2), q3 = __ldg(h + 3);
	float fx = __uint_as_float(q0.x), fy = __uint_as_float(q0.y);
	const float xInc = __uint_as_float(q0.z), yInc = __uint_as_float(q0.w);
	const float tInc = __uint_as_float(q1.x);
	const int count = (int) q1.y;
	const float zw0 = __uint_as_float(q1.z), zw1 = __uint_as_float(q1.w);
	const float iw0 = __uint_as_float(q2.x), iw1 = __uint_as_float(q2.y);
	float t = __uint_as_float(q2.z);
	const long long W = a.d.st.width;
	long long mine[SRPD_PX];
	#pragma unroll
	for (int k = 0; k < SRPD_PX; k++)
		mine[k] = valid[k] ? (long long) (y0 + 4 * k) * W + x : -1;
	/* lane k: k steps of the chain */
	for (int i = 0; i + 1 < count; i++)
		if (i < lane)
		{
			fx = __fadd_rn(fx, xInc);
			fy = __fadd_rn(fy, yInc);
			t = __fadd_rn(t, tInc);
		}
	const int myX = srpdRoundToInt(fx), myY = srpdRoundToInt(fy);
	/* does my fragment land in this warp's block (columns bx0 .. bx0+7, rows by0 .. by0+7)? */
	bool inBlock = false;
	if (lane < count)
	{
		const long long idx = (long long) myY * W + myX;
		if (idx >= 0 && idx < W * (long long) a.d.st.height)
		{
			const int bx0 = x - (lane % SRPD_BLK_W), by0 = y0 - (lane / SRPD_BLK_W);
			const bool inRow = myX >= 0 && myX < W;      /* the usual case needs no 64-bit division */
			const int pxl = inRow ? myX : (int) (idx % W), pyl = inRow ? myY : (int) (idx / W);
			inBlock = pxl >= bx0 && pxl < bx0 + SRPD_BLK_W && pyl >= by0 && pyl < by0 + SRPD_BLK_H;
		}
	}
	/* which fragments hit my pixels?  bit k of hits[j] = fragment k lands on my j-th pixel */
	uint32_t hits[SRPD_PX];
	#pragma unroll
	for (int j = 0; j < SRPD_PX; j++)
		hits[j] = 0u;
	for (uint32_t m = __ballot_sync(0xFFFFFFFFu, inBlock); m != 0u; m &= m - 1u)
	{
		const int k = __ffs(m) - 1;
		const int ipx = __shfl_sync(0xFFFFFFFFu, myX, k), ipy = __shfl_sync(0xFFFFFFFFu, myY, k);
		const long long idx = (long long) ipy * W + ipx;
		#pragma unroll
		for (int j = 0; j < SRPD_PX; j++)
			if (idx == mine[j]) hits[j] |= 1u << k;
	}
	/* every lane shades the fragments of its own pixels, in DDA order per pixel; different
	 * pixels are independent, so the lanes work side by side (a line rarely visits a pixel twice) */
	for (;;)
	{
		uint32_t any = hits[0];
		#pragma unroll
		for (int j = 1; j < SRPD_PX; j++)
			any |= hits[j];
		if (!__any_sync(0xFFFFFFFFu, any != 0u))
			break;
		const int which = pickPending(hits);
		const uint32_t hsel = getSel(hits, which);
		const int k = any ? __ffs(hsel) - 1 : 0;
		const int ipx = __shfl_sync(0xFFFFFFFFu, myX, k), ipy = __shfl_sync(0xFFFFFFFFu, myY, k);
		const float tk = __shfl_sync(0xFFFFFFFFu, t, k);
		if (any)
		{
			setSel(hits, which, hsel & (hsel - 1u));
			const float w0 = __fsub_rn(1.0f, tk);
			const float wgt[2] = { w0, tk };
			/* interpolateDepthAndWLine, interpolation.c:49-60 */
			const float recW = __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(iw0, w0), __fmul_rn(iw1, tk)));
			const float depth = __fadd_rn(__fmul_rn(zw0, w0), __fmul_rn(zw1, tk));
			Pixel cur = getSel(px, which);
			emitFragment<2, 0>(a.d.st, fr, cur, cnt, ipx, ipy, (float) ((double) ipx + 0.5), (float) ((double) ipy + 0.5),
			                depth, recW, recW, true, q3.w + idBase, rec + SRPD_REC_HEADER_BYTES, wgt);
			setSel(px, which, cur);
		}
	}
}

/* rasterizePoint for the thread's pixels, reference point.c:32-74 */
__device__ __forceinline__ void visitPoint(
	const SrpdTileArgs& a, const SrpdFrame& fr, const unsigned char* rec, uint32_t idBase, Pixel (&px)[SRPD_PX], FragCounters& cnt,
	int x, int y0, const bool (&valid)[SRPD_PX])
{
	const uint4* h = (const uint4*) rec;
	const uint4 q1 = __ldg(h + 1);
	if (x < (int) q1.x || x > (int) q1.y)
		return;
	const uint4 q0 = __ldg(h + 0);
	const float pcx = (float) ((double) x + 0.5);
	if (pcx < __uint_as_float(q0.x) || pcx >= __uint_as_float(q0.z))
		return;
	#pragma unroll 1
	for (int k = 0; k < SRPD_PX; k++)
	{
		const int y = y0 + 4 * k;
		if (!getSel(valid, k) || y < (int) q1.z || y > (int) q1.w)
			continue;
		const float pcy = (float) ((double) y + 0.5);
		if (pcy < __uint_as_float(q0.y) || pcy >= __uint_as_float(q0.w))
			continue;
		const uint4 q2 = __ldg(h + 2), q3 = __ldg(h + 3);
		Pixel cur = getSel(px, k);
		emitFragment<1, 0>(a.d.st, fr, cur, cnt, x, y, pcx, pcy, __uint_as_float(q2.x), 0.f, __uint_as_float(q2.y),
		                true, q3.w + idBase, rec + SRPD_REC_HEADER_BYTES, nullptr);
		setSel(px, k, cur);
	}
}

} // namespace

/* shared memory of the tile kernel (dynamic: with the per-warp step scratch it exceeds 48 KB) */
struct TileShared
{
	uint32_t ids[SRPD_TILE_THREADS];     /* record slots of the tile's list chunk */
	uint32_t idBase[SRPD_TILE_THREADS];  /* id prefix of their batches (records hold batch-local ids) */
	uint2    box[SRPD_TILE_THREADS];     /* their boxes */
	uint32_t warpCnt[32];
	uint32_t item[2];
};
template <int KIND> struct TileSharedK : TileShared {};
template <> struct TileSharedK<SRPD_KIND_TRIANGLE> : TileShared { WarpStep step[SRPD_TILE_WARPS]; };

/* One tile: filter the candidate list, visit the primitives in order, write the tile back. */
template <int KIND, int SIMPLE>
__device__ __forceinline__ void processTile(
	const SrpdTileArgs& a, const SrpdFrame& fr, uint32_t frame, int tileX, int tileY, TileSharedK<KIND>& sm, FragCounters& cnt)
{
	const SrpdState& st = a.d.st;
	const bool stencilEnabled = SIMPLE ? false : (bool) st.stencilEnabled;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

	/* candidate list of this tile */
	uint32_t begin = 0, end;
	const uint32_t* ids = nullptr;
	if (a.superOffsets && *a.listOverflow == 0u)
	{
		const uint32_t s = ((uint32_t) tileY >> a.superShift) * a.superX + ((uint32_t) tileX >> a.superShift);
		begin = a.superOffsets[s];
		end = a.superOffsets[s + 1];
		ids = a.listIds;
	}
	else
		end = a.frameCounts[2 * frame + 1];

	const unsigned char* records = a.records + (size_t) frame * a.recCapacity * a.recStride;
	const uint4* ordered = a.ordered + (size_t) frame * a.recCapacity;

	/* pixel ownership: warp w -> block (w % 4, w / 4) of 8 x SRPD_BLK_H pixels;
	 * lane -> column lane % 8, rows lane / 8 (+ 4 for the thread's second pixel) */
	const int tx0 = tileX * SRPD_TILE_W, ty0 = tileY * SRPD_TILE_H;
	const int bx0 = tx0 + (warp % (SRPD_TILE_W / SRPD_BLK_W)) * SRPD_BLK_W;
	const int by0 = ty0 + (warp / (SRPD_TILE_W / SRPD_BLK_W)) * SRPD_BLK_H;
	const int x = bx0 + (lane % SRPD_BLK_W), y0 = by0 + (lane / SRPD_BLK_W);

	Pixel px[SRPD_PX];
	bool valid[SRPD_PX];
	#pragma unroll
	for (int k = 0; k < SRPD_PX; k++)
	{
		const int y = y0 + 4 * k;
		valid[k] = x < st.width && y < st.height;
		px[k].color = 0u; px[k].depth = -1.0f; px[k].stencil = 0u; px[k].dirty = 0u;
		if (valid[k] && (!fr.clearPending || stencilEnabled))
		{
			const size_t pixelIndex = (size_t) y * st.width + x;
			if (!fr.clearPending)
			{
				px[k].color = fr.color[pixelIndex];
				if (st.depthTest)
					px[k].depth = fr.depth[pixelIndex];
			}
			if (stencilEnabled)
				px[k].stencil = fr.stencil[pixelIndex];
		}
	}

	for (uint32_t c = begin; c < end; c += SRPD_TILE_THREADS)
	{
		/* CTA: keep the candidates whose bbox touches the tile, in order */
		const uint32_t i = c + tid;
		bool hit = false;
		uint4 ent = make_uint4(0u, 0u, 0u, 0u);
		uint2 bb = make_uint2(0u, 0u);
		if (i < end)
		{
			const uint32_t rid = ids ? ids[i] : i;          /* position in primitive order */
			ent = ordered[rid];                              /* {box, record slot, id prefix} */
			bb = make_uint2(ent.x, ent.y);
			const int x0 = (int) (bb.x & 0xFFFFu), y0b = (int) (bb.x >> 16);
			const int x1 = (int) (bb.y & 0xFFFFu), y1b = (int) (bb.y >> 16);
			hit = x0 < tx0 + SRPD_TILE_W && x1 > tx0 && y0b < ty0 + SRPD_TILE_H && y1b > ty0;
		}
		const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
		/* (no barrier needed here: whoever gets this far has passed the previous chunk's second
		 * barrier, which every warp reaches only after it has read that chunk's counters; and the
		 * list itself is rewritten only after the next barrier, which every warp reaches only
		 * after it has finished walking the previous list) */
		if (lane == 0)
			sm.warpCnt[warp] = __popc(ballot);
		__syncthreads();
		const uint32_t wc = lane < SRPD_TILE_WARPS ? sm.warpCnt[lane] : 0u;
		const uint32_t total = __reduce_add_sync(0xFFFFFFFFu, wc);
		const uint32_t base = __reduce_add_sync(0xFFFFFFFFu, lane < warp ? wc : 0u);
		if (hit)
		{
			const uint32_t pos = base + __popc(ballot & ((1u << lane) - 1u));
			sm.ids[pos] = ent.z;             /* record slot */
			sm.idBase[pos] = ent.w;
			sm.box[pos] = bb;
		}
		__syncthreads();

		/* warp: visit, in order, the entries that touch this warp's block */
		for (uint32_t j0 = 0; j0 < total; j0 += 32)
		{
			const uint32_t j = j0 + lane;
			bool mine = false;
			uint32_t slot = 0u, idBase = 0u;
			int rowLo = 0, rowCnt = 0;
			if (j < total)
			{
				const uint2 b2 = sm.box[j];
				const int x0 = (int) (b2.x & 0xFFFFu), y0b = (int) (b2.x >> 16);
				const int x1 = (int) (b2.y & 0xFFFFu), y1b = (int) (b2.y >> 16);
				mine = x0 < bx0 + SRPD_BLK_W && x1 > bx0 && y0b < by0 + SRPD_BLK_H && y1b > by0;
				slot = sm.ids[j];
				idBase = sm.idBase[j];
				rowLo = max(y0b, by0) - by0;
				rowCnt = min(y1b, by0 + SRPD_BLK_H) - by0 - rowLo;
			}
			if constexpr (KIND == SRPD_KIND_TRIANGLE)
				visitTriangles<SIMPLE>(a, fr, records, mine, slot, idBase, rowLo, rowCnt, sm.step[warp], px, cnt, x, y0, bx0, by0, lane);
			else
			{
				uint32_t m = __ballot_sync(0xFFFFFFFFu, mine);
				while (m)
				{
					const int bit = __ffs(m) - 1;
					m &= m - 1;
					const unsigned char* rec = records + (size_t) sm.ids[j0 + bit] * a.recStride;
					const uint32_t recIdBase = sm.idBase[j0 + bit];
					if (KIND == SRPD_KIND_LINE)
						visitLine(a, fr, rec, recIdBase, px, cnt, x, y0, valid);
					else
						visitPoint(a, fr, rec, recIdBase, px, cnt, x, y0, valid);
				}
			}
		}
	}

	/* write-back, straight from the registers: the eight lanes of a block row hold eight
	 * neighbouring pixels, so every store instruction of the warp leaves as four full 32-byte
	 * sectors (one per row); no staging in shared memory and no barrier -- a warp that is done
	 * with its block moves on to the next tile's list.  With a pending clear every pixel is
	 * written (untouched ones with the clear values), otherwise only what a fragment changed. */
	#pragma unroll
	for (int k = 0; k < SRPD_PX; k++)
	{
		if (!valid[k])
			continue;
		const size_t pixelIndex = (size_t) (y0 + 4 * k) * st.width + x;
		if (fr.clearPending || (px[k].dirty & 1u))
			fr.color[pixelIndex] = px[k].color;
		if (fr.clearPending || (px[k].dirty & 2u))
			fr.depth[pixelIndex] = px[k].depth;
		if (px[k].dirty & 4u)
			fr.stencil[pixelIndex] = (uint8_t) px[k].stencil;
	}
}

/* A tile no primitive touches while a clear is pending: just write the clear values
 * (colour 0, depth -1; reference core/framebuffer.c:57-62), one 128-byte row per warp. */
__device__ __forceinline__ void clearTile(const SrpdState& st, const SrpdFrame& fr, int tileX, int tileY)
{
	const int x = tileX * SRPD_TILE_W + (threadIdx.x % SRPD_TILE_W);
	#pragma unroll
	for (int k = 0; k < SRPD_PX; k++)
	{
		const int y = tileY * SRPD_TILE_H + (threadIdx.x / SRPD_TILE_W) + k * (SRPD_TILE_H / SRPD_PX);
		if (x < st.width && y < st.height)
		{
			const size_t i = (size_t) y * st.width + x;
			fr.color[i] = 0u;
			fr.depth[i] = -1.0f;
		}
	}
}

/* Persistent tile kernel: the grid is sized to the machine (CTAs per SM x SM count) and the
 * CTAs pull work items -- groups of `tilesPerItem` consecutive tiles of one frame -- from an
 * atomic counter, so neither empty tiles (skipped through the occupancy bitmap the geometry
 * kernel filled) nor hundreds of frames of a batch cost a CTA launch each.
 *
 * Register budget: both launch-bound arguments are given explicitly (under device LTO a
 * missing minimum makes the linker's code generator cap the kernel at 64 registers and
 * spill).  Triangles and points fit two 512-thread CTAs per SM; the line walker does not. */
template <int KIND, bool BATCH, int SIMPLE>
__global__ void __launch_bounds__(SRPD_TILE_THREADS, KIND == SRPD_KIND_LINE ? SRPD_TILE_LINE_CTAS_PER_SM : SRPD_TILE_CTAS_PER_SM)
srpdTileKernel(const __grid_constant__ SrpdTileArgs a)
{
	extern __shared__ __align__(16) unsigned char srpdTileSmem[];
	TileSharedK<KIND>& sm = *reinterpret_cast<TileSharedK<KIND>*>(srpdTileSmem);

	srpdGridDependencyEnter();
	if (*a.abortFlag)      /* a scratch pool overflowed: the host repeats the draw with larger pools */
		return;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t rows = a.d.tileRow1 - a.d.tileRow0;
	const uint32_t tilesPerFrame = a.tilesX * rows;
	const uint32_t itemsPerFrame = (tilesPerFrame + a.tilesPerItem - 1) / a.tilesPerItem;
	const uint32_t nItems = itemsPerFrame * a.d.nFrames;

	FragCounters cnt;
	cnt.emitted = 0; cnt.shaded = 0;

	if (tid == 0)
		sm.item[0] = atomicAdd(a.workCounter, 1u);
	__syncthreads();
	for (uint32_t it = 0;; it++)
	{
		const uint32_t item = sm.item[it & 1];
		if (item >= nItems)
			break;
		if (tid == 0)      /* fetch the next item while this one is processed */
			sm.item[(it + 1) & 1] = atomicAdd(a.workCounter, 1u);
		/* BATCH: many frames, bindings in a device array; otherwise the one frame of the argument block */
		uint32_t frame = 0u, inFrame = item;
		if (BATCH)
		{
			frame = item / itemsPerFrame;
			inFrame = item - frame * itemsPerFrame;
		}
		const uint32_t first = inFrame * a.tilesPerItem;
		const uint32_t last = min(first + a.tilesPerItem, tilesPerFrame);
		SrpdFrame frCopy;
		if (BATCH)
			frCopy = a.frames[frame];
		else
		{
			frCopy = a.frame0;
			frCopy.uniform = a.uniformInline;      /* constant bank (kernels.cuh) */
		}
		const SrpdFrame& fr = frCopy;
		const uint32_t* occ = a.occupancy + (size_t) frame * a.occWordsPerFrame;
		/* first / tilesX by multiplication: tilesXInv = floor(2^40 / tilesX) + 1 is exact while
		 * first * tilesX < 2^40 (tiles per frame < 2^23, tilesX <= 2^11) */
		uint32_t rowInBand = (uint32_t) (((uint64_t) first * a.tilesXInv) >> 40);
		int tileX = (int) (first - rowInBand * a.tilesX);
		int tileY = (int) (a.d.tileRow0 + rowInBand);
		for (uint32_t t = first; t < last; t++)
		{
			const uint32_t tileIndex = (uint32_t) tileY * a.tilesX + (uint32_t) tileX;
			const bool occupied = (occ[tileIndex >> 5] >> (tileIndex & 31u)) & 1u;
			if (occupied)
				processTile<KIND, SIMPLE>(a, fr, frame, tileX, tileY, sm, cnt);
			else if (fr.clearPending)
				clearTile(a.d.st, fr, tileX, tileY);
			if (++tileX == (int) a.tilesX)
			{
				tileX = 0;
				tileY++;
			}
		}
		__syncthreads();   /* item[(it + 1) & 1] is visible; item[it & 1] may be overwritten next round */
	}

	/* counters: warp reduction, then one atomic per warp into one of the slots, so the atomics
	 * of a frame do not serialise on one L2 address */
	{
		const uint32_t e = __reduce_add_sync(0xFFFFFFFFu, cnt.emitted), s = __reduce_add_sync(0xFFFFFFFFu, cnt.shaded);
		if (lane == 0 && e)
		{
			SrpdStats* slot = a.stats + ((blockIdx.x * SRPD_TILE_WARPS + warp) & (SRPD_STATS_SLOTS - 1));
			atomicAdd(&slot->fragsEmitted, (unsigned long long) e);
			atomicAdd(&slot->fragsShaded, (unsigned long long) s);
		}
	}
}

/* srpFramebufferClear as a real memory operation (only needed when a pending clear has
 * to be materialised without a draw), reference core/framebuffer.c:57-62 */
__global__ void __launch_bounds__(256) srpdClearKernel(uint4* color, uint4* depth, size_t nVec, uint32_t* colorTail, float* depthTail, int nTail)
{
	srpdGridDependencyEnter();
	const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
	const uint32_t m1 = 0xBF800000u;   /* -1.0f */
	const uint4 minusOne = make_uint4(m1, m1, m1, m1);
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < nVec; i += (size_t) gridDim.x * blockDim.x)
	{
		color[i] = zero;
		depth[i] = minusOne;
	}
	if (blockIdx.x == 0 && (int) threadIdx.x < nTail)
	{
		colorTail[threadIdx.x] = 0u;
		depthTail[threadIdx.x] = -1.0f;
	}
}

void srpdLaunchClear(uint32_t* color, float* depth, size_t nPixels, cudaStream_t stream)
{
	const size_t nVec = nPixels / 4;
	const int nTail = (int) (nPixels % 4);
	unsigned grid = (unsigned) ((nVec + 255) / 256);
	if (grid > 148u * 16u) grid = 148u * 16u;
	if (grid == 0) grid = 1;
	srpdClearKernel<<<grid, 256, 0, stream>>>((uint4*) color, (uint4*) depth, nVec, color + nVec * 4, depth + nVec * 4, nTail);
}

template <int KIND, bool BATCH, int SIMPLE>
static void launchTileKernelS(const SrpdTileArgs& a, unsigned grid, cudaStream_t stream)
{
	static bool configured = false;
	const int bytes = (int) sizeof(TileSharedK<KIND>);
	if (!configured)
	{
		cudaFuncSetAttribute(srpdTileKernel<KIND, BATCH, SIMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
		configured = true;
	}
	srpdLaunchKernel(srpdTileKernel<KIND, BATCH, SIMPLE>, grid, SRPD_TILE_THREADS, (size_t) bytes, stream, a);
}
template <int KIND>
static void launchTileKernel(const SrpdTileArgs& a, unsigned grid, cudaStream_t stream)
{
	const SrpdState& st = a.d.st;
	const bool simple = KIND == SRPD_KIND_TRIANGLE && !st.scissorEnabled && !st.stencilEnabled && st.earlyDepth && st.allFloat;
	if constexpr (KIND == SRPD_KIND_TRIANGLE)
	{
		if (simple)
		{
			/* all floats PERSPECTIVE (mode 0, two bits each) and at most four pairs of them */
			const uint32_t used = st.nFloats >= 16 ? 0xFFFFFFFFu : ((1u << (2 * st.nFloats)) - 1u);
			const bool perspective = st.nFloats >= 1 && st.nFloats <= 8 && (st.floatModes & used) == 0u
				&& SRP_INTERPOLATION_MODE_PERSPECTIVE == 0;
			if (perspective)
			{
				if (a.frames) launchTileKernelS<KIND, true, 2>(a, grid, stream);
				else          launchTileKernelS<KIND, false, 2>(a, grid, stream);
			}
			else
			{
				if (a.frames) launchTileKernelS<KIND, true, 1>(a, grid, stream);
				else          launchTileKernelS<KIND, false, 1>(a, grid, stream);
			}
			return;
		}
	}
	if (a.frames) launchTileKernelS<KIND, true, 0>(a, grid, stream);
	else          launchTileKernelS<KIND, false, 0>(a, grid, stream);
}

void srpdLaunchTiles(const SrpdTileArgs& a0, cudaStream_t stream)
{
	SrpdTileArgs a = a0;
	const uint32_t rows = a.d.tileRow1 - a.d.tileRow0;
	if (rows == 0 || a.tilesX == 0)
		return;
	a.tilesXInv = (1ull << 40) / a.tilesX + 1ull;
	/* persistent grid: resident CTAs per SM x SM count (no more CTAs than work items) */
	const uint32_t tilesPerFrame = a.tilesX * rows;
	const uint64_t nItems = (uint64_t) ((tilesPerFrame + a.tilesPerItem - 1) / a.tilesPerItem) * a.d.nFrames;
	const uint32_t perSm = a.d.kind == SRPD_KIND_LINE ? SRPD_TILE_LINE_CTAS_PER_SM : SRPD_TILE_CTAS_PER_SM;
	uint64_t grid = (uint64_t) a.smCount * perSm;
	if (grid > nItems) grid = nItems;
	if (grid == 0) return;
	if (a.d.kind == SRPD_KIND_TRIANGLE)
		launchTileKernel<SRPD_KIND_TRIANGLE>(a, (unsigned) grid, stream);
	else if (a.d.kind == SRPD_KIND_LINE)
		launchTileKernel<SRPD_KIND_LINE>(a, (unsigned) grid, stream);
	else
		launchTileKernel<SRPD_KIND_POINT>(a, (unsigned) grid, stream);
}
