/* srp-b200 -- tile kernel: coverage, interpolation, fragment shading and the whole
 * per-fragment test sequence (sm_100a).
 *
 * Replaces the reference's immediate-mode inner loops
 *   src/raster/triangle.c:73-111 (rasterizeTriangle), line.c:34-77 (rasterizeLine),
 *   point.c:32-74 (rasterizePoint), fragment.c:63-125 (emitFragment),
 *   pipeline/interpolation.c:34-163, core/color.c:14-23 (colorPack)
 * and srpFramebufferClear (core/framebuffer.c:57-62), which is fused in here.
 *
 * ONE WARP = ONE WARP TILE of 32x8 pixels.  The warp tile's colour / depth / stencil live in the
 * warp's SHARED MEMORY for as long as it works on the tile: loaded once (or, with a pending
 * clear, just initialised), tested and updated there by every fragment of the draw, written
 * back once with 16-byte stores in full 128-byte rows.  The warp is the only one that touches
 * these pixels, and it visits the tile's primitives in primitive-id order, so the reference's
 * "later primitive wins" semantics (no blending, depth EQUAL / ALWAYS, stencil counters) hold
 * with no atomics; the warps of a CTA share nothing and there is no CTA barrier.  Per tile:
 *   - LIST: the warp scans the candidate list (the coarse bin of its supertile, or all records)
 *     32 entries at a time and appends those whose bounding box touches the tile to a ring in
 *     shared memory (ballot compaction, order kept); 32 collected entries make a step;
 *   - COVERAGE (triangles): "row lanes" -- one lane per (triangle, tile row) -- walk the
 *     barycentric chain to their row and along it.  Each edge function moves monotonically
 *     along the row (float addition of a constant is monotone), so a row's covered pixels are ONE
 *     interval: the lane walks to its first covered pixel, counts to its last, and a warp scan
 *     gives every lane its place in the warp's FRAGMENT QUEUE, where it leaves
 *     {lambda0, lambda1, lambda2, triangle, pixel} per covered pixel (queue = primitive order);
 *   - SHADING: the lanes take the queue 32 fragments at a time -- whichever pixels they belong
 *     to, so every lane has a fragment -- and run the fragment stage on the pixel's state in
 *     shared memory; two fragments of one pass that hit the same pixel (different triangles)
 *     are found with match.any and run one after the other, in queue order.
 * Exact arithmetic: a pixel's barycentrics are NOT evaluated in closed form; the reference's
 * incremental chain is replayed -- (y - minY) float additions of dlambda/dy from the value at
 * the bounding-box corner, then (x - minX) additions of dlambda/dx (triangle.c:102-109) --
 * which is what makes depth bit-exact; the row lane that decides coverage
 * is the one that walks the chain, so no pixel replays any part of it.  Lines replay the DDA
 * chain of line.c:56-76 the same way.
 *
 * Framebuffer traffic per tile: at most one read and one write of 9 B/px; with a pending
 * clear no read at all.  Algorithmic bytes of a launch: 9 B x the pixels of the frame(s). */
#include "kernels.cuh"

namespace {

struct FragCounters { uint32_t emitted, shaded; };

/* the pixel a fragment lands on: its three plane entries in the tile's shared-memory state */
struct PixelRef
{
	uint32_t* color;
	float* depth;
	uint8_t* stencil;
};

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

/* all-PERSPECTIVE floats (what Gouraud colours and texture coordinates are): no mode decode.
 * QUADS: the blobs are 16 bytes apart and 16-byte aligned (1-4 floats per vertex in a 16-byte
 * slot: records are 16-byte aligned and the header is 80 bytes), one 16-byte load per vertex */
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
template <int NV, int FLOATS>
__device__ __forceinline__ void interpolatePerspectiveQuad(const float4* b, const float* wgt, float rec, float* out)
{
	float4 in[NV];
	#pragma unroll
	for (int i = 0; i < NV; i++)
		in[i] = __ldg(b + i);
	float v[4] = { 0.f, 0.f, 0.f, 0.f };
	#pragma unroll
	for (int i = 0; i < NV; i++)
	{
		v[0] = __fadd_rn(v[0], __fmul_rn(in[i].x, wgt[i]));
		v[1] = __fadd_rn(v[1], __fmul_rn(in[i].y, wgt[i]));
		v[2] = __fadd_rn(v[2], __fmul_rn(in[i].z, wgt[i]));
		if (FLOATS == 4)
			v[3] = __fadd_rn(v[3], __fmul_rn(in[i].w, wgt[i]));
	}
	#pragma unroll
	for (int e = 0; e < FLOATS; e++)
		out[e] = __fmul_rn(v[e], rec);
}

/* emitFragment, reference src/raster/fragment.c:63-125.  `sx, sy` are the (unwrapped)
 * integer fragment coordinates the scissor test sees; interpolation of the varyings is
 * deferred until the early tests have passed (it is pure).
 * SIMPLE (compile-time) = 1: no scissor, no stencil, the shader does not write depth and all
 * varyings are floats -- the state almost every draw has; the tests on the draw state then
 * disappear from the fragment stage instead of being evaluated per fragment.
 * SIMPLE = 2: additionally every varying is PERSPECTIVE: the two mode bits per float are not
 * decoded per fragment at all.
 * `dirty` collects which planes the thread changed (bit0 colour, bit1 depth, bit2 stencil). */
template <int NV, int SIMPLE>
__device__ __forceinline__ void emitFragment(
	const SrpdState& st, const SrpdFrame& fr, const PixelRef& px, uint32_t& dirty, FragCounters& cnt,
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

	const float storedDepth = *px.depth;
	if (stencilEnabled)
	{
		const SrpdStencilFace& sf = frontFacing ? st.stencilFront : st.stencilBack;
		const uint8_t storedStencil = *px.stencil;
		if (!srpdCompareU8(sf.func, (uint8_t) (sf.ref & sf.mask), (uint8_t) (storedStencil & sf.mask)))
		{
			*px.stencil = srpdStencilWrite(storedStencil, srpdStencilOp(sf.sfailOp, storedStencil, sf.ref), sf.writeMask);
			dirty |= 4u;
			return;
		}
	}
	if (earlyDepth && st.depthTest && !srpdComparePass(srpdCompareMask(st.depthOp), depth, storedDepth))
	{
		if (stencilEnabled)
		{
			const SrpdStencilFace& sf = frontFacing ? st.stencilFront : st.stencilBack;
			const uint8_t storedStencil = *px.stencil;
			*px.stencil = srpdStencilWrite(storedStencil, srpdStencilOp(sf.dfailOp, storedStencil, sf.ref), sf.writeMask);
			dirty |= 4u;
		}
		return;
	}

	alignas(16) unsigned char interpolated[SRPD_MAX_VARYING_BYTES];
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
		{
			if (st.slotSize == 16)      /* 3 or 4 floats (launchTileKernel: 1..8 floats); warp-uniform */
			{
				if (st.nFloats == 3) interpolatePerspectiveQuad<NV, 3>((const float4*) blobs, wgt, rec, (float*) interpolated);
				else                 interpolatePerspectiveQuad<NV, 4>((const float4*) blobs, wgt, rec, (float*) interpolated);
			}
			else
				switch (pairs)
				{
					case 1:  interpolatePerspectivePairs<NV, 1>(b, slotPairs, wgt, rec, o); break;
					case 2:  interpolatePerspectivePairs<NV, 2>(b, slotPairs, wgt, rec, o); break;
					case 3:  interpolatePerspectivePairs<NV, 3>(b, slotPairs, wgt, rec, o); break;
					default: interpolatePerspectivePairs<NV, 4>(b, slotPairs, wgt, rec, o); break;
				}
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
	srpB200DeviceFS(st.fsProgramId, &in, &out);
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
				const uint8_t storedStencil = *px.stencil;
				*px.stencil = srpdStencilWrite(storedStencil, srpdStencilOp(sf.dfailOp, storedStencil, sf.ref), sf.writeMask);
				dirty |= 4u;
			}
			return;
		}
	}
	if (stencilEnabled)
	{
		const SrpdStencilFace& sf = frontFacing ? st.stencilFront : st.stencilBack;
		const uint8_t storedStencil = *px.stencil;
		*px.stencil = srpdStencilWrite(storedStencil, srpdStencilOp(sf.passOp, storedStencil, sf.ref), sf.writeMask);
		dirty |= 4u;
	}
	*px.color = srpdColorPack(out.color);
	dirty |= 1u;
	if (st.depthTest && st.depthWrite)
	{
		*px.depth = depth;
		dirty |= 2u;
	}
}

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

/* ---- shared memory: one of these per warp ------------------------------------------- */

/* what the shading lanes need of a triangle, copied from its record once per list step:
 * a = {z0*iw0, z1*iw1, z2*iw2, primitive id}, b = {iw0, iw1, iw2, record slot | checkpointed << 30 | frontFacing << 31} */
struct TriInfo { uint4 a, b; };
constexpr uint32_t SRPD_TRI_SLOT_MASK = 0x3FFFFFFFu;      /* record slots are < 2^30 (runtime.cu clamps the pools) */
constexpr uint32_t SRPD_TRI_CKPT_BIT = 1u << 30, SRPD_TRI_FRONT_BIT = 1u << 31;

constexpr int SRPD_QUEUE = 128;          /* fragment queue entries */
constexpr int SRPD_RING = 64;            /* list entries waiting for a step (< 32 left over + <= 32 new) */

/* the warp tile's pixels: plane entry of pixel (x, y), x = 0..31, y = 0..7.  Rows are 32 words;
 * the 4-pixel (16-byte) chunks of a row are XOR-swizzled with the row so that the pixels of one
 * column -- what the fragments of a small triangle in one shading pass are made of -- fall into
 * different banks, while a chunk stays a chunk for the 16-byte loads / stores of the write-back */
__device__ __forceinline__ int pixelEntry(int x, int y)
{
	return y * SRPD_WT_W + ((((x >> 2) ^ y) & 7) << 2) + (x & 3);      /* = CU_TENSOR_MAP_SWIZZLE_128B of a 128-byte row */
}

/* Shared memory of a CTA: first the colour and depth planes of its warp tiles, 1024 bytes each and
 * 1024-byte aligned (the 128-byte swizzle of the TMA tile store works on 1024-byte atoms), then
 * one scratch block per warp.  The kernels work through these views. */
struct WarpScratch
{
	alignas(16) uint8_t stencil[SRPD_WT_PIXELS];     /* row-major, 32 bytes per row */
	uint4 ring[SRPD_RING];                           /* {box.x, box.y, record slot, id prefix} of the tile's next primitives */
};
struct WarpScratchTri : WarpScratch
{
	uint4   frag[SRPD_QUEUE];            /* fragment queue: lambda0..2 (float bits), triangle | x << 5 | y << 10 */
	TriInfo tri[32];                     /* [triangle of the step] */
	uint8_t pair[SRPD_WT_H * 32];        /* work list of the row lanes: triangle * 8 + row */
};
template <int KIND> struct WarpScratchK { typedef WarpScratch type; };
template <> struct WarpScratchK<SRPD_KIND_TRIANGLE> { typedef WarpScratchTri type; };
constexpr int SRPD_PLANE_BYTES = SRPD_WT_PIXELS * 4;      /* 1024 */

struct WarpTile
{
	uint32_t* color;      /* swizzled: pixelEntry(x, y) */
	float*    depth;      /* swizzled: pixelEntry(x, y) */
	uint8_t*  stencil;    /* y * 32 + x */
	uint4*    ring;
};
struct WarpTileTri : WarpTile
{
	uint4*   frag;
	TriInfo* tri;
	uint8_t* pair;
};
template <int KIND> struct WarpTileK { typedef WarpTile type; };
template <> struct WarpTileK<SRPD_KIND_TRIANGLE> { typedef WarpTileTri type; };

__device__ __forceinline__ int stencilEntry(int x, int y) { return y * SRPD_WT_W + x; }

/* top-left rule as ONE comparison per edge: the reference accepts lambda when
 * lambda > 0 || (|lambda| <= 1e-9 && edgeTL) (triangle.c:82-87).  With F = the largest float
 * <= 1e-9 (srpdRoughlyZero) that is lambda > 0 for a non-TL edge and lambda >= -F, i.e.
 * lambda > nextbelow(-F), for a TL edge; NaN fails both forms. */
__device__ __forceinline__ float coverageThreshold(bool topLeft)
{
	return topLeft ? __uint_as_float(0xB0897060u) : 0.0f;      /* 0xB089705F = -F; one ulp further from zero */
}

/* A row lane's walk: the chain to the row's first pixel inside the tile -- from the box corner,
 * or from the (row, tile column) checkpoint of a large triangle -- then along the row to its
 * first covered pixel and on to its last.  The reference tests every pixel of the box row
 * (triangle.c:80-110); lambda_i(x) is a chain of float additions of one constant, hence monotone
 * in x, hence {x : lambda_i(x) passes} is a prefix or a suffix of the row for each edge and the
 * covered pixels are one interval: once coverage has begun and ended nothing further can be
 * covered, so the walk stops there.  Returns the number of covered pixels; `first` = the first
 * one (tile-relative x), (s0, s1, s2) = lambda there, (dx0, dx1, dx2) = dlambda/dx. */
__device__ __forceinline__ int coverTriangleRow(
	const unsigned char* rec, bool checkpointed, const float* ckptTable, int tx0, int y,
	float& s0, float& s1, float& s2, float& dx0, float& dx1, float& dx2, int& first)
{
	const uint4* h = (const uint4*) rec;
	const uint4 q0 = __ldg(h + 0), q1 = __ldg(h + 1), q2 = __ldg(h + 2);
	const int minX = (int) (q0.w & 0xFFFFu), maxX = (int) (q0.w >> 16);
	const int minY = (int) (q1.w & 0xFFFFu);
	dx0 = __uint_as_float(q1.x); dx1 = __uint_as_float(q1.y); dx2 = __uint_as_float(q1.z);
	const int xs = max(tx0, minX);
	const int n = min(tx0 + SRPD_WT_W, maxX) - xs;      /* > 0: the box touches the tile (and y is inside the box) */
	float l0, l1, l2;
	int nx;
	if (checkpointed)
	{
		/* large triangle: resume from the checkpoint of (row, this tile column), written by
		 * srpdCheckpointKernel with the reference's own sequence of additions */
		const uint32_t ckpt = __ldg((const uint32_t*) rec + 19);
		const int col0 = minX / SRPD_TILE_W;
		const int cols = (maxX - 1) / SRPD_TILE_W - col0 + 1;
		const float* e = ckptTable + 3 * ((size_t) (ckpt - 1) + (size_t) (y - minY) * cols + (tx0 / SRPD_TILE_W - col0));
		l0 = __ldg(e + 0); l1 = __ldg(e + 1); l2 = __ldg(e + 2);
		nx = 0;      /* the checkpoint is at max(tile column start, minX) = xs */
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
	for (; nx > 0; nx--)      /* (only a box that starts left of the tile) */
	{
		l0 = __fadd_rn(l0, dx0); l1 = __fadd_rn(l1, dx1); l2 = __fadd_rn(l2, dx2);
	}
	const uint32_t flags = q2.w;
	const float t0 = coverageThreshold(flags & 1u), t1 = coverageThreshold(flags & 2u), t2 = coverageThreshold(flags & 4u);
	/* to the first covered pixel */
	int i = 0;
	while (i < n && !(l0 > t0 && l1 > t1 && l2 > t2))
	{
		l0 = __fadd_rn(l0, dx0); l1 = __fadd_rn(l1, dx1); l2 = __fadd_rn(l2, dx2);
		i++;
	}
	first = xs - tx0 + i;
	s0 = l0; s1 = l1; s2 = l2;
	/* on to the last one */
	int count = 0;
	while (i < n && (l0 > t0 && l1 > t1 && l2 > t2))
	{
		l0 = __fadd_rn(l0, dx0); l1 = __fadd_rn(l1, dx1); l2 = __fadd_rn(l2, dx2);
		i++;
		count++;
	}
	return count;
}

/* Two fragments of one shading pass that land on the same pixel must run one after the other,
 * in queue order (the queue is in primitive order).  `key` = the pixel for lanes that hold a
 * fragment, a value no other lane has otherwise.  Returns this lane's turn (0 = first) and,
 * through `turns`, how many turns the pass needs (1 unless pixels collide). */
__device__ __forceinline__ uint32_t fragmentTurn(uint32_t key, int lane, uint32_t& turns)
{
	const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
	const uint32_t turn = __popc(peers & ((1u << lane) - 1u));
	turns = __reduce_max_sync(0xFFFFFFFFu, turn) + 1u;
	return turn;
}

/* fragment stage of one queued triangle fragment: depth / 1/w interpolation
 * (interpolateDepthAndWTriangle, interpolation.c:34-47) and emitFragment */
template <int SIMPLE>
__device__ __forceinline__ void shadeTriangleFragment(
	const SrpdTileArgs& a, const SrpdFrame& fr, const unsigned char* records, const uint4& frag, const TriInfo& ti,
	WarpTile& wt, int tx0, int ty0, uint32_t& dirty, FragCounters& cnt)
{
	const float l0 = __uint_as_float(frag.x), l1 = __uint_as_float(frag.y), l2 = __uint_as_float(frag.z);
	const int lx = (int) ((frag.w >> 5) & 31u), ly = (int) (frag.w >> 10);
	const uint4 A = ti.a, B = ti.b;
	const float wgt[3] = { l0, l1, l2 };
	const float iwSum = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(B.x), l0), __fmul_rn(__uint_as_float(B.y), l1)),
	                              __fmul_rn(__uint_as_float(B.z), l2));
	const float recW = __fdiv_rn(1.0f, iwSum);
	const float depth = __fadd_rn(__fadd_rn(__fmul_rn(__uint_as_float(A.x), l0), __fmul_rn(__uint_as_float(A.y), l1)),
	                              __fmul_rn(__uint_as_float(A.z), l2));
	const int x = tx0 + lx, y = ty0 + ly;
	const unsigned char* rec = records + (size_t) (B.w & SRPD_TRI_SLOT_MASK) * a.recStride;
	const int p = pixelEntry(lx, ly);
	PixelRef px;
	px.color = wt.color + p; px.depth = wt.depth + p; px.stencil = wt.stencil + stencilEntry(lx, ly);
	/* pixel centre: (float) ((double) x + 0.5) is exact, and so is the float sum for these magnitudes */
	emitFragment<3, SIMPLE>(a.d.st, fr, px, dirty, cnt, x, y, __fadd_rn((float) x, 0.5f), __fadd_rn((float) y, 0.5f),
	                depth, recW, recW, (B.w & SRPD_TRI_FRONT_BIT) != 0u, A.w, rec + SRPD_REC_HEADER_BYTES, wgt);
}

/* one shading pass: queue entries [f0, f0 + 32) */
template <int SIMPLE>
__device__ __forceinline__ void shadePass(
	const SrpdTileArgs& a, const SrpdFrame& fr, const unsigned char* records, WarpTileTri& wt, int f0, int nFrags, bool several,
	int tx0, int ty0, uint32_t& dirty, FragCounters& cnt, int lane)
{
	const int f = f0 + lane;
	const bool has = f < nFrags;
	uint4 frag = make_uint4(0u, 0u, 0u, 0u);
	if (has)
		frag = wt.frag[f];
	uint32_t turn = 0u, turns = 1u;
	if (several)
		turn = fragmentTurn(has ? (frag.w >> 5) : (uint32_t) (SRPD_WT_PIXELS + lane), lane, turns);
	for (uint32_t r = 0; r < turns; r++)
	{
		if (has && turn == r)
			shadeTriangleFragment<SIMPLE>(a, fr, records, frag, wt.tri[frag.w & 31u], wt, tx0, ty0, dirty, cnt);
		if (turns > 1u)
			__syncwarp();
	}
	if (several)
		__syncwarp();      /* a later pass may hold another triangle's fragment for one of this pass's pixels */
}

/* One step: the next `n` (<= 32) primitives of the tile, entries ring[head ..] (all of them touch
 * the tile).  Their shading data goes to shared memory; the (triangle, row) pairs -- only the tile
 * rows inside a triangle's box -- form a work list that the row lanes take 32 at a time:
 * coverage (coverTriangleRow), then a warp scan places every lane's fragments in the queue, in
 * (triangle, row, x) order, SRPD_QUEUE fragments per turn; the shading passes empty the queue;
 * lanes whose pixels did not fit take the next turn. */
template <int SIMPLE>
__device__ __forceinline__ void visitTriangles(
	const SrpdTileArgs& a, const SrpdFrame& fr, const unsigned char* records, WarpTileTri& wt, int head, int n,
	int tx0, int ty0, uint32_t& dirty, FragCounters& cnt, int lane)
{
	int rowLo = 0, rowCnt = 0;
	if (lane < n)
	{
		const uint4 ent = wt.ring[(head + lane) & (SRPD_RING - 1)];
		const int y0b = (int) (ent.x >> 16), y1b = (int) (ent.y >> 16);
		rowLo = max(y0b, ty0) - ty0;
		rowCnt = min(y1b, ty0 + SRPD_WT_H) - ty0 - rowLo;
		const unsigned char* rec = records + (size_t) ent.z * a.recStride;
		const uint4 q3 = __ldg((const uint4*) rec + 3), q4 = __ldg((const uint4*) rec + 4);
		const uint32_t flags = __ldg((const uint32_t*) rec + 11);
		TriInfo ti;
		ti.a = make_uint4(q3.x, q3.y, q3.z, q3.w + ent.w);      /* batch-local id + the batch's id prefix */
		ti.b = make_uint4(q4.x, q4.y, q4.z, ent.z | (q4.w ? SRPD_TRI_CKPT_BIT : 0u) | ((flags & 8u) ? SRPD_TRI_FRONT_BIT : 0u));
		wt.tri[lane] = ti;
	}
	const uint32_t inc = warpInclusiveScan((uint32_t) rowCnt, lane);
	const int nPairs = (int) __shfl_sync(0xFFFFFFFFu, inc, 31);
	{
		uint8_t* out = wt.pair + (inc - (uint32_t) rowCnt);
		for (int r = 0; r < rowCnt; r++)
			out[r] = (uint8_t) (lane * SRPD_WT_H + rowLo + r);
	}
	__syncwarp();
	int carried = 0;      /* fragments at the front of the queue that wait for a full pass (< 32) */
	for (int q0 = 0; q0 < nPairs; q0 += 32)
	{
		/* coverage: one lane per (triangle, row) */
		const int q = q0 + lane;
		uint32_t tRow = 0u;
		float l0 = 0.f, l1 = 0.f, l2 = 0.f, dx0 = 0.f, dx1 = 0.f, dx2 = 0.f;
		int x = 0, left = 0;
		if (q < nPairs)
		{
			tRow = wt.pair[q];
			const uint32_t slotWord = wt.tri[tRow / SRPD_WT_H].b.w;
			left = coverTriangleRow(records + (size_t) (slotWord & SRPD_TRI_SLOT_MASK) * a.recStride, (slotWord & SRPD_TRI_CKPT_BIT) != 0u,
			                        a.ckptTable, tx0, ty0 + (int) (tRow % SRPD_WT_H), l0, l1, l2, dx0, dx1, dx2, x);
		}
		/* do two triangles meet in this round?  (only then can two fragments share a pixel) */
		const uint32_t firstTri = __shfl_sync(0xFFFFFFFFu, tRow / SRPD_WT_H, 0);
		const bool several = carried > 0 || __any_sync(0xFFFFFFFFu, left != 0 && tRow / SRPD_WT_H != firstTri);
		uint32_t meta = (tRow / SRPD_WT_H) | ((uint32_t) x << 5) | ((tRow % SRPD_WT_H) << 10);
		while (__any_sync(0xFFFFFFFFu, left != 0))
		{
			/* a turn: the next fragments in strict (lane, x) = (primitive, row, x) order, as many as
			 * the queue holds: the lanes in front queue all they have left, one lane may get in
			 * only a part of its row, the lanes behind it wait for the next turn (a lane behind
			 * may belong to a later triangle that covers the same pixels) */
			const uint32_t want = (uint32_t) left;
			const uint32_t incF = warpInclusiveScan(want, lane);
			const uint32_t excl = incF - want;
			const int room = SRPD_QUEUE - carried;
			const int total = (int) __shfl_sync(0xFFFFFFFFu, incF, 31);
			const int nFrags = carried + min(total, room);
			const int k = excl >= (uint32_t) room ? 0 : min((int) want, room - (int) excl);
			uint4* out = wt.frag + carried + excl;
			for (int j = 0; j < k; j++)
			{
				/* lambda at the covered pixels: the chain goes on with the same additions */
				out[j] = make_uint4(__float_as_uint(l0), __float_as_uint(l1), __float_as_uint(l2), meta);
				l0 = __fadd_rn(l0, dx0); l1 = __fadd_rn(l1, dx1); l2 = __fadd_rn(l2, dx2);
				meta += 1u << 5;
			}
			left -= k;
			__syncwarp();
			/* full passes only while more fragments of this step are to come: what is left over
			 * (< 32) moves to the front of the queue and is shaded with the next turn's fragments,
			 * so that a pass has 32 fragments whichever rows and triangles they come from */
			const bool more = total > room || q0 + 32 < nPairs;
			const int nShade = more ? (nFrags & ~31) : nFrags;
			for (int f0 = 0; f0 < nShade; f0 += 32)
				shadePass<SIMPLE>(a, fr, records, wt, f0, nShade, several, tx0, ty0, dirty, cnt, lane);
			__syncwarp();      /* the next turn overwrites the queue */
			carried = nFrags - nShade;
			if (carried > 0 && nShade > 0)
			{
				uint4 keep = make_uint4(0u, 0u, 0u, 0u);
				if (lane < carried)
					keep = wt.frag[nShade + lane];
				__syncwarp();
				if (lane < carried)
					wt.frag[lane] = keep;
				__syncwarp();
			}
		}
	}
	if (carried > 0)      /* (the last rounds of the step had no fragments of their own) */
	{
		shadePass<SIMPLE>(a, fr, records, wt, 0, carried, true, tx0, ty0, dirty, cnt, lane);
		__syncwarp();
	}
	__syncwarp();      /* the next step overwrites the triangle data and the pair list */
}

/* rasterizeLine for the warp tile, reference line.c:34-77.  A record is a segment of
 * <= SRPD_LINE_SEG (16) consecutive DDA fragments with the chain state at its first one.  A step
 * is up to 32 records; their fragments -- 1 to 16 each -- are numbered in (record, DDA) order by a
 * warp scan over the records' fragment counts and taken 32 per pass, ONE LANE PER FRAGMENT
 * whichever records they belong to (a million two-pixel lines fill the warp as well as a few long
 * ones).  A lane finds its fragment's record by bisection over the scan, walks the chain to its
 * fragment -- k float additions per coordinate, exactly the reference's sequence -- and rounds it
 * to its pixel once.  A fragment that lands in this warp's tile is shaded by the lane that
 * walked to it, on the pixel's state in shared memory; a fragment is placed through its linear
 * index y*W + x, which is also how the reference's unchecked indexing wraps x == width onto the
 * next row (App. B-1).  Fragments of a pass that share a pixel run in lane order = (record, DDA)
 * order.  Must be called by all 32 lanes. */
__device__ __forceinline__ void visitLines(
	const SrpdTileArgs& a, const SrpdFrame& fr, const unsigned char* records, WarpTile& wt, int head, int n,
	int tx0, int ty0, uint32_t& dirty, FragCounters& cnt, int lane)
{
	uint32_t myCount = 0u;
	if (lane < n)
		myCount = __ldg((const uint32_t*) (records + (size_t) wt.ring[(head + lane) & (SRPD_RING - 1)].z * a.recStride) + 5);
	uint32_t incl = myCount;
	#pragma unroll
	for (int d = 1; d < 32; d <<= 1)
	{
		const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
		if (lane >= d) incl += v;
	}
	const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
	const long long W = a.d.st.width;
	for (uint32_t f0 = 0; f0 < total; f0 += 32u)
	{
		const uint32_t f = f0 + (uint32_t) lane;
		const bool have = f < total;
		/* the record of fragment f: the first one whose inclusive count exceeds f */
		int e = 0;
		#pragma unroll
		for (int step = 16; step > 0; step >>= 1)
		{
			const uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, e + step - 1);
			if (v <= f) e += step;
		}
		e = min(e, 31);
		const uint32_t inclE = __shfl_sync(0xFFFFFFFFu, incl, e), countE = __shfl_sync(0xFFFFFFFFu, myCount, e);
		bool inTile = false;
		int lx = 0, ly = 0, myX = 0, myY = 0;
		float t = 0.f;
		uint4 q1 = make_uint4(0u, 0u, 0u, 0u), q2 = q1, q3 = q1;
		const unsigned char* rec = nullptr;
		uint32_t idBase = 0u;
		if (have)
		{
			const int k = (int) (f - (inclE - countE));      /* my fragment's place in its segment */
			const uint4 ent = wt.ring[(head + e) & (SRPD_RING - 1)];
			rec = records + (size_t) ent.z * a.recStride;
			idBase = ent.w;
			const uint4* h = (const uint4*) rec;
			const uint4 q0 = __ldg(h + 0);
			q1 = __ldg(h + 1); q2 = __ldg(h + 2); q3 = __ldg(h + 3);
			float fx = __uint_as_float(q0.x), fy = __uint_as_float(q0.y);
			const float xInc = __uint_as_float(q0.z), yInc = __uint_as_float(q0.w);
			const float tInc = __uint_as_float(q1.x);
			t = __uint_as_float(q2.z);
			for (int i = 0; i < k; i++)
			{
				fx = __fadd_rn(fx, xInc);
				fy = __fadd_rn(fy, yInc);
				t = __fadd_rn(t, tInc);
			}
			myX = srpdRoundToInt(fx); myY = srpdRoundToInt(fy);
			/* does my fragment land in this warp's tile? */
			const long long idx = (long long) myY * W + myX;
			if (idx >= 0 && idx < W * (long long) a.d.st.height)
			{
				const bool inRow = myX >= 0 && myX < W;      /* the usual case needs no 64-bit division */
				const int pxl = inRow ? myX : (int) (idx % W), pyl = inRow ? myY : (int) (idx / W);
				inTile = pxl >= tx0 && pxl < tx0 + SRPD_WT_W && pyl >= ty0 && pyl < ty0 + SRPD_WT_H;
				lx = pxl - tx0; ly = pyl - ty0;
			}
		}
		if (!__any_sync(0xFFFFFFFFu, inTile))
			continue;
		uint32_t turns;
		const uint32_t turn = fragmentTurn(inTile ? (uint32_t) (ly * SRPD_WT_W + lx) : (uint32_t) (SRPD_WT_PIXELS + lane), lane, turns);
		for (uint32_t r = 0; r < turns; r++)
		{
			if (inTile && turn == r)
			{
				const float zw0 = __uint_as_float(q1.z), zw1 = __uint_as_float(q1.w);
				const float iw0 = __uint_as_float(q2.x), iw1 = __uint_as_float(q2.y);
				const float w0 = __fsub_rn(1.0f, t);
				const float wgt[2] = { w0, t };
				/* interpolateDepthAndWLine, interpolation.c:49-60 */
				const float recW = __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(iw0, w0), __fmul_rn(iw1, t)));
				const float depth = __fadd_rn(__fmul_rn(zw0, w0), __fmul_rn(zw1, t));
				const int p = pixelEntry(lx, ly);
				PixelRef px;
				px.color = wt.color + p; px.depth = wt.depth + p; px.stencil = wt.stencil + stencilEntry(lx, ly);
				emitFragment<2, 0>(a.d.st, fr, px, dirty, cnt, myX, myY, (float) ((double) myX + 0.5), (float) ((double) myY + 0.5),
				                depth, recW, recW, true, q3.w + idBase, rec + SRPD_REC_HEADER_BYTES, wgt);
			}
			if (turns > 1u)
				__syncwarp();
		}
		__syncwarp();      /* the next pass sees these fragments' pixels */
	}
}

/* rasterizePoint for the warp tile, reference point.c:32-74: a step of up to 32 points, ONE LANE
 * PER POINT.  A lane walks the pixels of its point's square inside the tile and runs the fragment
 * stage on each; points whose squares share no pixel are independent, so the lanes work side by
 * side.  Order where squares do overlap: every lane first learns which EARLIER points of the
 * step overlap its square (one shuffle of the packed pixel box per point); a lane runs once all
 * of those have run -- waves of mutually independent points, in primitive order per pixel. */
__device__ __forceinline__ void visitPoints(
	const SrpdTileArgs& a, const SrpdFrame& fr, const unsigned char* records, WarpTile& wt, int head, int n,
	int tx0, int ty0, uint32_t& dirty, FragCounters& cnt, int lane)
{
	const unsigned char* rec = nullptr;
	uint32_t idBase = 0u;
	int xLo = 1, xHi = 0, yLo = 1, yHi = 0;      /* empty */
	uint4 q0 = make_uint4(0u, 0u, 0u, 0u);
	if (lane < n)
	{
		const uint4 ent = wt.ring[(head + lane) & (SRPD_RING - 1)];
		rec = records + (size_t) ent.z * a.recStride;
		idBase = ent.w;
		const uint4* h = (const uint4*) rec;
		q0 = __ldg(h + 0);                                   /* the square: minX, minY, maxX, maxY (floats, half-open) */
		const uint4 q1 = __ldg(h + 1);                        /* pixel box, inclusive: minX, maxX, minY, maxY */
		xLo = max((int) q1.x, tx0); xHi = min(min((int) q1.y, tx0 + SRPD_WT_W - 1), a.d.st.width - 1);
		yLo = max((int) q1.z, ty0); yHi = min(min((int) q1.w, ty0 + SRPD_WT_H - 1), a.d.st.height - 1);
	}
	const bool any = xLo <= xHi && yLo <= yHi;
	/* tile-relative box in one word: x0 | x1 << 8 | y0 << 16 | y1 << 24 */
	const uint32_t box = any ? (uint32_t) (xLo - tx0) | ((uint32_t) (xHi - tx0) << 8) | ((uint32_t) (yLo - ty0) << 16) | ((uint32_t) (yHi - ty0) << 24) : 0xFFFFFFFFu;
	uint32_t blockers = 0u;      /* earlier points of the step whose squares overlap mine */
	for (int e = 0; e + 1 < n; e++)
	{
		const uint32_t other = __shfl_sync(0xFFFFFFFFu, box, e);
		if (e < lane && any && other != 0xFFFFFFFFu)
		{
			const int ox0 = (int) (other & 255u), ox1 = (int) ((other >> 8) & 255u), oy0 = (int) ((other >> 16) & 255u), oy1 = (int) (other >> 24);
			if (ox0 <= xHi - tx0 && ox1 >= xLo - tx0 && oy0 <= yHi - ty0 && oy1 >= yLo - ty0)
				blockers |= 1u << e;
		}
	}
	uint32_t pending = __ballot_sync(0xFFFFFFFFu, any);
	while (pending != 0u)
	{
		const bool go = ((pending >> lane) & 1u) && (blockers & pending) == 0u;
		if (go)
		{
			const uint4 q2 = __ldg((const uint4*) rec + 2), q3 = __ldg((const uint4*) rec + 3);
			for (int y = yLo; y <= yHi; y++)
			{
				const float pcy = (float) ((double) y + 0.5);
				if (pcy < __uint_as_float(q0.y) || pcy >= __uint_as_float(q0.w))
					continue;
				for (int x = xLo; x <= xHi; x++)
				{
					const float pcx = (float) ((double) x + 0.5);
					if (pcx < __uint_as_float(q0.x) || pcx >= __uint_as_float(q0.z))
						continue;
					const int p = pixelEntry(x - tx0, y - ty0);
					PixelRef px;
					px.color = wt.color + p; px.depth = wt.depth + p; px.stencil = wt.stencil + stencilEntry(x - tx0, y - ty0);
					emitFragment<1, 0>(a.d.st, fr, px, dirty, cnt, x, y, pcx, pcy, __uint_as_float(q2.x), 0.f, __uint_as_float(q2.y),
					                true, q3.w + idBase, rec + SRPD_REC_HEADER_BYTES, nullptr);
				}
			}
		}
		pending &= ~__ballot_sync(0xFFFFFFFFu, go);
		__syncwarp();      /* the next wave sees this one's pixels */
	}
}

/* ---- warp tile: global memory <-> shared memory ---------------------------------------
 * A tile row is 32 pixels = 128 bytes of colour / depth = eight 16-byte chunks of 4 pixels and
 * 32 bytes of stencil.  Lane l moves chunks l and l + 32 of the 64 chunks of a plane (chunk c =
 * row c / 8, pixels 4 * (c % 8) ..), so eight neighbouring lanes move one full 128-byte row.
 * The 16-byte path needs the row pitch and the plane base to keep the chunks aligned (width a
 * multiple of 4 pixels, 16 for stencil); other framebuffers take the element-wise path. */
struct TileIO
{
	bool vec4;       /* colour / depth rows can move as 16-byte chunks */
	bool vec16;      /* stencil rows can move as 16-byte chunks */
};

__device__ __forceinline__ TileIO tileIO(const SrpdState& st, const SrpdFrame& fr)
{
	TileIO io;
	io.vec4 = (st.width % 4) == 0 && ((((uintptr_t) fr.color) | ((uintptr_t) fr.depth)) & 15u) == 0;
	io.vec16 = (st.width % 16) == 0 && (((uintptr_t) fr.stencil) & 15u) == 0;
	return io;
}

/* the planes the draw reads (or, with a pending clear, the clear values) -> shared memory */
__device__ __forceinline__ void loadWarpTile(const SrpdState& st, const SrpdFrame& fr, const TileIO& io, bool stencilEnabled,
                                             WarpTile& wt, int tx0, int ty0, int lane)
{
	const bool loadColor = !fr.clearPending, loadDepth = !fr.clearPending && st.depthTest;
	#pragma unroll
	for (int k = 0; k < 2; k++)
	{
		const int c = lane + 32 * k;
		const int row = c / 8, col = (c % 8) * 4;
		const int x = tx0 + col, y = ty0 + row;
		const int e = pixelEntry(col, row);
		uint4 vc = make_uint4(0u, 0u, 0u, 0u);                                              /* colour 0 */
		uint4 vd = make_uint4(0xBF800000u, 0xBF800000u, 0xBF800000u, 0xBF800000u);          /* depth -1 */
		if ((loadColor || loadDepth) && y < st.height && x < st.width)
		{
			const size_t at = (size_t) y * st.width + x;
			if (io.vec4)
			{
				if (loadColor) vc = *(const uint4*) (fr.color + at);
				if (loadDepth) vd = *(const uint4*) (fr.depth + at);
			}
			else
			{
				uint32_t* pc = &vc.x; uint32_t* pd = &vd.x;
				for (int i = 0; i < 4; i++)
					if (x + i < st.width)
					{
						if (loadColor) pc[i] = fr.color[at + i];
						if (loadDepth) pd[i] = __float_as_uint(fr.depth[at + i]);
					}
			}
		}
		*(uint4*) (wt.color + e) = vc;
		*(uint4*) (wt.depth + e) = vd;
	}
	/* stencil: 64 chunks of 4 bytes, two per lane */
	#pragma unroll
	for (int k = 0; k < 2; k++)
	{
		const int c = lane + 32 * k;
		const int row = c / 8, col = (c % 8) * 4;
		const int x = tx0 + col, y = ty0 + row;
		uint32_t v = 0u;
		if (stencilEnabled && y < st.height && x < st.width)
		{
			const uint8_t* src = fr.stencil + (size_t) y * st.width + x;
			if (io.vec4 && (((uintptr_t) fr.stencil) & 3u) == 0)
				v = *(const uint32_t*) src;
			else
				for (int i = 0; i < 4; i++)
					if (x + i < st.width) v |= (uint32_t) src[i] << (8 * i);
		}
		*(uint32_t*) (wt.stencil + stencilEntry(col, row)) = v;
	}
}

/* shared memory -> the planes.  With a pending clear every pixel of colour and depth is written
 * (untouched ones with the clear values); otherwise only the planes that some fragment changed
 * (`dirty`: bit0 colour, bit1 depth, bit2 stencil) -- their other pixels go back as they were
 * loaded.  Full 128-byte rows of colour / depth, 32-byte rows of stencil. */
__device__ __forceinline__ void storeWarpTile(const SrpdState& st, const SrpdFrame& fr, const TileIO& io, uint32_t dirty,
                                              const WarpTile& wt, int tx0, int ty0, int lane)
{
	const bool storeColor = fr.clearPending || (dirty & 1u), storeDepth = fr.clearPending || (dirty & 2u);
	if (storeColor || storeDepth)
	{
		#pragma unroll
		for (int k = 0; k < 2; k++)
		{
			const int c = lane + 32 * k;
			const int row = c / 8, col = (c % 8) * 4;
			const int x = tx0 + col, y = ty0 + row;
			if (y >= st.height || x >= st.width)
				continue;
			const int e = pixelEntry(col, row);
			const size_t at = (size_t) y * st.width + x;
			const uint4 vc = *(const uint4*) (wt.color + e), vd = *(const uint4*) (wt.depth + e);
			if (io.vec4)
			{
				if (storeColor) *(uint4*) (fr.color + at) = vc;
				if (storeDepth) *(uint4*) (fr.depth + at) = vd;
			}
			else
			{
				const uint32_t* pc = &vc.x; const uint32_t* pd = &vd.x;
				for (int i = 0; i < 4; i++)
					if (x + i < st.width)
					{
						if (storeColor) fr.color[at + i] = pc[i];
						if (storeDepth) fr.depth[at + i] = __uint_as_float(pd[i]);
					}
			}
		}
	}
	if (dirty & 4u)
	{
		if (io.vec16)
		{
			/* lanes 0..15: one 16-pixel half row each */
			if (lane < 16)
			{
				const int row = lane / 2, col = (lane % 2) * 16;
				const int x = tx0 + col, y = ty0 + row;
				if (y < st.height && x < st.width)
					*(uint4*) (fr.stencil + (size_t) y * st.width + x) = *(const uint4*) (wt.stencil + stencilEntry(col, row));
			}
		}
		else
		{
			#pragma unroll
			for (int k = 0; k < 2; k++)
			{
				const int c = lane + 32 * k;
				const int row = c / 8, col = (c % 8) * 4;
				const int x = tx0 + col, y = ty0 + row;
				if (y >= st.height || x >= st.width)
					continue;
				const uint32_t v = *(const uint32_t*) (wt.stencil + stencilEntry(col, row));
				uint8_t* dst = fr.stencil + (size_t) y * st.width + x;
				for (int i = 0; i < 4; i++)
					if (x + i < st.width) dst[i] = (uint8_t) (v >> (8 * i));
			}
		}
	}
}

/* The same write-back as TMA tile stores (cp.async.bulk.tensor, SASS: UTMASTG): one lane hands the
 * whole 32x8 tile of a plane to the copy engine -- the tensor map (runtime.cu) describes the
 * row-major plane, a 32x8 box and, for colour / depth, the 128-byte swizzle the shared-memory
 * planes are kept in; parts of the box beyond the framebuffer edge are clipped by the hardware.
 * The generic-proxy writes of the fragments are made visible to the async proxy first; the
 * stores are committed as one bulk group, which the warp waits for (its READ of shared memory)
 * before it re-initialises the planes for its next tile. */
__device__ __forceinline__ uint32_t sharedAddress(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmaStorePlane(const CUtensorMap* map, const void* smem, int x, int y)
{
	asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
	             :: "l"(map), "r"(x), "r"(y), "r"(sharedAddress(smem)) : "memory");
}
__device__ __forceinline__ void storeWarpTileTma(const SrpdTileArgs& a, const SrpdFrame& fr, uint32_t dirty, const WarpTile& wt,
                                                 int tx0, int ty0, int lane)
{
	const bool storeColor = fr.clearPending || (dirty & 1u), storeDepth = fr.clearPending || (dirty & 2u);
	/* every lane publishes its own writes of the planes to the async proxy, then one lane issues */
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	__syncwarp();
	if (lane == 0)
	{
		if (storeColor) tmaStorePlane(&a.tmColor, wt.color, tx0, ty0);
		if (storeDepth) tmaStorePlane(&a.tmDepth, wt.depth, tx0, ty0);
		if (dirty & 4u) tmaStorePlane(&a.tmStencil, wt.stencil, tx0, ty0);
		asm volatile("cp.async.bulk.commit_group;" ::: "memory");
	}
}
__device__ __forceinline__ void tmaWaitSharedRead(int lane)
{
	if (lane == 0)
		asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
	__syncwarp();
}

} // namespace

/* One warp tile: load its state, collect the primitives that touch it, visit them in order, write it back. */
template <int KIND, int SIMPLE, bool BATCH_TMA>
__device__ __forceinline__ void processWarpTile(
	const SrpdTileArgs& a, const SrpdFrame& fr, uint32_t frame, int tileX, int tileY8, typename WarpTileK<KIND>::type& wt, FragCounters& cnt, int lane)
{
	const SrpdState& st = a.d.st;
	const bool stencilEnabled = SIMPLE ? false : (bool) st.stencilEnabled;

	/* candidate list of this tile */
	uint32_t begin = 0, end;
	const uint4* list = nullptr;      /* the supertile's list: copies of ordered-view entries */
	if (a.superOffsets && *a.listOverflow == 0u)
	{
		const uint32_t s = ((uint32_t) tileY8 >> (a.superShift + 1)) * a.superX + ((uint32_t) tileX >> a.superShift);
		begin = a.superOffsets[s];
		end = a.superOffsets[s + 1];
		list = a.listEntries;
	}
	else
		end = a.frameCounts[2 * frame + 1];

	const unsigned char* records = a.records + (size_t) frame * a.recCapacity * a.recStride;
	const uint4* ordered = a.ordered + (size_t) frame * a.recCapacity;
	const int tx0 = tileX * SRPD_WT_W, ty0 = tileY8 * SRPD_WT_H;
	const TileIO io = tileIO(st, fr);
	uint32_t dirty = 0u;

	/* TMA write-back: single-frame draws into planes of this process whose pitch keeps the tensor
	 * maps legal (runtime.cu sets the mask); bit2 (stencil) may be missing on its own */
	const uint32_t tma = BATCH_TMA ? a.tmaPlanes : 0u;
	if (tma)
		tmaWaitSharedRead(lane);      /* the previous tile's stores have read the planes */
	loadWarpTile(st, fr, io, stencilEnabled, wt, tx0, ty0, lane);
	__syncwarp();

	int head = 0, waiting = 0;      /* the ring: entries [head, head + waiting) */
	for (uint32_t c = begin; c < end || waiting > 0; c += 32)
	{
		/* keep the candidates whose box touches the tile, in order */
		const uint32_t i = c + lane;
		bool hit = false;
		uint4 ent = make_uint4(0u, 0u, 0u, 0u);
		if (i < end)
		{
			ent = list ? list[i] : ordered[i];              /* {box, record slot, id prefix}, in primitive order */
			const int x0 = (int) (ent.x & 0xFFFFu), y0b = (int) (ent.x >> 16);
			const int x1 = (int) (ent.y & 0xFFFFu), y1b = (int) (ent.y >> 16);
			hit = x0 < tx0 + SRPD_WT_W && x1 > tx0 && y0b < ty0 + SRPD_WT_H && y1b > ty0;
		}
		const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, hit);
		if (hit)
			wt.ring[(head + waiting + __popc(ballot & ((1u << lane) - 1u))) & (SRPD_RING - 1)] = ent;
		waiting += __popc(ballot);
		/* a step when 32 primitives are waiting, or what is left at the end of the list */
		if (waiting < 32 && c + 32 < end)
			continue;
		__syncwarp();
		const int n = min(waiting, 32);
		if constexpr (KIND == SRPD_KIND_TRIANGLE)
			visitTriangles<SIMPLE>(a, fr, records, wt, head, n, tx0, ty0, dirty, cnt, lane);
		else if constexpr (KIND == SRPD_KIND_LINE)
			visitLines(a, fr, records, wt, head, n, tx0, ty0, dirty, cnt, lane);
		else
			visitPoints(a, fr, records, wt, head, n, tx0, ty0, dirty, cnt, lane);
		head = (head + n) & (SRPD_RING - 1);
		waiting -= n;
	}

	dirty = __reduce_or_sync(0xFFFFFFFFu, dirty);
	__syncwarp();
	if (tma == 7u || (tma == 3u && !(dirty & 4u)))
		storeWarpTileTma(a, fr, dirty, wt, tx0, ty0, lane);
	else
		storeWarpTile(st, fr, io, dirty, wt, tx0, ty0, lane);
	__syncwarp();      /* the next tile re-initialises the state */
}

/* A warp tile no primitive touches while a clear is pending: just write the clear values
 * (colour 0, depth -1; reference core/framebuffer.c:57-62), full 128-byte rows. */
__device__ __forceinline__ void clearWarpTile(const SrpdState& st, const SrpdFrame& fr, int tileX, int tileY8, int lane)
{
	const bool vec4 = (st.width % 4) == 0 && ((((uintptr_t) fr.color) | ((uintptr_t) fr.depth)) & 15u) == 0;
	#pragma unroll
	for (int k = 0; k < 2; k++)
	{
		const int c = lane + 32 * k;
		const int x = tileX * SRPD_WT_W + (c % 8) * 4, y = tileY8 * SRPD_WT_H + c / 8;
		if (y >= st.height || x >= st.width)
			continue;
		const size_t at = (size_t) y * st.width + x;
		if (vec4)
		{
			*(uint4*) (fr.color + at) = make_uint4(0u, 0u, 0u, 0u);
			*(uint4*) (fr.depth + at) = make_uint4(0xBF800000u, 0xBF800000u, 0xBF800000u, 0xBF800000u);
		}
		else
			for (int i = 0; i < 4; i++)
				if (x + i < st.width)
				{
					fr.color[at + i] = 0u;
					fr.depth[at + i] = -1.0f;
				}
	}
}

/* Persistent tile kernel: the grid is sized to the machine (CTAs per SM x SM count) and every
 * WARP pulls work items -- groups of `tilesPerItem` consecutive warp tiles of one frame -- from an
 * atomic counter, so neither empty tiles (skipped through the occupancy bitmap the geometry
 * kernel filled) nor hundreds of frames of a batch cost a launch each, and a warp that draws a
 * heavy tile holds nobody up.
 *
 * Register budget: both launch-bound arguments are given explicitly (under device LTO a
 * missing minimum makes the linker's code generator cap the kernel at 64 registers and spill). */
template <int KIND, bool BATCH, int SIMPLE>
__global__ void __launch_bounds__(SRPD_TILE_THREADS, SRPD_TILE_CTAS_PER_SM)
srpdTileKernel(const __grid_constant__ SrpdTileArgs a)
{
	typedef typename WarpScratchK<KIND>::type Scratch;
	extern __shared__ __align__(1024) unsigned char srpdTileSmem[];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	typename WarpTileK<KIND>::type wt;
	{
		Scratch& sc = reinterpret_cast<Scratch*>(srpdTileSmem + SRPD_TILE_WARPS * 2 * SRPD_PLANE_BYTES)[warp];
		wt.color = reinterpret_cast<uint32_t*>(srpdTileSmem + (2 * warp + 0) * SRPD_PLANE_BYTES);
		wt.depth = reinterpret_cast<float*>(srpdTileSmem + (2 * warp + 1) * SRPD_PLANE_BYTES);
		wt.stencil = sc.stencil;
		wt.ring = sc.ring;
		if constexpr (KIND == SRPD_KIND_TRIANGLE)
		{
			wt.frag = sc.frag; wt.tri = sc.tri; wt.pair = sc.pair;
		}
	}

	srpdGridDependencyEnter();
	if (*a.abortFlag)      /* guard (never expected, runtime.cu): leave the framebuffer untouched */
		return;
	/* warp-tile rows of the row range: two per tile row, the last tile row may have one */
	const uint32_t rows8All = ((uint32_t) a.d.st.height + SRPD_WT_H - 1) / SRPD_WT_H;
	const uint32_t row8Lo = a.d.tileRow0 * 2u, row8Hi = min(a.d.tileRow1 * 2u, rows8All);
	const uint32_t rows = row8Hi > row8Lo ? row8Hi - row8Lo : 0u;
	const uint32_t tilesPerFrame = a.tilesX * rows;
	const uint32_t itemsPerFrame = (tilesPerFrame + a.tilesPerItem - 1) / a.tilesPerItem;
	const uint32_t nItems = itemsPerFrame * a.d.nFrames;

	FragCounters cnt;
	cnt.emitted = 0; cnt.shaded = 0;

	for (;;)
	{
		uint32_t item = 0u;
		if (lane == 0)
			item = atomicAdd(a.workCounter, 1u);
		item = __shfl_sync(0xFFFFFFFFu, item, 0);
		if (item >= nItems)
			break;
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
		int tileY8 = (int) (row8Lo + rowInBand);
		for (uint32_t t = first; t < last; t++)
		{
			/* (the occupancy bitmap has one bit per 32x16 tile) */
			const uint32_t tileIndex = ((uint32_t) tileY8 >> 1) * a.tilesX + (uint32_t) tileX;
			const bool occupied = (occ[tileIndex >> 5] >> (tileIndex & 31u)) & 1u;
			if (occupied)
				processWarpTile<KIND, SIMPLE, !BATCH>(a, fr, frame, tileX, tileY8, wt, cnt, lane);
			else if (fr.clearPending)
				clearWarpTile(a.d.st, fr, tileX, tileY8, lane);
			if (++tileX == (int) a.tilesX)
			{
				tileX = 0;
				tileY8++;
			}
		}
	}

	if (!BATCH && a.tmaPlanes && lane == 0)      /* the last tile's stores have left before the CTA gives up its shared memory */
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");

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
	const int bytes = (int) ((2 * SRPD_PLANE_BYTES + sizeof(typename WarpScratchK<KIND>::type)) * SRPD_TILE_WARPS);
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
	/* persistent grid: resident CTAs per SM x SM count (no more warps than work items) */
	const uint32_t rows8All = ((uint32_t) a.d.st.height + SRPD_WT_H - 1) / SRPD_WT_H;
	const uint32_t row8Hi = a.d.tileRow1 * 2u < rows8All ? a.d.tileRow1 * 2u : rows8All;
	const uint32_t rows8 = row8Hi > a.d.tileRow0 * 2u ? row8Hi - a.d.tileRow0 * 2u : 0u;
	const uint32_t tilesPerFrame = a.tilesX * rows8;
	const uint64_t nItems = (uint64_t) ((tilesPerFrame + a.tilesPerItem - 1) / a.tilesPerItem) * a.d.nFrames;
	uint64_t grid = (uint64_t) a.smCount * SRPD_TILE_CTAS_PER_SM;
	if (grid > (nItems + SRPD_TILE_WARPS - 1) / SRPD_TILE_WARPS) grid = (nItems + SRPD_TILE_WARPS - 1) / SRPD_TILE_WARPS;
	if (grid == 0) return;
	if (a.d.kind == SRPD_KIND_TRIANGLE)
		launchTileKernel<SRPD_KIND_TRIANGLE>(a, (unsigned) grid, stream);
	else if (a.d.kind == SRPD_KIND_LINE)
		launchTileKernel<SRPD_KIND_LINE>(a, (unsigned) grid, stream);
	else
		launchTileKernel<SRPD_KIND_POINT>(a, (unsigned) grid, stream);
}
