/* srp-b200 -- geometry front-end kernels (sm_100a).
 *
 * One WARP = one batch of SRPD_GEOM_PRIMS (30) consecutive input primitives of one frame; the
 * warps of a CTA share nothing and there is no CTA barrier.  Replaces, for that batch, the
 * reference's per-primitive loop in
 *   src/pipeline/primitive_assembly.c:37-135 (assembleTrianglesGeneric),
 *   :137-178 (assembleLines), :221-254 (assemblePoints)
 * and everything it calls: topology.c:24-83, core/buffer.c:102-128 (typed index fetch),
 * vertex_processing.c:38-74 (post-VS cache + vertex shader), clipping.c:68-258,
 * raster/triangle.c:113-160 / line.c:79-86 / point.c:76-79 (setup).
 *
 * Stages of a batch (srpdGeomKernel):
 *   1. topology + index fetch: every lane resolves the 1..3 vertex indices of its primitive;
 *   2. post-VS cache = index de-duplication in shared memory: the indices go into a warp-private
 *      128-slot open-addressing table, the distinct ones are numbered by a warp scan, and the
 *      user vertex shader runs ONCE per distinct index of the batch, outputs kept in shared memory;
 *   3. per primitive: outcodes, trivial accept / reject inline; lines, points and polygon-mode
 *      expansion on an out-of-line path; setup, face / degenerate culling, the exact zero-coverage
 *      test of tiny triangles -- counting primitive ids and records first;
 *   4. warp scan of both counts; the batch takes a contiguous range of record slots from the
 *      frame's bump allocator (arrival order) and publishes its totals;
 *   5. records are written with batch-local primitive ids.
 * A batch that contains a triangle crossing a clip plane, or a line that is clipped or longer
 * than two 16-fragment segments, is not processed by the main pass at all: it is appended to a
 * deferred list and a second launch (`CLIPPER`: one warp per CTA, more registers, scratch slots in
 * shared memory) takes all deferred batches at once (clipChunk, emitPreparedLines).
 * srpdBatchOrderKernel then restores the reference's serial order (primitive_assembly.c:64,
 * 90-91: `primitiveID++` over emitted primitives): it prefix-sums the batch totals in batch order
 * and writes the id-ordered view -- per stored primitive its bounding box, its record slot and
 * the id prefix of its batch -- that binning and tiles consume; the records are not touched again.
 *
 * HBM traffic per input triangle: 3 indices + (amortised) its vertices in; per STORED triangle
 * one record (80 B header + 3 blobs), one 8-byte box and one 16-byte ordered-view entry out. */
#include "kernels.cuh"

namespace {

struct ClipVert
{
	SrpdPos p;
	alignas(8) unsigned char vary[SRPD_MAX_VARYING_BYTES];
};

/* result of the count phase kept for the write phase (unclipped filled triangles only), so
 * that the common case runs its setup -- nine IEEE divisions and the f64 islands -- once */
struct FastTriangle
{
	bool valid;      /* the count phase went through the fast path */
	bool stored;
	SrpdTriSetup s;
};

__device__ __forceinline__ uint32_t fetchIndex(const SrpdDraw& d, uint64_t streamIndex)
{
	if (d.ib == nullptr)
		return (uint32_t) streamIndex;
	switch (d.ibElemSize)
	{
		case 1:  return ((const uint8_t*) d.ib)[streamIndex];
		case 2:  return ((const uint16_t*) d.ib)[streamIndex];
		case 4:  return ((const uint32_t*) d.ib)[streamIndex];
		default: return (uint32_t) ((const uint64_t*) d.ib)[streamIndex];
	}
}

/* stream indices of input primitive k, reference topology.c:24-48,62-83 */
__device__ __forceinline__ void resolveTopology(const SrpdDraw& d, uint32_t k, uint64_t s[3])
{
	const uint64_t base = d.startIndex;
	switch (d.topology)
	{
		case SRPD_TOPO_TRIANGLES:
			s[0] = base + 3ull * k; s[1] = s[0] + 1; s[2] = s[0] + 2; break;
		case SRPD_TOPO_TRIANGLE_STRIP:
		{
			const uint64_t odd = k & 1u;   /* odd triangles swap their first two vertices */
			s[0] = base + k + odd; s[1] = base + k + (1 - odd); s[2] = base + k + 2; break;
		}
		case SRPD_TOPO_TRIANGLE_FAN:
			s[0] = base; s[1] = base + k + 1; s[2] = base + k + 2; break;
		case SRPD_TOPO_LINES:
			s[0] = base + 2ull * k; s[1] = s[0] + 1; s[2] = 0; break;
		case SRPD_TOPO_LINE_STRIP:
			s[0] = base + k; s[1] = base + k + 1; s[2] = 0; break;
		case SRPD_TOPO_LINE_LOOP:
			s[0] = base + k; s[1] = base + ((uint64_t) (k + 1) % d.count); s[2] = 0; break;
		default:   /* points */
			s[0] = base + k; s[1] = 0; s[2] = 0; break;
	}
}

__device__ __forceinline__ uint32_t hashInsert(uint32_t* keys, uint32_t key)
{
	uint32_t h = (key * 2654435761u) >> SRPD_HASH_SHIFT;
	for (;;)
	{
		const uint32_t prev = atomicCAS(&keys[h], SRPD_HASH_EMPTY, key);
		if (prev == SRPD_HASH_EMPTY || prev == key)
			return h;
		h = (h + 1) & (SRPD_HASH_SLOTS - 1);
	}
}

/* Where a thread's outputs go.  WRITE = false: count only. */
struct Emitter
{
	const SrpdGeomArgs* a;
	unsigned char* records;     /* frame base */
	uint2* bboxes;              /* frame base */
	uint32_t* occupancy;        /* frame base */
	uint32_t idBase, storeBase; /* global bases of this input primitive (valid when writing) */
	uint32_t nEmit, nStore;
	uint32_t frame;
	bool overflow;
	/* tile-occupancy bits: normally set record by record (beginRecord); a lane that writes ONE record
	 * in a converged warp may leave its (word, mask) here instead, and the warp then sets the bits of
	 * all its lanes with one reduction per distinct word (srpdGeomKernel, step 5) */
	bool deferOcc;
	uint32_t occWord, occMask;
};
constexpr uint32_t SRPD_OCC_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ void copyBlobWords(const unsigned char* src, unsigned char* dst, int bytes)
{
	/* both sides are 8-byte aligned and `bytes` is a multiple of 8 */
	const uint2* s = (const uint2*) src;
	uint2* d = (uint2*) dst;
	for (int i = 0; i < bytes / 8; i++)
		d[i] = s[i];
}

__device__ void storeBlob(const SrpdState& st, const unsigned char* src, float invW, bool persp, unsigned char* dst)
{
	if (st.allFloat)
	{
		/* all-float layout: one pass over the floats, PERSPECTIVE ones pre-multiplied by 1/w */
		const float* s = (const float*) src;
		float* d = (float*) dst;
		uint32_t modes = st.floatModes;
		for (int e = 0; e < st.nFloats; e++, modes >>= 2)
			d[e] = (persp && (modes & 3u) == SRP_INTERPOLATION_MODE_PERSPECTIVE) ? SRP_FMUL(s[e], invW) : s[e];
		for (int e = st.nFloats; e < st.slotSize / 4; e++)
			d[e] = s[e];
		return;
	}
	copyBlobWords(src, dst, st.slotSize);
	if (!persp)
		return;
	for (int ai = 0; ai < st.nVaryings; ai++)
	{
		const SrpdVarying& at = st.varyings[ai];
		if (at.mode != SRP_INTERPOLATION_MODE_PERSPECTIVE)
			continue;
		if (at.type == SRP_FLOAT)
			for (int e = 0; e < at.nItems; e++)
				srpdStore<float>(dst + at.offset + 4 * e, SRP_FMUL(srpdLoad<float>(src + at.offset + 4 * e), invW));
		else if (at.type == SRP_DOUBLE)
			for (int e = 0; e < at.nItems; e++)
				srpdStore<double>(dst + at.offset + 8 * e, SRP_DMUL(srpdLoad<double>(src + at.offset + 8 * e), (double) invW));
	}
}

template <bool WRITE>
__device__ __forceinline__ unsigned char* beginRecord(Emitter& em, const uint32_t w[SRPD_REC_HEADER_WORDS],
                                                       uint16_t x0, uint16_t y0, uint16_t x1, uint16_t y1)
{
	if (!WRITE)
		return nullptr;
	const uint32_t slot = em.storeBase + em.nStore;
	if (slot >= em.a->recCapacity)
	{
		em.overflow = true;
		return nullptr;
	}
	unsigned char* rec = em.records + (size_t) slot * em.a->recStride;
	uint4* h = (uint4*) rec;
	h[0] = make_uint4(w[0], w[1], w[2], w[3]);
	h[1] = make_uint4(w[4], w[5], w[6], w[7]);
	h[2] = make_uint4(w[8], w[9], w[10], w[11]);
	h[3] = make_uint4(w[12], w[13], w[14], em.idBase + em.nEmit);
	h[4] = make_uint4(w[16], w[17], w[18], w[19]);
	em.bboxes[slot] = make_uint2((uint32_t) x0 | ((uint32_t) y0 << 16), (uint32_t) x1 | ((uint32_t) y1 << 16));
	/* occupancy bitmap: one bit per tile the box touches (row-wise word masks) */
	if (x1 > x0 && y1 > y0)
	{
		const uint32_t tx0 = x0 / SRPD_TILE_W, tx1 = (uint32_t) (x1 - 1) / SRPD_TILE_W;
		const uint32_t ty0 = y0 / SRPD_TILE_H, ty1 = (uint32_t) (y1 - 1) / SRPD_TILE_H;
		const uint32_t first = ty0 * em.a->tilesX + tx0, last = ty0 * em.a->tilesX + (tx1 < em.a->tilesX ? tx1 : em.a->tilesX - 1);
		if (em.deferOcc && ty0 == ty1 && ty0 < em.a->tilesY && (first >> 5) == (last >> 5))
		{
			const uint32_t lo = first & 31u, hi = last & 31u;
			em.occWord = first >> 5;
			em.occMask = (hi == 31u ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
		}
		else
		for (uint32_t ty = ty0; ty <= ty1 && ty < em.a->tilesY; ty++)
		{
			const uint32_t b0 = ty * em.a->tilesX + tx0, b1 = ty * em.a->tilesX + (tx1 < em.a->tilesX ? tx1 : em.a->tilesX - 1);
			for (uint32_t w = b0 >> 5; w <= (b1 >> 5); w++)
			{
				const uint32_t lo = w == (b0 >> 5) ? (b0 & 31u) : 0u, hi = w == (b1 >> 5) ? (b1 & 31u) : 31u;
				const uint32_t mask = (hi == 31u ? 0xFFFFFFFFu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
				if ((em.occupancy[w] & mask) != mask)      /* (measured: the pre-check beats unconditional reductions) */
					atomicOr(&em.occupancy[w], mask);
			}
		}
	}
	return rec + SRPD_REC_HEADER_BYTES;
}

/* setup + the exact zero-coverage test; returns false if the triangle is culled (no id) */
__device__ __forceinline__ bool setupAndClassify(const SrpdState& st, const SrpdPos p[3], SrpdTriSetup& s, bool& stored)
{
	if (!srpdSetupTriangle(st, p, s, stored))
		return false;
	/* A triangle whose bounding box is a few pixels (sub-pixel geometry, cfg4) usually covers
	 * no pixel centre at all.  Walking the reference's own chain over the box here decides
	 * that exactly; such a triangle keeps its primitive id but needs no record. */
	if (stored && srpdTriangleIsSmall(s) && !srpdSmallTriangleCoversAnyPixel(s))
		stored = false;
	/* sort-first strips: a primitive whose box misses this process's rows keeps its id (every
	 * rank runs the same scan) but needs no record here */
	if (stored && ((int) s.maxY <= st.stripY0 || (int) s.minY >= st.stripY1))
		stored = false;
	return true;
}

template <bool WRITE>
__device__ __forceinline__ void writeTriangle(Emitter& em, const SrpdState& st, SrpdTriSetup& s, bool stored, const unsigned char* const vary[3])
{
	if (stored)
	{
		if (WRITE)
		{
			/* Large triangle: reserve a checkpoint table (rows x tile columns) and queue it for
			 * the checkpoint pre-pass, so that no tile has to replay more than a tile's width of
			 * its barycentric chain.  If the table or the queue is full the triangle simply has
			 * no checkpoints and the tile kernel replays from the corner (slow, still exact). */
			const uint32_t rows = (uint32_t) (s.maxY - s.minY), width = (uint32_t) (s.maxX - s.minX);
			if (rows > SRPD_LARGE_EXTENT || width > SRPD_LARGE_EXTENT)
			{
				const uint32_t cols = (uint32_t) (s.maxX - 1) / SRPD_TILE_W - (uint32_t) s.minX / SRPD_TILE_W + 1;
				const uint32_t entries = rows * cols;
				const uint32_t slot = em.storeBase + em.nStore;
				/* the cursor is 64 bits wide: it cannot wrap, whatever the number of large triangles of a
				 * draw or batch keeps adding to it after the table is full */
				const uint64_t cap = em.a->ckptCapacity;
				const uint64_t off = atomicAdd(em.a->ckptCursor, (unsigned long long) entries);
				if (off + entries <= cap && slot < em.a->recCapacity)
				{
					const uint32_t q = atomicAdd(em.a->largeCount, 1u);
					if (q < em.a->largeCapacity)
					{
						em.a->largeList[q] = make_uint2(em.frame, slot);
						s.w[19] = (uint32_t) off + 1u;
					}
				}
			}
		}
		unsigned char* blobs = beginRecord<WRITE>(em, s.w, s.minX, s.minY, s.maxX, s.maxY);
		if (WRITE && blobs)
		{
			/* the winding normalisation either keeps the order or swaps v1 and v2: selects, not a
			 * run-time index (which would force `vary` -- and the caller's copy -- into local memory) */
			const bool swapped = s.order[1] != 1;
			storeBlob(st, vary[0], s.invW[0], true, blobs);
			storeBlob(st, swapped ? vary[2] : vary[1], s.invW[1], true, blobs + st.slotSize);
			storeBlob(st, swapped ? vary[1] : vary[2], s.invW[2], true, blobs + 2 * st.slotSize);
		}
		em.nStore++;
	}
	em.nEmit++;
}

template <bool WRITE>
__device__ __noinline__ void emitTriangle(Emitter& em, const SrpdState& st, const SrpdPos p[3], const unsigned char* const vary[3])
{
	SrpdTriSetup s;
	bool stored;
	if (setupAndClassify(st, p, s, stored))
		writeTriangle<WRITE>(em, st, s, stored, vary);
}

template <bool WRITE>
__device__ __noinline__ void emitLine(Emitter& em, const SrpdState& st, const SrpdPos p[2], const unsigned char* const vary[2])
{
	SrpdLineSetup ln;
	srpdSetupLine(st, p, ln);
	float x = ln.x0, y = ln.y0, t = 0.f;
	const int total = ln.steps + 1;
	for (int i = 0; i < total; i += SRPD_LINE_SEG)
	{
		SrpdLineSegment seg;
		const int n = total - i < SRPD_LINE_SEG ? total - i : SRPD_LINE_SEG;
		srpdLineSegment(st, ln, x, y, t, n, seg);
		if (!seg.any || (int) seg.maxY <= st.stripY0 || (int) seg.minY >= st.stripY1)      /* (outside this process's strip) */
			continue;
		unsigned char* blobs = beginRecord<WRITE>(em, seg.w, seg.minX, seg.minY, seg.maxX, seg.maxY);
		if (WRITE && blobs)
			for (int k = 0; k < 2; k++)
				storeBlob(st, vary[k], ln.invW[k], true, blobs + k * st.slotSize);
		em.nStore++;
	}
	em.nEmit++;
}

template <bool WRITE>
__device__ __noinline__ void emitPoint(Emitter& em, const SrpdState& st, const SrpdPos& p, const unsigned char* vary)
{
	SrpdPointSetup s;
	if (srpdSetupPoint(st, p, s) && (int) s.maxY > st.stripY0 && (int) s.minY < st.stripY1)
	{
		unsigned char* blob = beginRecord<WRITE>(em, s.w, s.minX, s.minY, s.maxX, s.maxY);
		if (WRITE && blob)
			storeBlob(st, vary, 1.0f, false, blob);
		em.nStore++;
	}
	em.nEmit++;
}

/* one (possibly clip-generated) triangle through the polygon mode,
 * reference primitive_assembly.c:80-129 */
template <bool WRITE>
__device__ void emitPolygonModeTriangle(Emitter& em, const SrpdState& st, const SrpdPos p[3], const unsigned char* const vary[3])
{
	if (st.polygonMode == SRP_POLYGON_MODE_FILL)
		emitTriangle<WRITE>(em, st, p, vary);
	else if (st.polygonMode == SRP_POLYGON_MODE_LINE)
		for (int j = 0; j < 3; j++)
		{
			const SrpdPos lp[2] = { p[j], p[(j + 1) % 3] };
			const unsigned char* const lv[2] = { vary[j], vary[(j + 1) % 3] };
			emitLine<WRITE>(em, st, lp, lv);
		}
	else
		for (int j = 0; j < 3; j++)
			emitPoint<WRITE>(em, st, p[j], vary[j]);
}

/* clipTriangle + expansion, reference clipping.c:68-119 and :199-258.
 *
 * Sutherland-Hodgman over the six planes in the reference's order (L, R, B, T, N, F) with its
 * emission rule -- for every edge (current, next): both inside -> emit next; crossing -> emit
 * the intersection, and next too when entering -- so the polygon's first vertex rotates from
 * plane to plane exactly as in the reference (that decides the fan and the provoking vertex).
 *
 * Few triangles cross a clip plane (cfg3: 2 k of a million), but the warps that own them decide
 * when the kernel ends, so the clipped path is built for LATENCY:
 *   - batches with such triangles are not processed where they are found: the main pass only
 *     notes them down, and a second launch of this kernel -- one warp per CTA, more registers,
 *     a private scratch area -- takes them all at once, so they run side by side from time zero
 *     instead of trailing behind the last ordinary batches (srpdLaunchGeom);
 *   - every clipped triangle of the batch gets a scratch slot in shared memory (vertex pool +
 *     polygon index lists; the 3 originals are copied from the post-VS cache, intersections
 *     are appended);
 *   - the owner lanes clip side by side, one lane per triangle, entirely in that scratch (no
 *     indexed local memory), keeping the edge's first endpoint and distance in registers;
 *   - the fan triangles (0, i, i+1) of ALL clipped triangles of the batch are then spread over
 *     the lanes -- 32 setups per pass through the polygon mode, whatever mix of triangles they
 *     came from -- and their id / record counts are left in the slots;
 *   - the slots stay allocated across the batch's scan, so the write phase only hands every
 *     fan triangle its id / record base and emits it: nothing is clipped twice. */
constexpr int SRPD_CLIP_NEW_VERTS = 12;    /* at most two intersections per plane */
constexpr int SRPD_CLIP_POOL = 3 + SRPD_CLIP_NEW_VERTS;
constexpr int SRPD_CLIP_SLOTS = 32;        /* scratch slots of a clipper warp: one per lane, every triangle of a batch may need one */
constexpr int SRPD_CLIP_MAX_FANS = SRPD_CLIP_MAX_VERTS - 2;

struct ClipSlot
{
	float4 pos[SRPD_CLIP_POOL];                 /* clip-space positions of the pool vertices           */
	uint32_t idBase, storeBase;                 /* the owner's bases (write phase)                     */
	uint16_t cnt[SRPD_CLIP_MAX_FANS][2];        /* per fan triangle: primitive ids, records            */
	uint8_t list[2][12];                        /* polygon (pool indices), ping-pong                   */
	uint8_t origSlot[3];                        /* post-VS cache slots of the 3 original vertices      */
	uint8_t n;                                  /* vertices of the clipped polygon                     */
	uint8_t cur, pad[3];                        /* which list holds it                                 */
};                                              /* followed by SRPD_CLIP_NEW_VERTS varyings blobs       */

struct LineSlot;
__host__ __device__ inline size_t clipSlotStride(int slotSize)
{
	/* (a line's slot -- LineSlot + two blobs -- fits as well: static_assert below) */
	size_t b = (sizeof(ClipSlot) + (size_t) SRPD_CLIP_NEW_VERTS * slotSize + 15) & ~(size_t) 15;
	if (b % 128 == 0) b += 16;                  /* keep the slots of neighbouring lanes on different banks */
	return b;
}

__device__ __forceinline__ const unsigned char* clipVary(const ClipSlot& sl, const unsigned char* fresh, const unsigned char* vvary, int slotSize, int v)
{
	return v < 3 ? vvary + (size_t) sl.origSlot[v] * slotSize : fresh + (size_t) (v - 3) * slotSize;
}

/* a fan triangle through the polygon mode; the filled case inline (the clipper warp has the
 * registers for it, and its latency is what the clipper pass consists of) */
template <bool WRITE>
__device__ __forceinline__ void emitFanTriangle(Emitter& em, const SrpdState& st, const SrpdPos p[3], const unsigned char* const vary[3])
{
	if (st.polygonMode == SRP_POLYGON_MODE_FILL)
	{
		SrpdTriSetup s;
		bool stored;
		if (setupAndClassify(st, p, s, stored))
			writeTriangle<WRITE>(em, st, s, stored, vary);
	}
	else
		emitPolygonModeTriangle<WRITE>(em, st, p, vary);
}

/* set-up of the first SRPD_FAN_CACHE fan triangles of a chunk, kept from the COUNT call for the
 * WRITE call (filled polygons): the clipper pass is the latency of ONE warp per deferred batch, and
 * a second set-up -- nine IEEE divisions and the f64 islands -- is a sixth of it */
constexpr int SRPD_FAN_CACHE = 32;
struct FanCache
{
	SrpdTriSetup s;
	uint32_t stored;      /* 0 / 1; 2 = culled (no id) */
};

struct ClipResult { uint32_t nEmit, nStore; bool overflow; };
enum { SRPD_CLIP_COUNT = 1, SRPD_CLIP_WRITE = 2 };

/* The clipped triangles of a batch, owned by the lanes of chunkMask (the r-th owner uses slot r).
 * mode: COUNT = clip + count the fan triangles (returns, in the owner lanes, the ids / records
 * their triangle produces); WRITE = emit them from idBase / storeBase of their owner, from the
 * slots the COUNT call left behind.  Called by all 32 lanes. */
__device__ __forceinline__ ClipResult clipChunk(
	Emitter em, const SrpdState& st, int mode, unsigned char* slots, FanCache* fanCache, uint32_t chunkMask,
	const float4* vpos, const unsigned char* vvary, uint32_t slot0, uint32_t slot1, uint32_t slot2,
	uint32_t idBase, uint32_t storeBase)
{
	ClipResult res;
	res.nEmit = 0u; res.nStore = 0u; res.overflow = false;
	const int lane = threadIdx.x & 31;
	const bool mine = (chunkMask >> lane) & 1u;
	const size_t stride = clipSlotStride(st.slotSize);
	const uint32_t mySlot = mine ? (uint32_t) __popc(chunkMask & ((1u << lane) - 1u)) : 0u;
	ClipSlot& sl = *reinterpret_cast<ClipSlot*>(slots + (size_t) mySlot * stride);
	unsigned char* fresh = reinterpret_cast<unsigned char*>(&sl + 1);

	if ((mode & SRPD_CLIP_COUNT) && mine)
	{
		sl.pos[0] = vpos[slot0]; sl.pos[1] = vpos[slot1]; sl.pos[2] = vpos[slot2];
		sl.origSlot[0] = (uint8_t) slot0; sl.origSlot[1] = (uint8_t) slot1; sl.origSlot[2] = (uint8_t) slot2;
		sl.list[0][0] = 0; sl.list[0][1] = 1; sl.list[0][2] = 2;
		int n = 3, nPool = 3, cur = 0;
		#pragma unroll      /* the plane becomes a constant: its distance is one add, not a jump table */
		for (int plane = 0; plane < 6; plane++)
		{
			if (n == 0)
				break;
			const uint8_t* src = sl.list[cur];
			uint8_t* dst = sl.list[cur ^ 1];
			int o = 0;
			int vi = src[0];
			float4 qa = sl.pos[vi];
			SrpdPos pa; pa.x = qa.x; pa.y = qa.y; pa.z = qa.z; pa.w = qa.w;
			float da = srpdPlaneDistance(pa, plane);
			for (int i = 0; i < n; i++)
			{
				const int vk = src[i + 1 == n ? 0 : i + 1];
				const float4 qb = sl.pos[vk];
				SrpdPos pb; pb.x = qb.x; pb.y = qb.y; pb.z = qb.z; pb.w = qb.w;
				const float db = srpdPlaneDistance(pb, plane);
				const bool ci = da >= 0, ni = db >= 0;
				if (ci && ni)
				{
					if (o < SRPD_CLIP_MAX_VERTS) dst[o++] = (uint8_t) vk;
				}
				else if (ci || ni)
				{
					const float diff = SRP_FSUB(da, db);
					if (!srpdRoughlyZero(diff))              /* clipping.c:226: otherwise the edge is skipped */
					{
						const float t = SRP_FDIV(da, diff);
						if (o < SRPD_CLIP_MAX_VERTS && nPool < SRPD_CLIP_POOL)
						{
							const SrpdPos np = srpdBlendPos(pa, pb, t);
							sl.pos[nPool] = make_float4(np.x, np.y, np.z, np.w);
							srpdBlendVaryings(st, clipVary(sl, fresh, vvary, st.slotSize, vi), clipVary(sl, fresh, vvary, st.slotSize, vk),
							                  SRP_FSUB(1.0f, t), t, fresh + (size_t) (nPool - 3) * st.slotSize);
							dst[o++] = (uint8_t) nPool++;
						}
						if (!ci && ni)
							if (o < SRPD_CLIP_MAX_VERTS) dst[o++] = (uint8_t) vk;
					}
				}
				pa = pb; da = db; vi = vk;
			}
			n = o;
			cur ^= 1;
		}
		sl.n = (uint8_t) n;
		sl.cur = (uint8_t) cur;
	}
	if ((mode & SRPD_CLIP_WRITE) && mine)
	{
		sl.idBase = idBase; sl.storeBase = storeBase;
	}
	__syncwarp();

	/* fan (0, i, i+1), clipping.c:103-118, of all the chunk's triangles, spread over the lanes */
	const uint32_t fans = mine ? (uint32_t) (sl.n > 2 ? sl.n - 2 : 0) : 0u;
	uint32_t inc = fans;
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
		if (lane >= o) inc += v;
	}
	const uint32_t nFans = __shfl_sync(0xFFFFFFFFu, inc, 31);
	const uint32_t fanBase = inc - fans;
	for (int pass = 0; pass < 2; pass++)
	{
		const bool writing = pass == 1;
		if (!(mode & (writing ? SRPD_CLIP_WRITE : SRPD_CLIP_COUNT)))
			continue;
		for (uint32_t f0 = 0; f0 < nFans; f0 += 32)
		{
			const uint32_t f = f0 + lane;
			const bool valid = f < nFans;
			/* whose fan triangle is f?  the last owner whose base is <= f */
			uint32_t oSlot = 0u, oBase = 0u;
			for (uint32_t m = chunkMask; m != 0u; m &= m - 1u)
			{
				const int L = __ffs(m) - 1;
				const uint32_t b = __shfl_sync(0xFFFFFFFFu, fanBase, L), sidx = __shfl_sync(0xFFFFFFFFu, mySlot, L);
				const uint32_t cnt = __shfl_sync(0xFFFFFFFFu, fans, L);
				if (cnt != 0u && f >= b) { oSlot = sidx; oBase = b; }
			}
			if (!valid)
				continue;
			ClipSlot& os = *reinterpret_cast<ClipSlot*>(slots + (size_t) oSlot * stride);
			const unsigned char* ofresh = reinterpret_cast<const unsigned char*>(&os + 1);
			const int i = (int) (f - oBase);
			SrpdPos tp[3];
			const unsigned char* tv[3];
			#pragma unroll
			for (int k = 0; k < 3; k++)
			{
				const int v = os.list[os.cur][k == 0 ? 0 : i + k];
				const float4 q = os.pos[v];
				tp[k].x = q.x; tp[k].y = q.y; tp[k].z = q.z; tp[k].w = q.w;
				tv[k] = clipVary(os, ofresh, vvary, st.slotSize, v);
			}
			em.idBase = 0; em.storeBase = 0; em.nEmit = 0; em.nStore = 0; em.overflow = false;
			const bool cached = st.polygonMode == SRP_POLYGON_MODE_FILL && f < (uint32_t) SRPD_FAN_CACHE;
			if (!writing)
			{
				if (cached)
				{
					FanCache& fc = fanCache[f];
					bool stored = false;
					const bool alive = setupAndClassify(st, tp, fc.s, stored);
					fc.stored = alive ? (stored ? 1u : 0u) : 2u;
					em.nEmit = alive ? 1u : 0u;
					em.nStore = stored ? 1u : 0u;
				}
				else
					emitFanTriangle<false>(em, st, tp, tv);
				os.cnt[i][0] = (uint16_t) em.nEmit; os.cnt[i][1] = (uint16_t) em.nStore;
			}
			else if (os.cnt[i][1] != 0)
			{
				uint32_t preE = 0u, preS = 0u;
				for (int j = 0; j < i; j++) { preE += os.cnt[j][0]; preS += os.cnt[j][1]; }
				em.idBase = os.idBase + preE;
				em.storeBase = os.storeBase + preS;
				if (cached)
				{
					SrpdTriSetup s = fanCache[f].s;      /* (stored == 1: it has a record) */
					writeTriangle<true>(em, st, s, true, tv);
				}
				else
					emitFanTriangle<true>(em, st, tp, tv);
				if (em.overflow) res.overflow = true;
			}
		}
		__syncwarp();
	}
	if (mine)
		for (uint32_t j = 0; j < fans; j++) { res.nEmit += sl.cnt[j][0]; res.nStore += sl.cnt[j][1]; }
	return res;
}

/* clipLine (Liang-Barsky), reference clipping.c:139-183: false if nothing of the line is left;
 * otherwise cp / cv are the end points to set up (originals, or blends kept in `a` / `b`) */
__device__ __forceinline__ bool clipLineEnds(const SrpdState& st, const SrpdPos p[2], const unsigned char* const vary[2],
                                             ClipVert& a, ClipVert& b, SrpdPos cp[2], const unsigned char* cv[2])
{
	cp[0] = p[0]; cp[1] = p[1];
	cv[0] = vary[0]; cv[1] = vary[1];
	const uint32_t c0 = srpdClipCode(p[0]), c1 = srpdClipCode(p[1]);
	if ((c0 | c1) == 0)
		return true;
	if ((c0 & c1) != 0)
		return false;

	float t0 = 0.f, t1 = 1.f;
	for (int plane = 0; plane < 6; plane++)
	{
		const float da = srpdPlaneDistance(p[0], plane);
		const float db = srpdPlaneDistance(p[1], plane);
		if (da < 0 && db < 0)
			return false;
		if (da < 0 || db < 0)
		{
			const float diff = SRP_FSUB(da, db);
			if (srpdRoughlyZero(diff))
				continue;
			const float t = SRP_FDIV(da, diff);
			if (da < 0)
				t0 = (t0 > t) ? t0 : t;     /* MAX(t0, t) */
			else
				t1 = (t1 > t) ? t : t1;     /* MIN(t1, t) */
			if (t0 > t1)
				return false;
		}
	}
	if (t0 > 0)
	{
		a.p = srpdBlendPos(p[0], p[1], t0);
		srpdBlendVaryings(st, vary[0], vary[1], SRP_FSUB(1.0f, t0), t0, a.vary);
		cp[0] = a.p; cv[0] = a.vary;
	}
	if (t1 < 1)
	{
		b.p = srpdBlendPos(p[0], p[1], t1);
		srpdBlendVaryings(st, vary[0], vary[1], SRP_FSUB(1.0f, t1), t1, b.vary);
		cp[1] = b.p; cv[1] = b.vary;
	}
	return true;
}

template <bool WRITE>
__device__ void processLine(Emitter& em, const SrpdState& st, const SrpdPos p[2], const unsigned char* const vary[2])
{
	ClipVert a, b;
	SrpdPos cp[2];
	const unsigned char* cv[2];
	if (!clipLineEnds(st, p, vary, a, b, cp, cv))
		return;
	const unsigned char* const cvc[2] = { cv[0], cv[1] };
	emitLine<WRITE>(em, st, cp, cvc);
}

/* ---- lines in the clipper pass ------------------------------------------------------------
 * A line is one lane's work in the main pass: its DDA chain is serial (float additions,
 * line.c:72-74) and so is the lane's walk over its 16-fragment segments.  That is fine for the
 * short lines a mesh is made of and hopeless for a line that crosses the screen (240 segments
 * while the other lanes of the warp wait): batches with a long line go to the clipper pass,
 * where the line's segments are spread over the 32 lanes.  Every lane first clips and sets up
 * its own line into a slot in shared memory (set-up + both varyings blobs as the records will
 * hold them); short lines are then emitted by their lanes side by side, long lines one after
 * the other by the whole warp: all lanes walk the chain through 32 segments together (the same
 * additions in the same order -- what is serial stays serial, it is 3 FADDs per fragment), lane
 * j keeps the state at the start of segment j and then boxes and writes that segment on its
 * own.  Record slots and ids are the owner lane's, segment order is kept by a ballot. */
struct LineSlot
{
	SrpdLineSetup ln;
	uint32_t valid, pad;
};                                              /* followed by the two varyings blobs of the records */
constexpr int SRPD_LINE_SHORT_SEGS = 2;         /* lines of up to this many segments stay with their lane */
static_assert(sizeof(LineSlot) <= sizeof(ClipSlot) && sizeof(LineSlot) % 8 == 0, "a line's slot lives in a clip slot");

__device__ __forceinline__ int lineSegments(const SrpdLineSetup& ln) { return (ln.steps + 1 + SRPD_LINE_SEG - 1) / SRPD_LINE_SEG; }

__device__ void prepareLine(const SrpdState& st, const SrpdPos p[2], const unsigned char* const vary[2], LineSlot& sl)
{
	ClipVert a, b;
	SrpdPos cp[2];
	const unsigned char* cv[2];
	if (!clipLineEnds(st, p, vary, a, b, cp, cv))
		return;
	srpdSetupLine(st, cp, sl.ln);
	unsigned char* blobs = reinterpret_cast<unsigned char*>(&sl + 1);
	for (int k = 0; k < 2; k++)
		storeBlob(st, cv[k], sl.ln.invW[k], true, blobs + k * st.slotSize);
	sl.valid = 1u;
}

template <bool WRITE>
__device__ __forceinline__ bool emitLineSegment(const Emitter& em, const SrpdState& st, const SrpdLineSegment& seg, const unsigned char* blobs,
                                                uint32_t id, uint32_t slot)
{
	if (!WRITE)
		return false;
	Emitter one = em;
	one.idBase = id; one.nEmit = 0u;
	one.storeBase = slot; one.nStore = 0u;
	one.overflow = false;
	unsigned char* dst = beginRecord<true>(one, seg.w, seg.minX, seg.minY, seg.maxX, seg.maxY);
	if (dst)
		copyBlobWords(blobs, dst, 2 * st.slotSize);
	return one.overflow;
}

/* count (WRITE = false: em.nEmit / em.nStore of the owner lanes) or write (em.idBase / em.storeBase
 * are the owners' bases) the prepared lines of a batch; all 32 lanes */
template <bool WRITE>
__device__ __noinline__ void emitPreparedLines(Emitter& em, const SrpdState& st, const unsigned char* slots, size_t stride, int lane)
{
	const LineSlot& mine = *reinterpret_cast<const LineSlot*>(slots + (size_t) lane * stride);
	const bool valid = mine.valid != 0u;
	const bool isShort = valid && lineSegments(mine.ln) <= SRPD_LINE_SHORT_SEGS;
	if (isShort)
	{
		const SrpdLineSetup ln = mine.ln;
		const unsigned char* blobs = reinterpret_cast<const unsigned char*>(&mine + 1);
		float x = ln.x0, y = ln.y0, t = 0.f;
		const int total = ln.steps + 1;
		uint32_t stored = 0u;
		for (int i = 0; i < total; i += SRPD_LINE_SEG)
		{
			SrpdLineSegment seg;
			const int n = total - i < SRPD_LINE_SEG ? total - i : SRPD_LINE_SEG;
			srpdLineSegment(st, ln, x, y, t, n, seg);
			if (!seg.any || (int) seg.maxY <= st.stripY0 || (int) seg.minY >= st.stripY1)
				continue;
			if (emitLineSegment<WRITE>(em, st, seg, blobs, em.idBase, em.storeBase + stored))
				em.overflow = true;
			stored++;
		}
		em.nEmit = 1u; em.nStore = stored;
	}
	__syncwarp();
	for (uint32_t m = __ballot_sync(0xFFFFFFFFu, valid && !isShort); m != 0u; m &= m - 1u)
	{
		const int owner = __ffs(m) - 1;
		const LineSlot& sl = *reinterpret_cast<const LineSlot*>(slots + (size_t) owner * stride);
		const unsigned char* blobs = reinterpret_cast<const unsigned char*>(&sl + 1);
		const SrpdLineSetup ln = sl.ln;
		const uint32_t ownerId = __shfl_sync(0xFFFFFFFFu, em.idBase, owner), ownerStore = __shfl_sync(0xFFFFFFFFu, em.storeBase, owner);
		const int total = ln.steps + 1;
		const int nSeg = lineSegments(ln);
		float x = ln.x0, y = ln.y0, t = 0.f;      /* the chain at the first fragment of the round's first segment: the same in every lane */
		uint32_t stored = 0u;
		bool overflow = false;
		for (int s0 = 0; s0 < nSeg; s0 += 32)
		{
			const int here = nSeg - s0 < 32 ? nSeg - s0 : 32;
			const int walk = (s0 + 32 < nSeg ? here : here - 1) * SRPD_LINE_SEG;      /* to the next round, or to the last segment's start */
			float mx = x, my = y, mt = t;
			for (int k = 0; ; k++)
			{
				if (k == lane * SRPD_LINE_SEG) { mx = x; my = y; mt = t; }
				if (k == walk)
					break;
				x = SRP_FADD(x, ln.xInc);
				y = SRP_FADD(y, ln.yInc);
				t = SRP_FADD(t, ln.tInc);
			}
			bool keep = false;
			SrpdLineSegment seg;
			if (lane < here)
			{
				const int i = (s0 + lane) * SRPD_LINE_SEG;
				const int n = total - i < SRPD_LINE_SEG ? total - i : SRPD_LINE_SEG;
				srpdLineSegment(st, ln, mx, my, mt, n, seg);
				keep = seg.any && !((int) seg.maxY <= st.stripY0 || (int) seg.minY >= st.stripY1);
			}
			const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
			if (keep)
				overflow |= emitLineSegment<WRITE>(em, st, seg, blobs, ownerId, ownerStore + stored + (uint32_t) __popc(ballot & ((1u << lane) - 1u)));
			stored += (uint32_t) __popc(ballot);
		}
		if (__any_sync(0xFFFFFFFFu, overflow))
			em.overflow = true;
		if (lane == owner)
		{
			em.nEmit = 1u; em.nStore = stored;
		}
	}
}

/* Lines, points and unclipped triangles in polygon modes LINE / POINT go through here.  Deliberately NOT inlined
 * (nor are the emit* helpers): inlined into the kernel at every call site the front-end grew to
 * 38 k instructions and, with the warps of an SM in different stages, stalled on instruction
 * fetch a third of the time; as functions the common path is a few thousand instructions. */
template <bool WRITE>
__device__ __noinline__ void processPrimitive(Emitter& em, const SrpdDraw& d, int nv,
                                                 const SrpdPos p[3], const unsigned char* const vary[3])
{
	if (nv == 3)      /* unclipped, polygon mode LINE / POINT (clipped triangles: clipChunk) */
		emitPolygonModeTriangle<WRITE>(em, d.st, p, vary);
	else if (nv == 2)
		processLine<WRITE>(em, d.st, p, vary);
	else if (!srpdClipPoint(p[0]))
		emitPoint<WRITE>(em, d.st, p[0], vary[0]);
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

} // namespace

/* per-warp shared memory: the post-VS cache of one batch */
struct GeomWarpShared
{
	uint32_t hashKey[SRPD_HASH_SLOTS];       /* open-addressing table of the batch's vertex indices   */
	uint8_t  hashDense[SRPD_HASH_SLOTS];     /* table slot -> dense number of the distinct index      */
	uint32_t uniq[SRPD_GEOM_MAX_VERTS];      /* dense number -> vertex index                          */
	float4   vpos[SRPD_GEOM_MAX_VERTS];      /* clip-space positions (VS output)                      */
};                                           /* followed by SRPD_GEOM_MAX_VERTS varyings blobs         */

/* per warp: the cache and its varyings blobs; a clipper warp adds the clip slots behind them */
__host__ __device__ inline size_t geomWarpBytes(int slotSize)
{
	return (sizeof(GeomWarpShared) + (size_t) SRPD_GEOM_MAX_VERTS * slotSize + 15) & ~(size_t) 15;
}
__host__ __device__ inline size_t geomClipSlotBytes(int slotSize)
{
	return ((size_t) SRPD_CLIP_SLOTS * clipSlotStride(slotSize) + 15) & ~(size_t) 15;
}
__host__ __device__ inline size_t geomCtaBytes(int slotSize, bool clipper)
{
	return clipper ? geomWarpBytes(slotSize) + geomClipSlotBytes(slotSize) + (size_t) SRPD_FAN_CACHE * sizeof(FanCache)
	               : (size_t) SRPD_GEOM_WARPS * geomWarpBytes(slotSize);
}

/* One WARP = one batch of SRPD_GEOM_PRIMS consecutive input primitives; the warps of a CTA are
 * independent of each other (no CTA barrier anywhere), so a warp that sits in a long-latency
 * step -- vertex fetch, the rare clipping path, the bump-allocator atomic -- never holds up the
 * other seven: the SM always has warps in different stages to issue from. */
/* CLIPPER = false: the main pass of a large draw, SRPD_GEOM_WARPS batches per CTA; a batch that
 * contains a triangle crossing a clip plane is appended to the deferred list and left alone.
 * CLIPPER = true: one warp per CTA with the clipper's scratch and a larger register budget;
 * it processes either the deferred list of the main pass or, for small draws (where a second
 * launch would cost more than it saves), every batch. */
template <bool BATCH, bool CLIPPER>
__global__ void __launch_bounds__(CLIPPER ? 32 : SRPD_GEOM_THREADS, CLIPPER ? 16 : SRPD_GEOM_CTAS_PER_SM)
srpdGeomKernel(const __grid_constant__ SrpdGeomArgs a)
{
	extern __shared__ __align__(16) unsigned char smem[];
	srpdGridDependencyEnter();
	const SrpdDraw& d = a.d;
	const SrpdState& st = d.st;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	constexpr int WARPS = CLIPPER ? 1 : SRPD_GEOM_WARPS;

	const size_t warpBytes = geomWarpBytes(st.slotSize);
	GeomWarpShared& ws = *reinterpret_cast<GeomWarpShared*>(smem + warp * warpBytes);
	unsigned char* vvary = reinterpret_cast<unsigned char*>(&ws + 1);
	unsigned char* clipSlots = smem + warpBytes;      /* CLIPPER only */
	FanCache* fanCache = reinterpret_cast<FanCache*>(clipSlots + geomClipSlotBytes(st.slotSize));
	const uint32_t nBatches = a.batchesPerFrame * d.nFrames;
	const int nv = (d.topology >= SRPD_TOPO_TRIANGLES) ? 3 : (d.topology == SRPD_TOPO_POINTS ? 1 : 2);

	/* one batch per warp; a clipper working off the deferred list strides through it */
	const bool fromList = CLIPPER && a.deferred;
	const uint64_t limit = fromList ? (uint64_t) min(*a.deferCount, nBatches) : (uint64_t) nBatches;
	const uint64_t stride = fromList ? (uint64_t) gridDim.x : (1ull << 40);
	for (uint64_t it = (uint64_t) blockIdx.x * WARPS + warp; it < limit; it += stride)
	{
	const uint32_t batch = fromList ? a.deferList[it] : (uint32_t) it;
	__syncwarp();      /* the previous batch no longer reads the cache */
	#pragma unroll
	for (int i = 0; i < SRPD_HASH_SLOTS / 32; i++)
		ws.hashKey[lane + 32 * i] = SRPD_HASH_EMPTY;
	const uint32_t frame = batch / a.batchesPerFrame;
	const uint32_t b = batch - frame * a.batchesPerFrame;
	/* single draw: the uniform block sits in the argument block (constant bank) */
	const void* uniform = BATCH ? a.frames[frame].uniform : (const void*) a.uniformInline;

	/* (a sub-draw covers input primitives [firstPrim, firstPrim + nInputPrims) of the draw) */
	const uint32_t kLocal = b * SRPD_GEOM_PRIMS + lane;
	const uint32_t k = d.firstPrim + kLocal;
	const bool active = lane < SRPD_GEOM_PRIMS && kLocal < d.nInputPrims;

	/* 1. topology + typed index fetch */
	uint32_t vi[3] = { 0, 0, 0 };
	uint32_t slot[3] = { 0, 0, 0 };
	if (active)
	{
		uint64_t s[3];
		resolveTopology(d, k, s);
		for (int i = 0; i < nv; i++)
			vi[i] = fetchIndex(d, s[i]);
	}
	__syncwarp();

	/* 2. post-VS cache: de-duplicate the batch's indices, shade each distinct one once.
	 * Points bypass the cache like the reference (primitive_assembly.c:239-240). */
	uint32_t nUniq;
	if (nv > 1)
	{
		if (active)
			for (int i = 0; i < nv; i++)
				slot[i] = hashInsert(ws.hashKey, vi[i]);
		__syncwarp();
		/* number the occupied slots: lane l owns slots 4l..4l+3 */
		constexpr int PER = SRPD_HASH_SLOTS / 32;
		uint32_t keys[PER];
		uint32_t occ = 0;
		#pragma unroll
		for (int i = 0; i < PER; i++)
		{
			keys[i] = ws.hashKey[lane * PER + i];
			occ += keys[i] != SRPD_HASH_EMPTY;
		}
		const uint32_t inc = warpInclusiveScan(occ, lane);
		nUniq = __shfl_sync(0xFFFFFFFFu, inc, 31);
		uint32_t dense = inc - occ;
		#pragma unroll
		for (int i = 0; i < PER; i++)
			if (keys[i] != SRPD_HASH_EMPTY)
			{
				ws.hashDense[lane * PER + i] = (uint8_t) dense;
				ws.uniq[dense] = keys[i];
				dense++;
			}
		__syncwarp();
		for (int i = 0; i < nv; i++)
			slot[i] = ws.hashDense[slot[i]];
	}
	else
	{
		const uint32_t first = b * SRPD_GEOM_PRIMS;
		nUniq = min((uint32_t) SRPD_GEOM_PRIMS, d.nInputPrims - first);
		if (active) ws.uniq[lane] = vi[0];
		slot[0] = lane;
		__syncwarp();
	}

	for (uint32_t u = lane; u < nUniq; u += 32)
	{
		const uint32_t vertexIndex = ws.uniq[u];
		SRPVertexShaderIn in;
		in.uniform = (SRPUniform*) uniform;
		in.vertex = (SRPVertex*) (d.vb + (size_t) vertexIndex * d.vbStride);
		in.vertexID = vertexIndex;
		SRPVertexShaderOut out;
		out.clipPosition[0] = 0.f; out.clipPosition[1] = 0.f; out.clipPosition[2] = 0.f; out.clipPosition[3] = 0.f;
		out.varyings = (SRPVarying*) (vvary + (size_t) u * st.slotSize);
		srpB200DeviceVS(st.vsProgramId, &in, &out);
		ws.vpos[u] = make_float4(out.clipPosition[0], out.clipPosition[1], out.clipPosition[2], out.clipPosition[3]);
	}
	__syncwarp();

	/* 3. count */
	SrpdPos p[3];
	const unsigned char* vary[3];
	for (int i = 0; i < 3; i++)
	{
		const float4 q = ws.vpos[active && i < nv ? slot[i] : 0];
		p[i].x = q.x; p[i].y = q.y; p[i].z = q.z; p[i].w = q.w;
		vary[i] = vvary + (size_t) (active && i < nv ? slot[i] : 0) * st.slotSize;
	}
	Emitter em;
	em.a = &a;
	em.records = a.records + (size_t) frame * a.recCapacity * a.recStride;
	em.bboxes = a.bboxes + (size_t) frame * a.recCapacity;
	em.occupancy = a.occupancy + (size_t) frame * a.occWordsPerFrame;
	em.idBase = 0; em.storeBase = 0; em.nEmit = 0; em.nStore = 0; em.overflow = false;
	em.deferOcc = false; em.occWord = SRPD_OCC_NONE; em.occMask = 0u;
	em.frame = frame;
	FastTriangle fast;
	fast.valid = false; fast.stored = false;
	/* outcodes once, here: a primitive entirely outside one clip plane is rejected on the spot
	 * (clipping.c:75-76,146-147 -- the usual fate of most of a mesh that surrounds the camera)
	 * and only primitives that really cross a plane take the out-of-line path */
	uint32_t codeOr = 0u, codeAnd = 0u;
	if (nv >= 2)
	{
		const uint32_t c0 = srpdClipCode(p[0]), c1 = srpdClipCode(p[1]);
		const uint32_t c2 = nv == 3 ? srpdClipCode(p[2]) : c1;
		codeOr = c0 | c1 | c2;
		codeAnd = c0 & c1 & c2;
	}
	const bool needsClip = active && nv == 3 && codeAnd == 0u && codeOr != 0u;
	const uint32_t clipMask = __ballot_sync(0xFFFFFFFFu, needsClip);
	/* a line that is clipped or long (an estimate is enough: either pass emits the same records) */
	bool longLine = false;
	if (!CLIPPER && nv == 2 && active && codeAnd == 0u)
	{
		longLine = codeOr != 0u;
		if (!longLine)
		{
			const float ax = __fdividef(p[0].x, p[0].w), ay = __fdividef(p[0].y, p[0].w);
			const float bx = __fdividef(p[1].x, p[1].w), by = __fdividef(p[1].y, p[1].w);
			const float len = fmaxf(fabsf(bx - ax) * (float) st.width, fabsf(by - ay) * (float) st.height) * 0.5f;
			longLine = !(len <= (float) (SRPD_LINE_SHORT_SEGS * SRPD_LINE_SEG));
		}
	}
	if (!CLIPPER && (clipMask != 0u || __any_sync(0xFFFFFFFFu, longLine)))
	{
		/* a triangle of this batch crosses a clip plane, or one of its lines is long: the whole
		 * batch goes to the clipper pass */
		if (lane == 0)
			a.deferList[atomicAdd(a.deferCount, 1u)] = batch;
		continue;
	}
	constexpr bool LINES_BY_WARP = CLIPPER;      /* emitPreparedLines */
	if (LINES_BY_WARP && nv == 2)
	{
		LineSlot& sl = *reinterpret_cast<LineSlot*>(clipSlots + (size_t) lane * clipSlotStride(st.slotSize));
		sl.valid = 0u;
		if (active && codeAnd == 0u)
		{
			const SrpdPos sp[2] = { p[0], p[1] };
			const unsigned char* const sv[2] = { vary[0], vary[1] };
			prepareLine(st, sp, sv, sl);
		}
		__syncwarp();
		emitPreparedLines<false>(em, st, clipSlots, clipSlotStride(st.slotSize), lane);
	}
	else if (active && codeAnd == 0u && !needsClip)
	{
		if (nv == 3 && st.polygonMode == SRP_POLYGON_MODE_FILL)
		{
			/* unclipped filled triangle: set up once, remember the result for the write phase */
			fast.valid = true;
			if (setupAndClassify(st, p, fast.s, fast.stored))
			{
				em.nEmit = 1;
				em.nStore = fast.stored ? 1 : 0;
			}
		}
		else
		{
			/* copies whose addresses escape into the out-of-line path; `em`, `p` and `vary` stay in registers */
			Emitter slow = em;
			const SrpdPos sp[3] = { p[0], p[1], p[2] };
			const unsigned char* const sv[3] = { vary[0], vary[1], vary[2] };
			processPrimitive<false>(slow, d, nv, sp, sv);
			em.nEmit = slow.nEmit; em.nStore = slow.nStore;
		}
	}
	/* triangles that cross a clip plane (clipChunk): clipped and counted now, in slots that stay
	 * as they are until the batch has written its records */
	if (CLIPPER && clipMask != 0u)
	{
		const ClipResult r = clipChunk(em, st, SRPD_CLIP_COUNT, clipSlots, fanCache, clipMask, ws.vpos, vvary, slot[0], slot[1], slot[2], 0u, 0u);
		if (needsClip)
		{
			em.nEmit = r.nEmit; em.nStore = r.nStore;
		}
	}
	const uint32_t myEmit = em.nEmit, myStore = em.nStore;

	/* 4. warp scan of both counts.  No warp ever waits for another one: the batch takes a
	 * contiguous range of record slots from the frame's bump allocator (arrival order) and leaves
	 * its counts behind; the batch-order prefix sums and the id-ordered view of the records are
	 * built afterwards (srpdBatchOrderKernel). */
	const uint32_t incE = warpInclusiveScan(myEmit, lane);
	const uint32_t incS = warpInclusiveScan(myStore, lane);
	const uint32_t totE = __shfl_sync(0xFFFFFFFFu, incE, 31), totS = __shfl_sync(0xFFFFFFFFu, incS, 31);
	uint32_t physBase = 0u;
	if (lane == 0)
	{
		physBase = totS ? atomicAdd(&a.frameBump[frame], totS) : 0u;
		a.batchInfo[batch] = make_uint4(physBase, totE, totS, 0u);
		/* totals of the scan chunk this batch belongs to (srpdBatchOrderKernel adds up the chunks before its own) */
		uint2* cs = a.chunkSums + (size_t) frame * a.chunksPerFrame + b / SRPD_SCAN_CHUNK;
		if (totE) atomicAdd(&cs->x, totE);
		if (totS) atomicAdd(&cs->y, totS);
	}
	if (totS == 0u)
		continue;
	physBase = __shfl_sync(0xFFFFFFFFu, physBase, 0);

	/* 5. write, in id order (batch-local ids) */
	em.idBase = incE - myEmit;
	em.storeBase = physBase + incS - myStore;
	if (CLIPPER && clipMask != 0u)
	{
		const bool ovf = clipChunk(em, st, SRPD_CLIP_WRITE, clipSlots, fanCache, clipMask, ws.vpos, vvary, slot[0], slot[1], slot[2], em.idBase, em.storeBase).overflow;
		if (ovf)
		{
			atomicAdd(&a.stats->overflow, 1ull);
			atomicExch(a.abortFlag, 1u);
		}
	}
	if (LINES_BY_WARP && nv == 2)
	{
		em.nEmit = 0; em.nStore = 0;
		emitPreparedLines<true>(em, st, clipSlots, clipSlotStride(st.slotSize), lane);
		if (em.overflow)
		{
			atomicAdd(&a.stats->overflow, 1ull);
			atomicExch(a.abortFlag, 1u);
		}
	}
	else if (active && myStore > 0 && !needsClip)
	{
		em.nEmit = 0; em.nStore = 0;
		if (fast.valid)
		{
			em.deferOcc = true;
			writeTriangle<true>(em, st, fast.s, fast.stored, vary);
		}
		else
		{
			Emitter slow = em;
			const SrpdPos sp[3] = { p[0], p[1], p[2] };
			const unsigned char* const sv[3] = { vary[0], vary[1], vary[2] };
			processPrimitive<true>(slow, d, nv, sp, sv);
			em.overflow = slow.overflow;
		}
		if (em.overflow)
		{
			atomicAdd(&a.stats->overflow, 1ull);
			atomicExch(a.abortFlag, 1u);
		}
	}
	/* occupancy bits of the batch's unclipped filled triangles: the triangles of a batch are
	 * neighbours, so the lanes' bits fall into a word or two -- one reduction per distinct word and
	 * nothing to wait for (a load-and-test per record stalled every batch for an L2 round trip) */
	{
		const uint32_t peers = __match_any_sync(0xFFFFFFFFu, em.occWord);
		const uint32_t bits = __reduce_or_sync(peers, em.occMask);
		if (em.occWord != SRPD_OCC_NONE && lane == __ffs(peers) - 1)
			atomicOr(&em.occupancy[em.occWord], bits);
	}
	}   /* batch */
}

/* Batch order -> primitive order, one CTA per chunk of SRPD_SCAN_CHUNK batches of a frame.
 * (1) Exclusive prefix sums of (ids, records) over the batches, in batch order: this is the
 * reference's serial `primitiveID++`.  The geometry warps have already accumulated every
 * chunk's totals, so a CTA adds up the chunks in front of its own and scans its own batches
 * (one per thread, warp-shuffle scans).  (2) The id-ordered view: the chunk's records occupy a
 * contiguous range of positions in primitive order; the threads stride over those positions
 * (an even split whatever the batches' record counts -- most batches of a mesh around the
 * camera store nothing), find the owning batch in the chunk's prefix array in shared memory and
 * write {bounding box, record slot, id prefix of the batch} at its position: 16 bytes, write-only
 * (the records keep their batch-local ids; the tile kernel adds the prefix when it shades).
 * Binning and the tile kernel only ever walk this view. */
constexpr int SRPD_ORDER_THREADS = 1024;      /* the first SRPD_SCAN_CHUNK of them scan; all of them move records */
__global__ void __launch_bounds__(SRPD_ORDER_THREADS)
srpdBatchOrderKernel(const __grid_constant__ SrpdGeomArgs a)
{
	constexpr int WARPS = SRPD_SCAN_CHUNK / 32;
	__shared__ uint32_t sWarpE[WARPS], sWarpS[WARPS];
	__shared__ uint32_t sBase[2];
	__shared__ uint32_t sOrd[SRPD_SCAN_CHUNK + 1];      /* records in front of the batch within the chunk; [CHUNK] = the chunk's total */
	__shared__ uint32_t sPhys[SRPD_SCAN_CHUNK];         /* the batch's first record slot */
	__shared__ uint32_t sIds[SRPD_SCAN_CHUNK];          /* primitive ids in front of the batch within the frame */
	srpdGridDependencyEnter();
	const uint32_t frame = blockIdx.x / a.chunksPerFrame, chunk = blockIdx.x - frame * a.chunksPerFrame;
	const uint32_t first = frame * a.batchesPerFrame;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint2* cs = a.chunkSums + (size_t) frame * a.chunksPerFrame;
	if (warp == 0)
	{
		uint32_t pe = 0, ps = 0;
		for (uint32_t c = lane; c < chunk; c += 32)
		{
			const uint2 v = cs[c];
			pe += v.x; ps += v.y;
		}
		pe = __reduce_add_sync(0xFFFFFFFFu, pe);
		ps = __reduce_add_sync(0xFFFFFFFFu, ps);
		if (lane == 0) { sBase[0] = pe; sBase[1] = ps; }
	}
	const bool scanner = tid < SRPD_SCAN_CHUNK;      /* (whole warps) */
	const uint32_t b = chunk * SRPD_SCAN_CHUNK + tid;
	uint32_t e = 0, s = 0, phys0 = 0;
	if (scanner && b < a.batchesPerFrame)
	{
		const uint4 info = a.batchInfo[first + b];
		phys0 = info.x; e = info.y; s = info.z;
	}
	uint32_t incE = 0, incS = 0;
	if (scanner)
	{
		incE = warpInclusiveScan(e, lane); incS = warpInclusiveScan(s, lane);
		if (lane == 31) { sWarpE[warp] = incE; sWarpS[warp] = incS; }
	}
	__syncthreads();
	uint32_t inE = incE - e, inS = incS - s;      /* exclusive, within the chunk */
	if (scanner)
		for (int w = 0; w < warp; w++) { inE += sWarpE[w]; inS += sWarpS[w]; }
	/* a later sub-draw of a split draw continues the ids of the previous one (draw_types.h) */
	const uint32_t carry = a.d.chunkIndex ? a.idCarry[frame] : 0u;
	const uint32_t baseE = sBase[0] + carry, baseS = sBase[1];
	if (scanner)
	{
		sOrd[tid] = inS;
		sPhys[tid] = phys0;
		sIds[tid] = baseE + inE;
	}
	if (tid == SRPD_SCAN_CHUNK - 1)
	{
		sOrd[SRPD_SCAN_CHUNK] = inS + s;
		if (chunk == a.chunksPerFrame - 1)
		{
			/* the frame's totals (the last thread of the last chunk has seen everything) */
			const uint32_t pe = baseE + inE + e, ps = baseS + inS + s;
			a.frameCounts[2 * frame + 0] = pe;
			a.frameCounts[2 * frame + 1] = ps < a.recCapacity ? ps : a.recCapacity;
			a.idCarryOut[frame] = pe;
			if (ps > a.recCapacity)
				atomicMax(&a.needed[0], ps);
			atomicAdd(&a.stats->primsIn, (unsigned long long) a.d.nInputPrims);
			atomicAdd(&a.stats->primsEmitted, (unsigned long long) (pe - carry));
			atomicAdd(&a.stats->primsStored, (unsigned long long) ps);
		}
	}
	__syncthreads();
	const uint32_t total = sOrd[SRPD_SCAN_CHUNK];
	const size_t base = (size_t) frame * a.recCapacity;
	for (uint32_t j = tid; j < total; j += SRPD_ORDER_THREADS)
	{
		/* the owning batch: the last one with sOrd[t] <= j (batches without records share their
		 * successor's value and are stepped over) */
		uint32_t lo = 0, hi = SRPD_SCAN_CHUNK;
		#pragma unroll
		for (int step = 0; step < 8; step++)      /* log2(SRPD_SCAN_CHUNK) */
		{
			const uint32_t mid = (lo + hi) >> 1;
			if (sOrd[mid] <= j) lo = mid; else hi = mid;
		}
		const uint32_t phys = sPhys[lo] + (j - sOrd[lo]), ord = baseS + j;
		if (phys >= a.recCapacity || ord >= a.recCapacity)
			continue;      /* (cannot happen: the pools hold the worst case; kept as a guard) */
		/* write-only: the record keeps its batch-local id, the view carries the batch's id prefix */
		const uint2 bb = a.bboxes[base + phys];
		a.ordered[base + ord] = make_uint4(bb.x, bb.y, phys, sIds[lo]);
	}
}
static_assert(SRPD_SCAN_CHUNK == 256, "srpdBatchOrderKernel: one batch per thread, 8 search steps");

static int gGeomLaunches = 0;
int srpdGeomLaunchCount(void) { return gGeomLaunches; }

template <bool BATCH, bool CLIPPER>
static void launchGeomKernel(const SrpdGeomArgs& a, unsigned grid, cudaStream_t stream)
{
	static bool configured = false;
	if (!configured)
	{
		cudaFuncSetAttribute(srpdGeomKernel<BATCH, CLIPPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		configured = true;
	}
	srpdLaunchKernel(srpdGeomKernel<BATCH, CLIPPER>, grid, CLIPPER ? 32 : SRPD_GEOM_THREADS, geomCtaBytes(a.d.st.slotSize, CLIPPER), stream, a);
}

/* Small draws: one launch, every batch by a clipper warp.  Large draws: the main pass, then the
 * clipper pass over the batches the main pass deferred (their number is only known on the device:
 * the grid is sized for the machine and strides through the list).  Returns the launches made. */
int srpdLaunchGeom(const SrpdGeomArgs& a0, cudaStream_t stream)
{
	SrpdGeomArgs a = a0;
	const unsigned batches = a.batchesPerFrame * a.d.nFrames;
	int launches = 0;
	if (batches <= SRPD_GEOM_SMALL_DRAW_BATCHES)
	{
		a.deferred = 0;
		if (a.frames) launchGeomKernel<true, true>(a, batches, stream);
		else          launchGeomKernel<false, true>(a, batches, stream);
		launches = 1;
	}
	else
	{
		a.deferred = 0;
		const unsigned grid = (batches + SRPD_GEOM_WARPS - 1) / SRPD_GEOM_WARPS;
		if (a.frames) launchGeomKernel<true, false>(a, grid, stream);
		else          launchGeomKernel<false, false>(a, grid, stream);
		a.deferred = 1;
		unsigned clipGrid = a.smCount * 16u;
		if (clipGrid > batches) clipGrid = batches;
		if (a.frames) launchGeomKernel<true, true>(a, clipGrid, stream);
		else          launchGeomKernel<false, true>(a, clipGrid, stream);
		launches = 2;
	}
	srpdLaunchKernel(srpdBatchOrderKernel, a.d.nFrames * a.chunksPerFrame, SRPD_ORDER_THREADS, 0, stream, a);
	launches += 1;
	gGeomLaunches += launches;
	return launches;
}
