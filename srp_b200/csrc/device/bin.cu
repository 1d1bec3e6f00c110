/* srp-b200 -- sort-middle coarse binning that preserves primitive order (sm_100a).
 *
 * The reference has no binning (it rasterises primitive-at-a-time, src/pipeline/draw.c:
 * 101-122); this stage exists so that a tile CTA only looks at the primitives near it
 * while still visiting them in primitive-id order.  The id-ordered view (srpdBatchOrderKernel)
 * lists the stored primitives in primitive order, so "in order" means "in increasing position".
 *
 * Bins are supertiles of 2^k x 2^k tiles (k = 3: 256x128 px, down to k = 1 for dense
 * sub-pixel geometry; chosen per draw).  Three passes over the bounding boxes of the id-ordered
 * view (16-byte entries, 8 bytes of box each; the records themselves are not touched):
 *   count : one CTA per chunk of 2048 records counts, per supertile, how many of the
 *           chunk's records overlap it            -> chunkCounts[chunk][supertile]
 *   scan  : per supertile an exclusive scan over the chunks, then (by the CTA that finishes
 *           last) an exclusive scan over the supertiles' totals
 *                                                 -> superOffsets[s] + chunkCounts[c][s] is
 *                                                    the write cursor of (chunk c, supertile s)
 *   fill  : one CTA per chunk, each warp walks its share of the records 32 at a time; a
 *           record's slot in a supertile list is cursor + (number of EARLIER lanes whose box
 *           also covers that supertile) -- a warp-scan compaction, evaluated through
 *           ballot/match for the common one-supertile case -- so every list comes out sorted
 *           by record index without global atomics or a sort.
 * A tile then filters its supertile's list against its own rectangle in shared memory
 * (raster.cu). */
#include "kernels.cuh"

namespace {

/* inclusive supertile rectangle of a pixel bbox: sx0 | sy0<<8 | sx1<<16 | sy1<<24 */
__device__ __forceinline__ uint32_t superRect(uint2 bb, uint32_t superX, uint32_t superY, uint32_t superShift)
{
	const uint32_t x0 = bb.x & 0xFFFFu, y0 = bb.x >> 16, x1 = bb.y & 0xFFFFu, y1 = bb.y >> 16;
	if (x1 <= x0 || y1 <= y0)
		return 0x00000101u;   /* empty: sx0 = 1 > sx1 = 0 */
	const uint32_t SW = (uint32_t) SRPD_TILE_W << superShift, SH = (uint32_t) SRPD_TILE_H << superShift;
	uint32_t sx0 = x0 / SW, sy0 = y0 / SH, sx1 = (x1 - 1) / SW, sy1 = (y1 - 1) / SH;
	if (sx1 >= superX) sx1 = superX - 1;
	if (sy1 >= superY) sy1 = superY - 1;
	return sx0 | (sy0 << 8) | (sx1 << 16) | (sy1 << 24);
}
__device__ __forceinline__ bool rectEmpty(uint32_t r) { return (r & 0xFFu) > ((r >> 16) & 0xFFu); }

/* Records per chunk: SRPD_BIN_CHUNK, or a quarter of it while the draw stores so few records that
 * full chunks would leave most of the machine idle (cfg3: 130 k records = 64 chunks of 2048 on 148
 * SMs).  Every binning kernel derives it from the same record count. */
__device__ __forceinline__ uint32_t binChunkRecords(uint32_t nStored, uint32_t nSuper)
{
	if (nSuper > SRPD_BIN_SMEM_SUPERS)
	{
		/* many supertiles (dense draws of small primitives): the fill runs one WARP per chunk
		 * (srpdBinFillWarpKernel); at least 512 records per chunk, at most SRPD_BIN_WARP_CHUNKS chunks */
		uint32_t r = (nStored + SRPD_BIN_WARP_CHUNKS - 1) / SRPD_BIN_WARP_CHUNKS;
		r = r < 512u ? 512u : r;
		return (r + 31u) & ~31u;
	}
	return nStored <= SRPD_BIN_SMALL_RECORDS ? SRPD_BIN_CHUNK / 4 : SRPD_BIN_CHUNK;
}

} // namespace

__global__ void __launch_bounds__(SRPD_BIN_THREADS)
srpdBinCountKernel(const __grid_constant__ SrpdBinArgs a)
{
	extern __shared__ uint32_t sCount[];
	srpdGridDependencyEnter();
	const uint32_t nSuper = a.superX * a.superY;
	const uint32_t nStored = a.frameCounts[1];
	const uint32_t chunkRecords = binChunkRecords(nStored, nSuper);
	const uint32_t nChunks = (nStored + chunkRecords - 1) / chunkRecords;
	/* the grid is sized for the machine, not for the record capacity: the CTAs stride over the
	 * chunks that exist (nobody reads the counts of chunks past the end) */
	for (uint32_t chunk = blockIdx.x; chunk < nChunks; chunk += gridDim.x)
	{
		for (uint32_t s = threadIdx.x; s < nSuper; s += SRPD_BIN_THREADS)
			sCount[s] = 0;
		const uint32_t first = chunk * chunkRecords;
		__syncthreads();      /* (the counters are zero) */
		for (uint32_t part = 0; part < chunkRecords; part += SRPD_BIN_CHUNK)
		{
			/* all of the thread's boxes are requested before the first one is used (the loads overlap) */
			constexpr int PER = SRPD_BIN_CHUNK / SRPD_BIN_THREADS;
			uint2 bb[PER];
			#pragma unroll
			for (int k = 0; k < PER; k++)
			{
				const uint32_t o = part + k * SRPD_BIN_THREADS + threadIdx.x;
				bb[k] = (o < chunkRecords && first + o < nStored) ? *(const uint2*) (a.ordered + first + o) : make_uint2(0u, 0u);      /* an empty box */
			}
			#pragma unroll
			for (int k = 0; k < PER; k++)
			{
				const uint32_t rect = superRect(bb[k], a.superX, a.superY, a.superShift);
				if (rectEmpty(rect))
					continue;
				for (uint32_t sy = (rect >> 8) & 0xFFu; sy <= (rect >> 24); sy++)
					for (uint32_t sx = rect & 0xFFu; sx <= ((rect >> 16) & 0xFFu); sx++)
						atomicAdd(&sCount[sy * a.superX + sx], 1u);
			}
		}
		__syncthreads();
		for (uint32_t s = threadIdx.x; s < nSuper; s += SRPD_BIN_THREADS)
			a.chunkCounts[(size_t) chunk * nSuper + s] = sCount[s];
		__syncthreads();
	}
}

/* Scan pass 2 (run by the last CTA of pass 1): exclusive scan of the supertile totals ->
 * superOffsets (warp-shuffle scans: lanes, then the warp totals, one barrier pair per 256 supertiles) */
constexpr int SRPD_SCAN_COL_WARPS = 16;
__device__ __forceinline__ void scanSuperTotals(const SrpdBinArgs& a, uint32_t nSuper)
{
	__shared__ uint32_t sWarp[SRPD_SCAN_COL_WARPS];
	__shared__ uint32_t sCarry;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	constexpr uint32_t T = SRPD_SCAN_COL_WARPS * 32;
	if (threadIdx.x == 0)
		sCarry = 0;
	__syncthreads();
	for (uint32_t s0 = 0; s0 < nSuper; s0 += T)
	{
		const uint32_t s = s0 + threadIdx.x;
		const uint32_t run = s < nSuper ? __ldcg(a.superTotals + s) : 0;      /* (written by other CTAs: not through L1) */
		uint32_t v = run;
		#pragma unroll
		for (uint32_t o = 1; o < 32; o <<= 1)
		{
			const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, v, o);
			if (lane >= o) v += n;
		}
		if (lane == 31)
			sWarp[warp] = v;
		__syncthreads();
		uint32_t before = 0, all = 0;
		#pragma unroll
		for (int w = 0; w < SRPD_SCAN_COL_WARPS; w++)
		{
			const uint32_t t = sWarp[w];
			if ((uint32_t) w < warp) before += t;
			all += t;
		}
		const uint32_t carry = sCarry;
		const uint32_t excl = carry + before + v - run;
		if (s < nSuper)
			a.superOffsets[s] = excl < a.listCapacity ? excl : a.listCapacity;
		__syncthreads();
		if (threadIdx.x == 0)
			sCarry = carry + all;
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		const uint32_t total = sCarry;
		a.superOffsets[nSuper] = total < a.listCapacity ? total : a.listCapacity;
		if (total > a.listCapacity)
		{
			/* the coarse lists do not fit: this draw's tiles scan all records instead (slow but
			 * exact, no host round trip), and the host is told how large the pool has to be */
			*a.listOverflow = 1u;
			a.needed[1] = total;
			*(volatile uint32_t*) (a.hostNotes + 1) = total;
			__threadfence_system();
		}
	}
}

/* Scan pass 1: exclusive scan over the chunks for every supertile column, total per supertile.
 * A CTA takes 32 columns (one per lane: coalesced rows); its 8 warps split the chunk rows, sum
 * their parts (loads only, 32 in flight), exchange the 8 partial sums per column through shared
 * memory and then rewrite their rows with the running sums -- the serial depth is two batches of
 * loads per 256 chunks instead of one dependent batch per 8 (or 32) chunks. */
__global__ void __launch_bounds__(SRPD_SCAN_COL_WARPS * 32)
srpdBinScanColumnsKernel(const __grid_constant__ SrpdBinArgs a)
{
	__shared__ uint32_t sPart[SRPD_SCAN_COL_WARPS][32];
	srpdGridDependencyEnter();
	const uint32_t nSuper = a.superX * a.superY;
	const uint32_t nStored = a.frameCounts[1];
	const uint32_t chunkRecords = binChunkRecords(nStored, nSuper);
	const uint32_t nChunks = (nStored + chunkRecords - 1) / chunkRecords;
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t s = blockIdx.x * 32u + lane;
	const bool valid = s < nSuper;
	const uint32_t rows = (nChunks + SRPD_SCAN_COL_WARPS - 1) / SRPD_SCAN_COL_WARPS;
	const uint32_t c0 = min(warp * rows, nChunks), c1 = min(c0 + rows, nChunks);
	uint32_t* col = a.chunkCounts + s;

	uint32_t sum = 0;
	if (valid)
	{
		uint32_t c = c0;
		for (; c + 32 <= c1; c += 32)
		{
			uint32_t n[32];
			#pragma unroll
			for (int u = 0; u < 32; u++)
				n[u] = col[(size_t) (c + u) * nSuper];
			#pragma unroll
			for (int u = 0; u < 32; u++)
				sum += n[u];
		}
		for (; c < c1; c++)
			sum += col[(size_t) c * nSuper];
	}
	sPart[warp][lane] = sum;
	__syncthreads();
	uint32_t run = 0, total = 0;
	#pragma unroll
	for (int w = 0; w < SRPD_SCAN_COL_WARPS; w++)
	{
		const uint32_t p = sPart[w][lane];
		if ((uint32_t) w < warp) run += p;
		total += p;
	}
	if (valid)
	{
		uint32_t c = c0;
		for (; c + 32 <= c1; c += 32)
		{
			uint32_t n[32];
			#pragma unroll
			for (int u = 0; u < 32; u++)
				n[u] = col[(size_t) (c + u) * nSuper];
			#pragma unroll
			for (int u = 0; u < 32; u++)
			{
				col[(size_t) (c + u) * nSuper] = run;
				run += n[u];
			}
		}
		for (; c < c1; c++)
		{
			const uint32_t n = col[(size_t) c * nSuper];
			col[(size_t) c * nSuper] = run;
			run += n;
		}
		if (warp == 0)
			a.superTotals[s] = total;
	}
	/* the CTA that finishes last turns the supertile totals into the list offsets (this used to be
	 * a kernel of its own: one launch and one drain less in every binned draw) */
	__shared__ uint32_t sLast;
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0)
		sLast = atomicAdd(a.scanTicket, 1u) == gridDim.x - 1 ? 1u : 0u;
	__syncthreads();
	if (!sLast)
		return;
	__threadfence();
	scanSuperTotals(a, nSuper);
}

/* Fill: one CTA per chunk, FILL_WARPS warps; warp w owns the chunk's records
 * [w*SPAN, (w+1)*SPAN) with SPAN = SRPD_BIN_CHUNK / FILL_WARPS.
 *   1. every lane loads its SPAN/32 bounding boxes up front (the loads overlap) and the warp
 *      counts its records per supertile                     -> sCnt[w][s]
 *   2. exclusive scan over the warps for every supertile, seeded with the chunk's global
 *      cursor                                               -> sCnt[w][s] = first slot of warp w
 *   3. every warp walks its records 32 at a time in record order: the lanes of a step OR
 *      their lane bit into a per-supertile mask, a record's slot is the warp cursor + the
 *      population count of the mask below its own lane (a warp-scan compaction keyed by
 *      supertile), then the last covering lane advances the cursor.
 * Lists therefore come out sorted by record index with no atomics on global memory. */
template <int FILL_WARPS>
__global__ void __launch_bounds__(FILL_WARPS * 32, 1)
srpdBinFillKernel(const __grid_constant__ SrpdBinArgs a)
{
	constexpr int ROUNDS = SRPD_BIN_CHUNK / FILL_WARPS / 32;      /* steps of 32 per warp, at most */
	extern __shared__ uint32_t sFill[];                   /* [FILL_WARPS][nSuper] cursors, then [FILL_WARPS][nSuper] lane masks */
	srpdGridDependencyEnter();
	const uint32_t nSuper = a.superX * a.superY;
	const uint32_t nStored = a.frameCounts[1];
	const uint32_t chunkRecords = binChunkRecords(nStored, nSuper);
	const uint32_t nChunks = (nStored + chunkRecords - 1) / chunkRecords;
	const uint32_t span = chunkRecords / FILL_WARPS;              /* records per warp */
	const int rounds = (int) (span / 32);
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	uint32_t* sCnt = sFill + (size_t) warp * nSuper;
	uint32_t* sMask = sFill + (size_t) (FILL_WARPS + warp) * nSuper;
	for (uint32_t chunk = blockIdx.x; chunk < nChunks; chunk += gridDim.x)      /* grid sized for the machine */
	{
	const uint32_t first = chunk * chunkRecords;

	for (uint32_t i = tid; i < 2 * FILL_WARPS * nSuper; i += FILL_WARPS * 32)
		sFill[i] = 0;
	uint32_t rect[ROUNDS];
	const uint32_t warpFirst = first + warp * span;
	#pragma unroll
	for (int r = 0; r < ROUNDS; r++)
	{
		const uint32_t rec = warpFirst + r * 32 + lane;
		rect[r] = (r < rounds && rec < nStored) ? superRect(*(const uint2*) (a.ordered + rec), a.superX, a.superY, a.superShift) : 0x00000101u;
	}
	__syncthreads();
	#pragma unroll
	for (int r = 0; r < ROUNDS; r++)
		if (!rectEmpty(rect[r]))
			for (uint32_t sy = (rect[r] >> 8) & 0xFFu; sy <= (rect[r] >> 24); sy++)
				for (uint32_t sx = rect[r] & 0xFFu; sx <= ((rect[r] >> 16) & 0xFFu); sx++)
					atomicAdd(&sCnt[sy * a.superX + sx], 1u);
	__syncthreads();
	for (uint32_t s = tid; s < nSuper; s += FILL_WARPS * 32)
	{
		uint32_t run = a.superOffsets[s] + a.chunkCounts[(size_t) chunk * nSuper + s];
		for (int w = 0; w < FILL_WARPS; w++)
		{
			const uint32_t c = sFill[(size_t) w * nSuper + s];
			sFill[(size_t) w * nSuper + s] = run;
			run += c;
		}
	}
	__syncthreads();

	#pragma unroll
	for (int r = 0; r < ROUNDS; r++)
	{
		const uint32_t rec = warpFirst + r * 32 + lane;
		const uint32_t rc = rect[r];
		const bool empty = rectEmpty(rc);
		if (__ballot_sync(0xFFFFFFFFu, !empty) == 0)
			continue;
		const uint32_t sx0 = rc & 0xFFu, sy0 = (rc >> 8) & 0xFFu, sx1 = (rc >> 16) & 0xFFu, sy1 = rc >> 24;
		/* a list entry is the record's entry of the ordered view itself (box, slot, id prefix): the
		 * tile kernel's scan of a list is then one coalesced stream, not a gather behind an index */
		uint4 ent = make_uint4(0u, 0u, 0u, 0u);
		if (!empty)
			ent = a.ordered[rec];
		/* which lanes of this step cover supertile s: one bit per lane */
		if (!empty)
			for (uint32_t sy = sy0; sy <= sy1; sy++)
				for (uint32_t sx = sx0; sx <= sx1; sx++)
					atomicOr(&sMask[sy * a.superX + sx], 1u << lane);
		__syncwarp();
		/* slot = cursor + number of EARLIER lanes covering the same supertile */
		if (!empty)
			for (uint32_t sy = sy0; sy <= sy1; sy++)
				for (uint32_t sx = sx0; sx <= sx1; sx++)
				{
					const uint32_t s = sy * a.superX + sx;
					const uint32_t pos = sCnt[s] + __popc(sMask[s] & ((1u << lane) - 1u));
					if (pos < a.listCapacity)
						a.listEntries[pos] = ent;
				}
		__syncwarp();
		/* the last covering lane advances the cursor (nobody reads the cursors in this phase) ... */
		if (!empty)
			for (uint32_t sy = sy0; sy <= sy1; sy++)
				for (uint32_t sx = sx0; sx <= sx1; sx++)
				{
					const uint32_t s = sy * a.superX + sx;
					const uint32_t m = sMask[s];
					if ((m >> lane) == 1u)
						sCnt[s] += __popc(m);
				}
		__syncwarp();
		/* ... and every lane takes its own bit out of the masks again for the next step (atomically:
		 * the covering lanes of a supertile do it side by side) */
		if (!empty)
			for (uint32_t sy = sy0; sy <= sy1; sy++)
				for (uint32_t sx = sx0; sx <= sx1; sx++)
					atomicAnd(&sMask[sy * a.superX + sx], ~(1u << lane));
		__syncwarp();
	}
	__syncthreads();      /* the next chunk re-initialises the cursors */
	}   /* chunk */
}

/* Fill for draws with MANY supertiles (dense small primitives: cfg4's 10 M sub-pixel triangles bin
 * into 4080 supertiles), where a per-warp cursor matrix no longer fits shared memory.  One WARP
 * per chunk; its cursors are its own row of chunkCounts in global memory (after the scans:
 * records of earlier chunks per supertile), advanced in place.  The (record, supertile) pairs of a
 * step of 32 records are taken in (lane, cell) order, 32 pairs at a time: match.any groups the
 * pairs of a supertile, a pair's slot is the group's cursor + its rank in the group -- the same
 * warp-scan compaction keyed by supertile, without a mask per supertile -- and the group's first
 * lane advances the cursor before the next 32 pairs.  Lists come out sorted by record position. */
__global__ void __launch_bounds__(SRPD_BIN_THREADS)
srpdBinFillWarpKernel(const __grid_constant__ SrpdBinArgs a)
{
	srpdGridDependencyEnter();
	const uint32_t nSuper = a.superX * a.superY;
	const uint32_t nStored = a.frameCounts[1];
	const uint32_t chunkRecords = binChunkRecords(nStored, nSuper);
	const uint32_t nChunks = (nStored + chunkRecords - 1) / chunkRecords;
	const int lane = threadIdx.x & 31;
	const uint32_t warpsPerGrid = gridDim.x * (SRPD_BIN_THREADS / 32);
	const uint32_t lt = (1u << lane) - 1u;
	for (uint32_t chunk = blockIdx.x * (SRPD_BIN_THREADS / 32) + (threadIdx.x >> 5); chunk < nChunks; chunk += warpsPerGrid)
	{
		uint32_t* cursor = a.chunkCounts + (size_t) chunk * nSuper;
		const uint32_t first = chunk * chunkRecords;
		const uint32_t last = min(first + chunkRecords, nStored);
		for (uint32_t r0 = first; r0 < last; r0 += 32)
		{
			const uint32_t rec = r0 + lane;
			const uint32_t rc = rec < last ? superRect(*(const uint2*) (a.ordered + rec), a.superX, a.superY, a.superShift) : 0x00000101u;
			const bool empty = rectEmpty(rc);
			const uint32_t sx0 = rc & 0xFFu, sy0 = (rc >> 8) & 0xFFu, sx1 = (rc >> 16) & 0xFFu, sy1 = rc >> 24;
			const uint32_t w = empty ? 0u : sx1 - sx0 + 1u, cells = empty ? 0u : w * (sy1 - sy0 + 1u);
			/* pairs in (lane, cell) order */
			uint32_t inc = cells;
			#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
				if (lane >= o) inc += v;
			}
			const uint32_t nPairs = __shfl_sync(0xFFFFFFFFu, inc, 31);
			const uint32_t excl = inc - cells;
			for (uint32_t p0 = 0; p0 < nPairs; p0 += 32)
			{
				const uint32_t p = p0 + lane;
				/* the owner of pair p: the last lane whose first pair is <= p (lanes without cells share
				 * their successor's value and are stepped over) */
				int lo = 0;
				#pragma unroll
				for (int step = 16; step > 0; step >>= 1)
				{
					const uint32_t e = __shfl_sync(0xFFFFFFFFu, excl, (lo + step) & 31);
					if (lo + step < 32 && e <= p) lo += step;
				}
				const uint32_t oExcl = __shfl_sync(0xFFFFFFFFu, excl, lo), oW = __shfl_sync(0xFFFFFFFFu, w, lo);
				const uint32_t oX0 = __shfl_sync(0xFFFFFFFFu, sx0, lo), oY0 = __shfl_sync(0xFFFFFFFFu, sy0, lo);
				const bool valid = p < nPairs;
				uint32_t s = 0xFFFFFFFFu - (uint32_t) lane;      /* a key nobody shares */
				if (valid)
				{
					const uint32_t cell = p - oExcl;
					s = (oY0 + cell / oW) * a.superX + oX0 + cell % oW;
				}
				const uint32_t peers = __match_any_sync(0xFFFFFFFFu, s);
				if (valid)
				{
					const int leader = __ffs(peers) - 1;
					uint32_t cur = 0u;
					if (lane == leader)
						cur = a.superOffsets[s] + cursor[s];
					cur = __shfl_sync(peers, cur, leader);
					const uint32_t pos = cur + __popc(peers & lt);
					if (pos < a.listCapacity)
						a.listEntries[pos] = a.ordered[r0 + (uint32_t) lo];
					if (lane == leader)
						cursor[s] += __popc(peers);
				}
				__syncwarp();      /* the next 32 pairs read the advanced cursors */
			}
		}
	}
}

static int gBinLaunches = 0;
int srpdBinLaunchCount(void) { return gBinLaunches; }

void srpdLaunchBin(const SrpdBinArgs& a, cudaStream_t stream, cudaEvent_t joinBeforeFill)
{
	const uint32_t nSuper = a.superX * a.superY;
	uint32_t grid = a.smCount * 3u;      /* the CTAs stride over the chunks that turn out to exist */
	if (grid > a.nChunksMax) grid = a.nChunksMax;
	if (grid == 0) grid = 1;
	srpdLaunchKernel(srpdBinCountKernel, grid, SRPD_BIN_THREADS, nSuper * sizeof(uint32_t), stream, a);
	srpdLaunchKernel(srpdBinScanColumnsKernel, (nSuper + 31) / 32, SRPD_SCAN_COL_WARPS * 32, 0, stream, a);
	/* work of another stream that the kernel AFTER the fill needs (the checkpoint pre-pass, for
	 * the tiles) joins here: the fill then starts after it, and the tile kernel's programmatic
	 * dependency on the fill covers it */
	if (joinBeforeFill)
		cudaStreamWaitEvent(stream, joinBeforeFill, 0);
	/* as many warps per chunk as the cursor matrix allows in shared memory; beyond
	 * SRPD_BIN_SMEM_SUPERS supertiles one warp per chunk with its cursors in global memory */
	const size_t budget = 160 * 1024;
	if (nSuper > SRPD_BIN_SMEM_SUPERS)
	{
		uint32_t fillGrid = a.smCount * 4u;
		if (fillGrid > (SRPD_BIN_WARP_CHUNKS + 7u) / 8u) fillGrid = (SRPD_BIN_WARP_CHUNKS + 7u) / 8u;
		srpdLaunchKernel(srpdBinFillWarpKernel, fillGrid, SRPD_BIN_THREADS, 0, stream, a);
	}
	else if (2 * 8 * (size_t) nSuper * sizeof(uint32_t) <= budget / 2)
	{
		const size_t bytes = 2 * 8 * (size_t) nSuper * sizeof(uint32_t);
		static bool configured8 = false;
		if (!configured8) { cudaFuncSetAttribute(srpdBinFillKernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) budget); configured8 = true; }
		srpdLaunchKernel(srpdBinFillKernel<8>, grid, 8 * 32, bytes, stream, a);
	}
	else if (2 * 4 * (size_t) nSuper * sizeof(uint32_t) <= budget)
	{
		const size_t bytes = 2 * 4 * (size_t) nSuper * sizeof(uint32_t);
		static bool configured4 = false;
		if (!configured4) { cudaFuncSetAttribute(srpdBinFillKernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) budget); configured4 = true; }
		srpdLaunchKernel(srpdBinFillKernel<4>, grid, 4 * 32, bytes, stream, a);
	}
	else
	{
		const size_t bytes = 2 * 2 * (size_t) nSuper * sizeof(uint32_t);
		static bool configured2 = false;
		if (!configured2) { cudaFuncSetAttribute(srpdBinFillKernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) budget); configured2 = true; }
		srpdLaunchKernel(srpdBinFillKernel<2>, grid, 2 * 32, bytes, stream, a);
	}
	gBinLaunches += 3;
}
