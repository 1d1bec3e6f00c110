/* srp-b200 -- sort-middle coarse binning that preserves primitive order (sm_100a).
 *
 * The reference has no binning (it rasterises primitive-at-a-time, src/pipeline/draw.c:
 * 101-122); this stage exists so that a tile CTA only looks at the primitives near it
 * while still visiting them in primitive-id order.  Records are already stored in id
 * order by the geometry kernel, so "in order" means "in increasing record index".
 *
 * Bins are supertiles of 8x8 tiles (256x128 px).  Three passes over the 8-byte bounding
 * boxes (HBM-bound; the records themselves are not touched):
 *   count : one CTA per chunk of 2048 records counts, per supertile, how many of the
 *           chunk's records overlap it            -> chunkCounts[chunk][supertile]
 *   scan  : per supertile an exclusive scan over the chunks, then an exclusive scan over
 *           the supertiles' totals                -> superOffsets[s] + chunkCounts[c][s] is
 *                                                    the write cursor of (chunk c, supertile s)
 *   fill  : one warp per chunk walks its records 32 at a time; a record's slot in a
 *           supertile list is cursor + (number of EARLIER lanes whose box also covers that
 *           supertile) -- a warp-scan compaction, evaluated through ballot/match for the
 *           common one-supertile case -- so every list comes out sorted by record index
 *           without atomics or a sort.
 * A tile then filters its supertile's list against its own rectangle in shared memory
 * (raster.cu). */
#include "kernels.cuh"

namespace {

/* inclusive supertile rectangle of a pixel bbox: sx0 | sy0<<8 | sx1<<16 | sy1<<24 */
__device__ __forceinline__ uint32_t superRect(uint2 bb, uint32_t superX, uint32_t superY)
{
	const uint32_t x0 = bb.x & 0xFFFFu, y0 = bb.x >> 16, x1 = bb.y & 0xFFFFu, y1 = bb.y >> 16;
	if (x1 <= x0 || y1 <= y0)
		return 0x00000101u;   /* empty: sx0 = 1 > sx1 = 0 */
	constexpr uint32_t SW = SRPD_TILE_W * SRPD_SUPER_W, SH = SRPD_TILE_H * SRPD_SUPER_H;
	uint32_t sx0 = x0 / SW, sy0 = y0 / SH, sx1 = (x1 - 1) / SW, sy1 = (y1 - 1) / SH;
	if (sx1 >= superX) sx1 = superX - 1;
	if (sy1 >= superY) sy1 = superY - 1;
	return sx0 | (sy0 << 8) | (sx1 << 16) | (sy1 << 24);
}
__device__ __forceinline__ bool rectContains(uint32_t r, uint32_t sx, uint32_t sy)
{
	return sx >= (r & 0xFFu) && sx <= ((r >> 16) & 0xFFu) && sy >= ((r >> 8) & 0xFFu) && sy <= (r >> 24);
}
__device__ __forceinline__ bool rectEmpty(uint32_t r) { return (r & 0xFFu) > ((r >> 16) & 0xFFu); }
__device__ __forceinline__ bool rectSingle(uint32_t r)
{
	return (r & 0xFFu) == ((r >> 16) & 0xFFu) && ((r >> 8) & 0xFFu) == (r >> 24);
}

} // namespace

__global__ void __launch_bounds__(SRPD_BIN_THREADS)
srpdBinCountKernel(const __grid_constant__ SrpdBinArgs a)
{
	extern __shared__ uint32_t sCount[];
	const uint32_t nSuper = a.superX * a.superY;
	const uint32_t nStored = a.frameCounts[1];
	for (uint32_t s = threadIdx.x; s < nSuper; s += SRPD_BIN_THREADS)
		sCount[s] = 0;
	__syncthreads();
	const uint32_t first = blockIdx.x * SRPD_BIN_CHUNK;
	for (uint32_t o = threadIdx.x; o < SRPD_BIN_CHUNK; o += SRPD_BIN_THREADS)
	{
		const uint32_t r = first + o;
		if (r >= nStored)
			break;
		const uint32_t rect = superRect(a.bboxes[r], a.superX, a.superY);
		if (rectEmpty(rect))
			continue;
		for (uint32_t sy = (rect >> 8) & 0xFFu; sy <= (rect >> 24); sy++)
			for (uint32_t sx = rect & 0xFFu; sx <= ((rect >> 16) & 0xFFu); sx++)
				atomicAdd(&sCount[sy * a.superX + sx], 1u);
	}
	__syncthreads();
	for (uint32_t s = threadIdx.x; s < nSuper; s += SRPD_BIN_THREADS)
		a.chunkCounts[(size_t) blockIdx.x * nSuper + s] = sCount[s];
}

__global__ void __launch_bounds__(1024)
srpdBinScanKernel(const __grid_constant__ SrpdBinArgs a)
{
	/* single CTA; thread s owns supertile column s (coalesced across s) */
	__shared__ uint32_t sTotals[1024];
	__shared__ uint32_t sCarry;
	const uint32_t nSuper = a.superX * a.superY;
	const uint32_t nStored = a.frameCounts[1];
	const uint32_t nChunks = (nStored + SRPD_BIN_CHUNK - 1) / SRPD_BIN_CHUNK;
	if (threadIdx.x == 0)
		sCarry = 0;
	__syncthreads();
	for (uint32_t s0 = 0; s0 < nSuper; s0 += 1024)
	{
		const uint32_t s = s0 + threadIdx.x;
		uint32_t run = 0;
		if (s < nSuper)
			for (uint32_t c = 0; c < nChunks; c++)
			{
				const size_t at = (size_t) c * nSuper + s;
				const uint32_t n = a.chunkCounts[at];
				a.chunkCounts[at] = run;
				run += n;
			}
		sTotals[threadIdx.x] = run;
		__syncthreads();
		/* exclusive scan of the 1024 totals (Hillis-Steele in shared memory) */
		uint32_t v = run;
		for (uint32_t o = 1; o < 1024; o <<= 1)
		{
			const uint32_t add = threadIdx.x >= o ? sTotals[threadIdx.x - o] : 0;
			__syncthreads();
			v += add;
			sTotals[threadIdx.x] = v;
			__syncthreads();
		}
		const uint32_t carry = sCarry;
		const uint32_t excl = carry + v - run;
		if (s < nSuper)
			a.superOffsets[s] = excl < a.listCapacity ? excl : a.listCapacity;
		__syncthreads();
		if (threadIdx.x == 1023)
			sCarry = carry + v;
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		const uint32_t total = sCarry;
		a.superOffsets[nSuper] = total < a.listCapacity ? total : a.listCapacity;
		if (total > a.listCapacity)
		{
			atomicAdd(&a.stats->overflow, 1ull);
			atomicExch(a.abortFlag, 1u);
		}
	}
}

__global__ void __launch_bounds__(32)
srpdBinFillKernel(const __grid_constant__ SrpdBinArgs a)
{
	extern __shared__ uint32_t sCursor[];          /* [nSuper] write cursors of this chunk, then [32] lane rects */
	const uint32_t nSuper = a.superX * a.superY;
	uint32_t* sRect = sCursor + nSuper;
	const uint32_t nStored = a.frameCounts[1];
	const uint32_t first = blockIdx.x * SRPD_BIN_CHUNK;
	if (first >= nStored)
		return;
	const int lane = threadIdx.x;
	for (uint32_t s = lane; s < nSuper; s += 32)
		sCursor[s] = a.superOffsets[s] + a.chunkCounts[(size_t) blockIdx.x * nSuper + s];
	__syncwarp();

	for (uint32_t o = 0; o < SRPD_BIN_CHUNK; o += 32)
	{
		const uint32_t r = first + o + lane;
		if (first + o >= nStored)
			break;
		uint32_t rect = 0x00000101u;
		if (r < nStored)
			rect = superRect(a.bboxes[r], a.superX, a.superY);
		const bool empty = rectEmpty(rect);
		const uint32_t anyMulti = __ballot_sync(0xFFFFFFFFu, !empty && !rectSingle(rect));
		if (anyMulti == 0)
		{
			/* every record of this step covers exactly one supertile: rank among the lanes
			 * that picked the same supertile = position in that supertile's list */
			const uint32_t s = empty ? 0xFFFFFFFFu : ((rect >> 8) & 0xFFu) * a.superX + (rect & 0xFFu);
			const uint32_t peers = __match_any_sync(0xFFFFFFFFu, s);
			if (!empty)
			{
				const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
				const uint32_t pos = sCursor[s] + rank;
				if (pos < a.listCapacity)
					a.listIds[pos] = r;
			}
			__syncwarp();
			if (!empty && (peers >> lane) == 1u)     /* highest lane of the group advances the cursor */
				sCursor[s] += __popc(peers);
			__syncwarp();
		}
		else
		{
			/* general step: a record may cover a rectangle of supertiles */
			sRect[lane] = rect;
			__syncwarp();
			if (!empty)
				for (uint32_t sy = (rect >> 8) & 0xFFu; sy <= (rect >> 24); sy++)
					for (uint32_t sx = rect & 0xFFu; sx <= ((rect >> 16) & 0xFFu); sx++)
					{
						uint32_t rank = 0;
						for (int l = 0; l < lane; l++)
							rank += rectContains(sRect[l], sx, sy);
						const uint32_t pos = sCursor[sy * a.superX + sx] + rank;
						if (pos < a.listCapacity)
							a.listIds[pos] = r;
					}
			__syncwarp();
			if (!empty)
				for (uint32_t sy = (rect >> 8) & 0xFFu; sy <= (rect >> 24); sy++)
					for (uint32_t sx = rect & 0xFFu; sx <= ((rect >> 16) & 0xFFu); sx++)
					{
						uint32_t before = 0;
						bool last = true;
						for (int l = 0; l < 32; l++)
						{
							const bool c = rectContains(sRect[l], sx, sy);
							if (l < lane) before += c;
							if (l > lane && c) last = false;
						}
						if (last)
							sCursor[sy * a.superX + sx] += before + 1;
					}
			__syncwarp();
		}
	}
}

static int gBinLaunches = 0;
int srpdBinLaunchCount(void) { return gBinLaunches; }

void srpdLaunchBin(const SrpdBinArgs& a, cudaStream_t stream)
{
	const uint32_t nSuper = a.superX * a.superY;
	srpdBinCountKernel<<<a.nChunksMax, SRPD_BIN_THREADS, nSuper * sizeof(uint32_t), stream>>>(a);
	srpdBinScanKernel<<<1, 1024, 0, stream>>>(a);
	srpdBinFillKernel<<<a.nChunksMax, 32, (nSuper + 32) * sizeof(uint32_t), stream>>>(a);
	gBinLaunches += 3;
}
