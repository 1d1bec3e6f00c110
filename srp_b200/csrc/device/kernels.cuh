/* srp-b200 internal -- kernel argument blocks and launch geometry shared by the CUDA
 * translation units (geom.cu, bin.cu, raster.cu, runtime.cu). */
#pragma once
#include <cuda.h>              /* CUtensorMap (type only: the driver entry point is fetched at run time) */
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "core.cuh"

/* User program table (defined by SRP_B200_DEFINE_PROGRAM_TABLE in the executable's
 * shader translation unit, resolved at device-link time). */
extern "C" __device__ void srpB200DeviceVS(int programId, SRPVertexShaderIn* in, SRPVertexShaderOut* out);
extern "C" __device__ void srpB200DeviceFS(int programId, SRPFragmentShaderIn* in, SRPFragmentShaderOut* out);

/* ---- launch geometry -------------------------------------------------------------- */
/* Tiles.  Geometry, occupancy bitmap, coarse binning and the strips' row ranges work on 32x16
 * tiles (SRPD_TILE_W x SRPD_TILE_H).  The tile kernel rasterises WARP TILES of 32x8 pixels --
 * half a tile: one 128-byte colour row per tile row, 8 rows -- each owned by one warp that
 * shares nothing with the other warps of its CTA (no CTA barrier anywhere).  128-thread CTAs,
 * eight per SM (32 warps, <= 64 registers, ~6.6 KB of shared memory per warp: the warp tile's
 * pixels, a 128-entry fragment queue, the step's triangle data and a ring of list entries). */
#ifndef SRPD_TILE_CTAS_PER_SM
#define SRPD_TILE_CTAS_PER_SM 8
#endif
constexpr int SRPD_TILE_W = 32;
constexpr int SRPD_TILE_H = 16;
constexpr int SRPD_WT_W = 32;            /* warp tile */
constexpr int SRPD_WT_H = 8;
constexpr int SRPD_WT_PIXELS = SRPD_WT_W * SRPD_WT_H;
constexpr int SRPD_TILE_THREADS = 128;
constexpr int SRPD_TILE_WARPS = SRPD_TILE_THREADS / 32;
static_assert(SRPD_WT_W == SRPD_TILE_W && SRPD_TILE_H == 2 * SRPD_WT_H, "a tile is two warp tiles, one above the other");
/* coarse bins ("supertiles") are 2^k x 2^k tiles with k = superShift chosen per draw from the
 * expected record density: 8x8 tiles (256x128 px) for ordinary meshes down to 2x2 for
 * millions of sub-pixel primitives, so that a tile never filters more than ~1-2 k candidates */
constexpr int SRPD_SUPER_SHIFT_MAX = 3;

/* geometry: one WARP per batch of SRPD_GEOM_PRIMS consecutive input primitives.  30 rather than
 * 32: a batch of 2n triangles of a regular grid or strip references n + 2 .. 2n + 2 distinct
 * vertices, so 30 keeps the distinct vertices of such a batch within one 32-lane pass of the
 * vertex shader (a 33rd vertex would cost a second, almost empty pass) */
#ifndef SRPD_GEOM_WARP_PRIMS
#define SRPD_GEOM_WARP_PRIMS 30
#endif
constexpr int SRPD_GEOM_PRIMS = SRPD_GEOM_WARP_PRIMS;
#ifndef SRPD_GEOM_WARPS_PER_CTA
#define SRPD_GEOM_WARPS_PER_CTA 4
#endif
constexpr int SRPD_GEOM_WARPS = SRPD_GEOM_WARPS_PER_CTA;
#ifndef SRPD_GEOM_CTAS_PER_SM
#define SRPD_GEOM_CTAS_PER_SM (32 / SRPD_GEOM_WARPS_PER_CTA)
#endif
constexpr int SRPD_SCAN_CHUNK = 256;     /* batches per CTA of the batch-order kernel (one per thread) */
/* draws of at most this many batches run as one launch of clipper warps (geom.cu) */
constexpr unsigned SRPD_GEOM_SMALL_DRAW_BATCHES = 512;
constexpr int SRPD_GEOM_THREADS = 32 * SRPD_GEOM_WARPS;
constexpr int SRPD_GEOM_MAX_VERTS = 3 * 32;
constexpr int SRPD_HASH_SLOTS = 128;     /* post-VS cache: open-addressing table in smem, load <= 3/4 */
constexpr int SRPD_HASH_SHIFT = 25;
constexpr uint32_t SRPD_HASH_EMPTY = 0xFFFFFFFFu;
constexpr int SRPD_CLIP_MAX_VERTS = 10;  /* a triangle against 6 planes has <= 9 vertices  */

/* A single-frame draw carries its uniform block inside the kernel argument block (constant
 * bank): the shaders' uniform reads -- matrices, light parameters, texture pointers, the same
 * for every thread -- then are constant-bank operands instead of generic loads.  Larger
 * uniforms, NULL uniforms and batches go through the per-frame device bindings. */
constexpr int SRPD_INLINE_UNIFORM_BYTES = 1024;

constexpr int SRPD_STATS_SLOTS = 1024;   /* SrpdStats[slots]: counters are spread to avoid same-address atomics */

constexpr int SRPD_BIN_THREADS = 256;
constexpr int SRPD_BIN_CHUNK = 2048;     /* records per coarse-binning CTA                  */
constexpr uint32_t SRPD_BIN_SMALL_RECORDS = 1u << 18;   /* up to here chunks are SRPD_BIN_CHUNK / 4 (bin.cu) */
constexpr uint32_t SRPD_BIN_SMEM_SUPERS = 1024;         /* more supertiles than this: one warp per chunk, cursors in global memory */
constexpr uint32_t SRPD_BIN_WARP_CHUNKS = 4096;         /* chunks of that path at most (chunks grow beyond 512 records instead) */

/* Per-draw zero-filled header in front of the scan state: word 0 coarse-list overflow flag, 1 abort flag,
 * 2 tile work counter, 3 records a frame needed (max over frames), 4 coarse-list entries needed,
 * 6 number of large triangles, 7 deferred batches, 8..11 tile work counters of the bands,
 * 12-13 checkpoint-table cursor (entries, 64-bit) */
constexpr int SRPD_DRAW_HEADER_BYTES = 64;
constexpr int SRPD_MAX_BANDS = 4;


struct SrpdGeomArgs
{
	SrpdDraw d;
	SrpdFrame frame0;                 /* used when frames == nullptr (single draw)       */
	const SrpdFrame* frames;          /* device array [nFrames]; nullptr: frame0 + uniformInline */
	alignas(16) unsigned char uniformInline[SRPD_INLINE_UNIFORM_BYTES];
	unsigned char* records;           /* [nFrames][recCapacity] records of recStride B   */
	uint2* bboxes;                    /* [nFrames][recCapacity] x0|y0<<16, x1|y1<<16 (half-open, pixels) */
	uint32_t recCapacity;
	uint32_t recStride;
	uint4* ordered;                   /* [nFrames][recCapacity] the id-ordered view: entry i = the i-th stored primitive
	                                     in primitive order: {box.x, box.y, record slot, id prefix of its batch} */
	uint32_t* idCarry;                /* [nFrames] ids handed out by the previous sub-draw of a split draw (read if d.chunkIndex) */
	uint32_t* idCarryOut;             /* [nFrames] written by every sub-draw (the other one of two arrays: the CTAs of srpdBatchOrderKernel run concurrently) */
	uint32_t* frameBump;              /* [nFrames], zeroed per draw: record slots handed out */
	uint32_t* deferCount;             /* header word 7, zeroed per draw: batches the main pass left to the clipper pass */
	uint32_t* deferList;              /* [nFrames * batchesPerFrame] their indices              */
	uint32_t deferred;                /* set by srpdLaunchGeom: this launch works off the list  */
	uint32_t smCount;
	uint4* batchInfo;                 /* [nFrames * batchesPerFrame] {first slot, ids, records, -} */
	uint32_t* abortFlag;              /* zeroed per draw; set when a scratch pool overflows: the
	                                     tile kernel then leaves the framebuffer untouched  */
	uint32_t* needed;                 /* [0] records needed per frame (max), [1] coarse-list entries needed */
	uint32_t batchesPerFrame;
	uint32_t chunksPerFrame;          /* ceil(batchesPerFrame / SRPD_SCAN_CHUNK) */
	uint2* chunkSums;                 /* [nFrames][chunksPerFrame] {ids, records}, zeroed per draw */
	uint32_t* frameCounts;            /* [nFrames][2]: ids emitted, records stored       */
	uint32_t* occupancy;              /* [nFrames][occWordsPerFrame], zeroed per draw: tiles touched by a stored record */
	uint32_t occWordsPerFrame;
	uint32_t tilesX, tilesY;
	/* large triangles: barycentric checkpoints (checkpoint.cu) */
	unsigned long long* ckptCursor;   /* header words 12-13 (64-bit: cannot wrap): entries handed out */
	uint32_t* largeCount;             /* header word 6                                   */
	uint2* largeList;                 /* [largeCapacity] {frame, record}                 */
	uint32_t largeCapacity;
	uint32_t ckptCapacity;            /* entries (3 floats each) in the checkpoint table */
	SrpdStats* stats;
};

/* Barycentric checkpoints of large triangles: for every pixel row of the bounding box and
 * every tile column it spans, lambda at the first covered pixel of that column.  Entry
 * (row, col) of a triangle lives at ckptTable[3 * (offset + row * cols + col)]. */
constexpr int SRPD_LARGE_EXTENT = 96;    /* a triangle is "large" when its box exceeds this in x or y */
struct SrpdCkptArgs
{
	const unsigned char* records;
	uint32_t recCapacity, recStride;
	const uint32_t* largeCount;
	const uint2* largeList;
	uint32_t largeCapacity;
	float* ckptTable;
};
void srpdLaunchCheckpoints(const SrpdCkptArgs& a, cudaStream_t stream);

struct SrpdBinArgs
{
	const uint4* ordered;             /* the id-ordered view (.x, .y = bounding box) */
	const uint32_t* frameCounts;      /* [0][1] = number of stored records               */
	uint32_t nChunksMax;              /* chunks at full record capacity                  */
	uint32_t smCount;
	uint32_t superX, superY;          /* supertile grid                                  */
	uint32_t superShift;              /* supertile = (1 << superShift)^2 tiles           */
	uint32_t* superTotals;            /* [nSuper] entries per supertile (scan pass 1 -> 2)*/
	uint32_t* chunkCounts;            /* [nChunksMax][nSuper]                            */
	uint32_t* superOffsets;           /* [nSuper + 1]                                    */
	uint4* listEntries;               /* [listCapacity] copies of the ordered view's entries {box, slot, id prefix}, id order per supertile */
	uint32_t listCapacity;
	uint32_t* listOverflow;           /* zeroed per draw; set when the coarse lists do not fit listCapacity: the tile
	                                     kernel then ignores them and lets every tile scan all records (slow, exact) */
	uint32_t* scanTicket;             /* header word 5, zeroed per draw: CTAs of the column scan that have finished */
	uint32_t* needed;
	uint32_t* hostNotes;              /* pinned, mapped: [1] = coarse-list entries a draw needed (the host raises the pool) */
	SrpdStats* stats;
};

struct SrpdTileArgs
{
	SrpdDraw d;
	SrpdFrame frame0;
	const SrpdFrame* frames;          /* nullptr: frame0 + uniformInline                 */
	alignas(16) unsigned char uniformInline[SRPD_INLINE_UNIFORM_BYTES];
	const unsigned char* records;
	const uint4* ordered;             /* the id-ordered view: {box.x, box.y, record slot, id prefix} */
	uint32_t recCapacity;
	uint32_t recStride;
	const uint32_t* frameCounts;
	const uint32_t* superOffsets;     /* nullptr: direct path, every tile scans all records */
	const uint4* listEntries;
	const uint32_t* listOverflow;     /* != 0: the coarse lists are incomplete, scan all records instead */
	uint32_t superX;
	uint32_t superShift;
	uint32_t tilesX, tilesY;
	const uint32_t* abortFlag;
	const uint32_t* occupancy;        /* [nFrames][occWordsPerFrame] bit per tile: some record's box touches it */
	uint32_t occWordsPerFrame;
	uint32_t* workCounter;            /* zeroed per draw: next work item of the persistent tile kernel */
	uint32_t tilesPerItem;            /* consecutive tiles of one frame per work item */
	uint64_t tilesXInv;               /* floor(2^40 / tilesX) + 1, filled by srpdLaunchTiles */
	const float* ckptTable;           /* barycentric checkpoints of large triangles     */
	uint32_t smCount;
	SrpdStats* stats;
	/* TMA tile stores of the write-back (single-frame draws): tensor maps of frame0's planes with
	 * a 32x8 box; tmaPlanes = planes that have one (bit0 colour, bit1 depth, bit2 stencil; 0 = none) */
	uint32_t tmaPlanes;
	alignas(64) CUtensorMap tmColor;
	alignas(64) CUtensorMap tmDepth;
	alignas(64) CUtensorMap tmStencil;
};

/* ---- programmatic dependent launch ---------------------------------------------------
 * A draw is a chain of ~10 dependent kernels, several of them only 3-10 us long, so the gaps
 * between them (grid drain + launch latency of the next one) are a visible share of a frame.
 * Every kernel is launched with the programmatic-stream-serialisation attribute and starts with
 * srpdGridDependencyEnter(): `griddepcontrol.wait` blocks until the preceding kernel of the
 * stream has completed and its memory is visible (so the dependency itself is unchanged --
 * nothing is read or written before it), then `griddepcontrol.launch_dependents` lets the NEXT
 * kernel's CTAs be scheduled as soon as all of this kernel's CTAs have got that far: they become
 * resident on SMs as these drain and sit in their own wait, so the next kernel starts the
 * moment this one ends.  Without the launch attribute both instructions are no-ops.
 * SRP_B200_PDL=0/1 switches the attribute (runtime.cu). */
__device__ __forceinline__ void srpdGridDependencyEnter()
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#ifndef SRPD_PDL_DEFAULT
#define SRPD_PDL_DEFAULT 1
#endif
bool srpdPdlEnabled(void);

template <typename Arg>
inline cudaError_t srpdLaunchKernel(void (*kernel)(Arg), unsigned grid, unsigned block, size_t smemBytes, cudaStream_t stream, const Arg& a)
{
	cudaLaunchConfig_t cfg;
	memset(&cfg, 0, sizeof cfg);
	cfg.gridDim = dim3(grid, 1, 1);
	cfg.blockDim = dim3(block, 1, 1);
	cfg.dynamicSmemBytes = smemBytes;
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = srpdPdlEnabled() ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kernel, a);
}

/* launchers (defined next to their kernels) */
int srpdLaunchGeom(const SrpdGeomArgs& a, cudaStream_t stream);
void srpdLaunchBin(const SrpdBinArgs& a, cudaStream_t stream, cudaEvent_t joinBeforeFill);
void srpdLaunchTiles(const SrpdTileArgs& a, cudaStream_t stream);
void srpdLaunchClear(uint32_t* color, float* depth, size_t nPixels, cudaStream_t stream);
int srpdGeomLaunchCount(void);
int srpdBinLaunchCount(void);
