/* srp-b200 -- CUDA runtime layer behind the thin C ABI of srpcu.h.
 *
 * Owns what the reference's src/memory (per-draw bump arena, src/memory/arena.c:69-116)
 * owned, re-designed for the device: one stream, grow-only scratch pools in HBM that
 * every draw re-uses (primitive records, bounding boxes, scan state, coarse-bin lists,
 * uniform ring), and the submission of the kernels of one draw:
 *
 *   memset(scan state)  ->  geometry  ->  [count, scan, fill coarse bins]  ->  tiles
 *
 * Nothing here falls back to the CPU: without a usable sm_100 device every entry point
 * fails and says why. */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "kernels.cuh"
#include "srpcu.h"

namespace {

struct Pool
{
	void* ptr = nullptr;
	size_t bytes = 0;
};

struct Runtime
{
	bool ready = false;
	bool failed = false;
	int device = -1;
	cudaStream_t stream = nullptr;
	cudaStream_t copyStream = nullptr;     /* band-wise downloads overlapping the tile kernel */
	cudaEvent_t bandEvents[SRPD_MAX_BANDS] = {};
	cudaEvent_t copyDone = nullptr;
	cudaEvent_t submitted = nullptr;       /* async downloads: "everything enqueued so far" on the submission stream */
	cudaStream_t auxStream = nullptr;      /* work off the critical path of a draw (checkpoint pre-pass) */
	cudaStream_t uploadStream = nullptr;   /* srp*BufferCopyData: ordered behind the last draw that reads the buffer only */
	cudaEvent_t uploadDone = nullptr;
	cudaEvent_t laneFence = nullptr;       /* cross-lane ordering (srpcuOrderBehindLane, uploads) */
	/* pinned staging ring for per-draw host data (uniform blocks, frame bindings): the caller's
	 * memory has been read when the draw call returns, whatever kind of memory it is */
	unsigned char* ring = nullptr;
	size_t ringBytes = 0, ringHead = 0;
	cudaEvent_t ringWrapped = nullptr;     /* recorded on the submission stream when the ring wraps */
	bool ringWrapPending = false;
	uint32_t* hostNotes = nullptr;         /* pinned + mapped, written by kernels: [0] draws that hit a pool limit (a bug
	                                          guard), [1] coarse-list entries a draw needed */
	uint32_t* hostNotesDev = nullptr;      /* device alias */
	unsigned long long guardSeen = 0;
	size_t poolBudget = 0;                 /* bytes the per-record scratch arrays of one sub-draw may take */
	cudaEvent_t geomDone = nullptr, ckptDone = nullptr;
	SrpcuMirror mirror = { nullptr, nullptr, nullptr };
	int* mirrorDone = nullptr;
	bool copyPending = false;
	std::string lastError;
	Pool records, bboxes, scan, frameCounts, chunkCounts, superOffsets, superTotals, listIds, uniforms, frames;
	Pool ckptTable, largeList, ordered, batchInfo, deferList, idCarry;
	int smCount = 148;
	uint64_t listFloor = 0;                /* minimum coarse-list capacity, raised when a draw reported it needed more */
	SrpdStats* stats = nullptr;            /* device, SRPD_STATS_SLOTS slots */
	SrpdStats* statsHost = nullptr;        /* pinned */
	unsigned long long launches = 0, h2d = 0, d2h = 0;
	int forceBinning = -1;                 /* SRP_B200_BINNING=0/1 overrides the heuristic */
	uint32_t binThreshold = 4096;
	bool ckptAside = true;               /* checkpoint pre-pass beside the binning kernels (auxiliary stream) */
	/* optional per-stage timing: 4 events per draw (start, after geometry, after binning,
	 * after tiles) on the submission stream, summed when collected */
	bool profile = false;
	std::vector<cudaEvent_t> freeEvents;
	std::vector<cudaEvent_t> pendingEvents;   /* groups of 4 */
	double stageMs[3] = { 0, 0, 0 };
	unsigned long long stageDraws = 0;
};

/* Lanes: independent copies of the whole submission state -- stream, scratch pools, staging ring,
 * counters -- so that consecutive frames enqueued on different lanes overlap on the device (the
 * low-occupancy front-end kernels of one frame run under the tile kernel of another).  Everything
 * below addresses the current lane through `g`; srpcuSetLane selects it.  Lanes share the device
 * and nothing else; the few cross-lane orderings (uploads, a framebuffer changing lanes) are
 * explicit events. */
Runtime gLanes[SRPCU_MAX_LANES];
int gLane = 0;
#define g gLanes[gLane]
int gRequestedDevice = -1;
bool gProfile = false;

struct LaneScope
{
	int keep;
	explicit LaneScope(int lane) : keep(gLane) { gLane = lane; }
	~LaneScope() { gLane = keep; }
};


bool fail(const char* what, cudaError_t e)
{
	char buf[512];
	snprintf(buf, sizeof buf, "srp-b200: %s failed: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
	g.lastError = buf;
	return false;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(#call, e_); return 1; } } while (0)

bool grow(Pool& p, size_t bytes)
{
	if (bytes <= p.bytes)
		return true;
	if (p.ptr)
	{
		cudaStreamSynchronize(g.stream);
		cudaFree(p.ptr);
		p.ptr = nullptr; p.bytes = 0;
	}
	size_t want = bytes + bytes / 4;
	want = (want + 255) & ~(size_t) 255;
	cudaError_t e = cudaMalloc(&p.ptr, want);
	if (e != cudaSuccess)
	{
		e = cudaMalloc(&p.ptr, bytes);
		want = bytes;
	}
	if (e != cudaSuccess)
		return fail("cudaMalloc(scratch pool)", e);
	p.bytes = want;
	return true;
}

cudaEvent_t takeEvent()
{
	cudaEvent_t e = nullptr;
	if (!g.freeEvents.empty()) { e = g.freeEvents.back(); g.freeEvents.pop_back(); }
	else cudaEventCreate(&e);
	return e;
}
void mark()
{
	if (!g.profile) return;
	cudaEvent_t e = takeEvent();
	cudaEventRecord(e, g.stream);
	g.pendingEvents.push_back(e);
}

} // namespace

/* programmatic dependent launch of the draw's kernel chain (kernels.cuh); SRP_B200_PDL=0 turns
 * the launch attribute off (the kernels' griddepcontrol instructions are then no-ops) */
bool srpdPdlEnabled(void)
{
	static int enabled = -1;
	if (enabled < 0)
	{
		const char* e = getenv("SRP_B200_PDL");
		enabled = e ? (atoi(e) != 0) : SRPD_PDL_DEFAULT;
	}
	return enabled != 0;
}

namespace {

int envInt(const char* name, int fallback)
{
	const char* v = getenv(name);
	return (v && *v) ? atoi(v) : fallback;
}

} // namespace

/* ---- TMA tensor maps of framebuffer planes (tile write-back, raster.cu) -------------------------
 * A plane is a row-major 2-D tensor [height][width]; the box is one 32x8 warp tile.  Colour and
 * depth use the 128-byte swizzle (the layout the tile kernel keeps them in in shared memory),
 * stencil none.  The driver's encoder is fetched through the runtime (no link against libcuda);
 * encoded maps are cached per (plane, size). */
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct CachedMap { const void* plane; int width, height, kind; CUtensorMap map; };
std::vector<CachedMap> gTensorMaps;

EncodeTiledFn encodeTiled()
{
	static EncodeTiledFn fn = nullptr;
	static bool tried = false;
	if (!tried)
	{
		tried = true;
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
			fn = (EncodeTiledFn) p;
		else
			cudaGetLastError();
	}
	return fn;
}

/* kind: 0 colour (u32), 1 depth (f32), 2 stencil (u8).  false: no map (pitch / alignment / driver) */
bool planeTensorMap(const void* plane, int width, int height, int kind, CUtensorMap* out)
{
	const size_t elem = kind == 2 ? 1 : 4;
	if (getenv("SRP_B200_NO_TMA") || !plane || ((uintptr_t) plane & 15u) || ((size_t) width * elem) % 16 != 0)
		return false;
	for (const CachedMap& c : gTensorMaps)
		if (c.plane == plane && c.width == width && c.height == height && c.kind == kind)
		{
			*out = c.map;
			return true;
		}
	EncodeTiledFn enc = encodeTiled();
	if (!enc)
		return false;
	const cuuint64_t dims[2] = { (cuuint64_t) width, (cuuint64_t) height };
	const cuuint64_t strides[1] = { (cuuint64_t) width * elem };
	const cuuint32_t box[2] = { (cuuint32_t) SRPD_WT_W, (cuuint32_t) SRPD_WT_H };
	const cuuint32_t estr[2] = { 1, 1 };
	CachedMap c;
	c.plane = plane; c.width = width; c.height = height; c.kind = kind;
	const CUresult r = enc(&c.map, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8,
	                       2, const_cast<void*>(plane), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                       kind == 2 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
	                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS)
		return false;
	if (gTensorMaps.size() >= 256)
		gTensorMaps.clear();
	gTensorMaps.push_back(c);
	*out = c.map;
	return true;
}

} // namespace

extern "C" {

void srpcuSetDevice(int device) { gRequestedDevice = device; }

const char* srpcuLastError(void) { return g.lastError.c_str(); }
void* srpcuStream(void) { return srpcuInit() == 0 ? (void*) g.stream : nullptr; }
int srpcuTileWidth(void) { return SRPD_TILE_W; }
int srpcuTileHeight(void) { return SRPD_TILE_H; }

const char* srpcuVersion(void)
{
	static char buf[128];
	snprintf(buf, sizeof buf, "srp-b200 sm_100a tile %dx%d warp-tile %dx%d geom-batch %d line-seg %d",
	         SRPD_TILE_W, SRPD_TILE_H, SRPD_WT_W, SRPD_WT_H, SRPD_GEOM_PRIMS, SRPD_LINE_SEG);
	return buf;
}

int srpcuInit(void)
{
	if (g.ready)
		return 0;
	if (g.failed)
		return 1;
	g.failed = true;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
	{
		g.lastError = std::string("srp-b200: no CUDA device available (") + cudaGetErrorString(e)
			+ "); this library has no CPU path";
		return 1;
	}
	int dev = gRequestedDevice;
	if (dev < 0) dev = envInt("SRP_B200_DEVICE", -1);
	if (dev < 0) dev = envInt("LOCAL_RANK", 0);
	if (dev >= count) dev = dev % count;
	CU(cudaSetDevice(dev));
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, dev));
	if (prop.major != 10)
	{
		char buf[256];
		snprintf(buf, sizeof buf, "srp-b200: device %d (%s, sm_%d%d) is not an sm_100 part; the kernels are built for sm_100a only",
		         dev, prop.name, prop.major, prop.minor);
		g.lastError = buf;
		return 1;
	}
	g.device = dev;
	g.smCount = prop.multiProcessorCount;
	CU(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithFlags(&g.copyStream, cudaStreamNonBlocking));
	for (int i = 0; i < SRPD_MAX_BANDS; i++)
		CU(cudaEventCreateWithFlags(&g.bandEvents[i], cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&g.copyDone, cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&g.submitted, cudaEventDisableTiming));
	CU(cudaStreamCreateWithFlags(&g.auxStream, cudaStreamNonBlocking));
	CU(cudaEventCreateWithFlags(&g.geomDone, cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&g.ckptDone, cudaEventDisableTiming));
	CU(cudaStreamCreateWithFlags(&g.uploadStream, cudaStreamNonBlocking));
	CU(cudaEventCreateWithFlags(&g.uploadDone, cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&g.laneFence, cudaEventDisableTiming));
	CU(cudaEventCreateWithFlags(&g.ringWrapped, cudaEventDisableTiming));
	g.ringBytes = (size_t) 8 << 20;
	CU(cudaMallocHost((void**) &g.ring, g.ringBytes));
	CU(cudaHostAlloc((void**) &g.hostNotes, 64, cudaHostAllocMapped));
	memset(g.hostNotes, 0, 64);
	CU(cudaHostGetDevicePointer((void**) &g.hostNotesDev, g.hostNotes, 0));
	{
		/* scratch budget of one sub-draw: a sixth of the device memory unless told otherwise (the
		 * worst case of a million lines on a 4K framebuffer, 243 segment records each, is 29 GB) */
		const int mb = envInt("SRP_B200_POOL_BUDGET_MB", 0);
		g.poolBudget = mb > 0 ? (size_t) mb << 20 : prop.totalGlobalMem / 6;
	}
	CU(cudaMalloc(&g.stats, sizeof(SrpdStats) * SRPD_STATS_SLOTS));
	CU(cudaMemset(g.stats, 0, sizeof(SrpdStats) * SRPD_STATS_SLOTS));
	CU(cudaMallocHost(&g.statsHost, sizeof(SrpdStats) * SRPD_STATS_SLOTS));
	g.forceBinning = envInt("SRP_B200_BINNING", -1);
	g.binThreshold = (uint32_t) envInt("SRP_B200_BIN_THRESHOLD", 4096);
	g.ckptAside = envInt("SRP_B200_CKPT_ASIDE", 1) != 0;
	g.profile = gProfile;
	/* buffer uploads issued on other lanes so far: this lane's draws come after them too */
	for (int l = 0; l < SRPCU_MAX_LANES; l++)
		if (l != gLane && gLanes[l].ready)
			CU(cudaStreamWaitEvent(g.stream, gLanes[l].uploadDone, 0));
	g.failed = false;
	g.ready = true;
	return 0;
}

int srpcuSetLane(int lane)
{
	if (lane < 0 || lane >= SRPCU_MAX_LANES)
		return 1;
	gLane = lane;
	return 0;
}
int srpcuLane(void) { return gLane; }
int srpcuLaneCount(void) { return SRPCU_MAX_LANES; }

/* everything enqueued on `other` so far happens before what the current lane enqueues from now on */
int srpcuOrderBehindLane(int other)
{
	if (other < 0 || other >= SRPCU_MAX_LANES || other == gLane || !gLanes[other].ready)
		return 0;
	if (srpcuInit()) return 1;
	Runtime& o = gLanes[other];
	CU(cudaEventRecord(o.laneFence, o.stream));
	CU(cudaStreamWaitEvent(g.stream, o.laneFence, 0));
	if (o.copyPending)
	{
		/* band-wise or asynchronous downloads still reading the planes */
		CU(cudaEventRecord(o.laneFence, o.copyStream));
		CU(cudaStreamWaitEvent(g.stream, o.laneFence, 0));
	}
	return 0;
}

void* srpcuMalloc(size_t bytes)
{
	if (srpcuInit()) return nullptr;
	void* p = nullptr;
	if (bytes == 0) bytes = 16;
	cudaError_t e = cudaMalloc(&p, bytes);
	if (e != cudaSuccess) { fail("cudaMalloc", e); return nullptr; }
	cudaMemsetAsync(p, 0, bytes, g.stream);
	cudaStreamSynchronize(g.stream);      /* uploads run on their own stream */
	return p;
}
void srpcuFree(void* p)
{
	if (p && g.ready) { cudaStreamSynchronize(g.stream); cudaFree(p); }
}
void* srpcuMallocHost(size_t bytes)
{
	if (srpcuInit()) return nullptr;
	void* p = nullptr;
	if (bytes == 0) bytes = 16;
	cudaError_t e = cudaMallocHost(&p, bytes);
	if (e != cudaSuccess) { fail("cudaMallocHost", e); return nullptr; }
	memset(p, 0, bytes);
	return p;
}
void srpcuFreeHost(void* p)
{
	if (p && g.ready) { cudaStreamSynchronize(g.stream); cudaFreeHost(p); }
}
void* srpcuMallocManaged(size_t bytes)
{
	if (srpcuInit()) return nullptr;
	void* p = nullptr;
	if (bytes == 0) bytes = 16;
	cudaError_t e = cudaMallocManaged(&p, bytes);
	if (e != cudaSuccess) { fail("cudaMallocManaged", e); return nullptr; }
	return p;
}
void srpcuFreeManaged(void* p)
{
	if (p && g.ready) { cudaStreamSynchronize(g.stream); cudaFree(p); }
}
void srpcuPrefetchToDevice(void* p, size_t bytes)
{
	if (!g.ready || !p) return;
	cudaMemAdvise(p, bytes, cudaMemAdviseSetReadMostly, g.device);
	cudaMemPrefetchAsync(p, bytes, g.device, g.stream);
}

/* Pinned staging for small per-draw host data: copies `bytes` from the caller into the ring and
 * returns the staged address (nullptr if it cannot fit).  Space is reused after a wrap, once the
 * submission stream has passed the event recorded at the wrap. */
static unsigned char* ringReserve(size_t bytes)
{
	const size_t need = (bytes + 255) & ~(size_t) 255;
	if (need > g.ringBytes / 2)
		return nullptr;
	if (g.ringHead + need > g.ringBytes)
	{
		/* everything staged so far has been consumed once the stream gets here */
		cudaEventRecord(g.ringWrapped, g.stream);
		cudaEventSynchronize(g.ringWrapped);
		g.ringHead = 0;
	}
	unsigned char* at = g.ring + g.ringHead;
	g.ringHead += need;
	return at;
}
static unsigned char* stageHostData(const void* src, size_t bytes)
{
	unsigned char* at = ringReserve(bytes);
	if (at)
		memcpy(at, src, bytes);
	return at;
}

/* srp*BufferCopyData: memcpy semantics -- when the call returns the caller may reuse `src`.
 * The copy runs on the upload stream, ordered behind `lastUse` (the last draw that reads the
 * destination; may be null) instead of behind everything the submission stream still has
 * queued, and later draws are ordered behind the copy.  Pageable sources are staged by the CUDA
 * runtime before the call returns; pinned, registered and device sources are waited for. */
int srpcuUpload(void* dst, const void* src, size_t bytes, void* lastUse)
{
	if (srpcuInit()) return 1;
	if (bytes == 0) return 0;
	if (lastUse)
		CU(cudaStreamWaitEvent(g.uploadStream, (cudaEvent_t) lastUse, 0));
	/* `lastUse` was recorded on one lane; draws of the others may read the buffer as well: behind
	 * everything they have enqueued so far */
	for (int l = 0; l < SRPCU_MAX_LANES; l++)
		if (l != gLane && gLanes[l].ready)
		{
			CU(cudaEventRecord(gLanes[l].laneFence, gLanes[l].stream));
			CU(cudaStreamWaitEvent(g.uploadStream, gLanes[l].laneFence, 0));
		}
	CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, g.uploadStream));
	cudaPointerAttributes attr;
	bool pageable = false;
	if (cudaPointerGetAttributes(&attr, src) == cudaSuccess)
		pageable = attr.type == cudaMemoryTypeUnregistered;
	else
		cudaGetLastError();
	if (!pageable)
		CU(cudaStreamSynchronize(g.uploadStream));
	CU(cudaEventRecord(g.uploadDone, g.uploadStream));
	for (int l = 0; l < SRPCU_MAX_LANES; l++)
		if (gLanes[l].ready)
			CU(cudaStreamWaitEvent(gLanes[l].stream, g.uploadDone, 0));
	g.h2d += bytes;
	return 0;
}
/* plain stream-ordered copy on the submission stream (framebuffer uploads; synchronised by the caller) */
int srpcuUploadInStream(void* dst, const void* src, size_t bytes)
{
	if (srpcuInit()) return 1;
	if (bytes == 0) return 0;
	CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, g.stream));
	g.h2d += bytes;
	return 0;
}
int srpcuRecordEvent(void* event)
{
	if (!g.ready || !event) return 0;
	CU(cudaEventRecord((cudaEvent_t) event, g.stream));
	return 0;
}
int srpcuDownload(void* dstHost, const void* srcDevice, size_t bytes)
{
	if (srpcuInit()) return 1;
	if (bytes == 0) return 0;
	CU(cudaMemcpyAsync(dstHost, srcDevice, bytes, cudaMemcpyDeviceToHost, g.stream));
	g.d2h += bytes;
	return 0;
}
void srpcuSetMirrorForNextDraw(const SrpcuMirror* mirror, int* done)
{
	g.mirror = mirror ? *mirror : SrpcuMirror{ nullptr, nullptr, nullptr };
	g.mirrorDone = done;
	if (done) *done = 0;
}

void* srpcuNewEvent(void)
{
	if (srpcuInit()) return nullptr;
	cudaEvent_t e = nullptr;
	cudaError_t err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
	if (err != cudaSuccess) { fail("cudaEventCreate", err); return nullptr; }
	return (void*) e;
}
void srpcuFreeEvent(void* event)
{
	if (event && g.ready) cudaEventDestroy((cudaEvent_t) event);
}
int srpcuDownloadPlanesAsync(const SrpcuMirror* host, const void* dColor, const void* dDepth, const void* dStencil,
                             size_t nPixels, void* doneEvent)
{
	if (srpcuInit()) return 1;
	CU(cudaEventRecord(g.submitted, g.stream));
	CU(cudaStreamWaitEvent(g.copyStream, g.submitted, 0));
	if (host->color)
	{
		CU(cudaMemcpyAsync(host->color, dColor, nPixels * 4, cudaMemcpyDeviceToHost, g.copyStream));
		g.d2h += nPixels * 4;
	}
	if (host->depth)
	{
		CU(cudaMemcpyAsync(host->depth, dDepth, nPixels * 4, cudaMemcpyDeviceToHost, g.copyStream));
		g.d2h += nPixels * 4;
	}
	if (host->stencil)
	{
		CU(cudaMemcpyAsync(host->stencil, dStencil, nPixels, cudaMemcpyDeviceToHost, g.copyStream));
		g.d2h += nPixels;
	}
	CU(cudaEventRecord((cudaEvent_t) doneEvent, g.copyStream));
	g.copyPending = true;
	return 0;
}
int srpcuHostWaitEvent(void* event)
{
	if (!g.ready || !event) return 0;
	CU(cudaEventSynchronize((cudaEvent_t) event));
	return 0;
}
int srpcuStreamWaitEvent(void* event)
{
	if (!g.ready || !event) return 0;
	CU(cudaStreamWaitEvent(g.stream, (cudaEvent_t) event, 0));
	return 0;
}

static int synchronizeLane(void)
{
	if (!g.ready) return 0;
	if (g.copyPending)
	{
		CU(cudaStreamSynchronize(g.copyStream));
		g.copyPending = false;
	}
	CU(cudaStreamSynchronize(g.stream));
	CU(cudaGetLastError());
	if (g.hostNotes && g.hostNotes[2])
	{
		char buf[160];
		snprintf(buf, sizeof buf, "srp-b200: srpB200StreamWait gave up waiting for a flag to reach %u (a peer never signalled)", g.hostNotes[2]);
		g.hostNotes[2] = 0;
		g.lastError = buf;
		return 1;
	}
	return 0;
}
/* waits for everything enqueued on every lane */
int srpcuSynchronize(void)
{
	int err = synchronizeLane();
	for (int l = 0; l < SRPCU_MAX_LANES && !err; l++)
		if (l != gLane && gLanes[l].ready)
		{
			std::string message;
			{
				LaneScope scope(l);
				err = synchronizeLane();
				if (err) message = g.lastError;
			}
			if (err) g.lastError = message;
		}
	return err;
}

int srpcuClearPlanes(uint32_t* color, float* depth, size_t nPixels)
{
	if (srpcuInit()) return 1;
	srpdLaunchClear(color, depth, nPixels, g.stream);
	g.launches++;
	CU(cudaGetLastError());
	return 0;
}

int srpcuDraw(const SrpdDraw* dIn, const SrpdFrame* framesHost,
              const void* uniforms, size_t uniformBytes, size_t uniformStride)
{
	if (srpcuInit()) return 1;
	SrpdDraw d = *dIn;
	const SrpdState& st = d.st;
	const uint32_t nFrames = d.nFrames;
	if (nFrames == 0 || d.nInputPrims == 0)
		return 0;

	/* A single frame whose uniform fits travels inside the kernel argument blocks (constant
	 * bank, kernels.cuh); everything else -- batches, large or NULL uniforms -- binds per-frame
	 * device copies: one 256-byte aligned block per frame */
	static const bool noInline = getenv("SRP_B200_NO_INLINE_UNIFORM") != nullptr;      /* tests: force the bound path */
	const bool inlineUniform = nFrames == 1 && uniforms && uniformBytes > 0 && uniformBytes <= (size_t) SRPD_INLINE_UNIFORM_BYTES && !noInline;
	const size_t ublock = (uniformBytes + 255) & ~(size_t) 255;
	unsigned char* uniDev = nullptr;
	if (uniforms && uniformBytes && !inlineUniform)
	{
		if (!grow(g.uniforms, ublock * nFrames)) return 1;
		uniDev = (unsigned char*) g.uniforms.ptr;
		/* through the pinned staging ring, so that the caller's uniform memory -- pageable or
		 * pinned -- has been read when the draw call returns (the reference reads it during the
		 * call; a caller may overwrite it right afterwards) */
		const size_t packed = ublock * (nFrames - 1) + uniformBytes;
		unsigned char* staged = ringReserve(packed);
		if (staged)
		{
			for (uint32_t f = 0; f < nFrames; f++)
				memcpy(staged + (size_t) f * ublock, (const unsigned char*) uniforms + (size_t) f * uniformStride, uniformBytes);
			CU(cudaMemcpyAsync(uniDev, staged, packed, cudaMemcpyHostToDevice, g.stream));
		}
		else
		{
			/* larger than the ring: copy and wait */
			if (nFrames == 1)
				CU(cudaMemcpyAsync(uniDev, uniforms, uniformBytes, cudaMemcpyDefault, g.stream));
			else
				CU(cudaMemcpy2DAsync(uniDev, ublock, uniforms, uniformStride, uniformBytes, nFrames, cudaMemcpyDefault, g.stream));
			CU(cudaStreamSynchronize(g.stream));
		}
		g.h2d += uniformBytes * nFrames;
	}

	/* frame bindings */
	SrpdFrame frame0 = framesHost[0];
	frame0.uniform = uniDev;
	const SrpdFrame* framesDev = nullptr;
	if (!inlineUniform)
	{
		if (!grow(g.frames, sizeof(SrpdFrame) * nFrames)) return 1;
		const size_t bytes = sizeof(SrpdFrame) * nFrames;
		SrpdFrame* tmp = (SrpdFrame*) malloc(bytes);
		if (!tmp) { g.lastError = "srp-b200: out of host memory"; return 1; }
		for (uint32_t f = 0; f < nFrames; f++)
		{
			tmp[f] = framesHost[f];
			tmp[f].uniform = uniDev ? uniDev + (size_t) f * ublock : nullptr;
		}
		unsigned char* staged = stageHostData(tmp, bytes);
		cudaError_t e = cudaMemcpyAsync(g.frames.ptr, staged ? (const void*) staged : (const void*) tmp, bytes, cudaMemcpyHostToDevice, g.stream);
		if (e == cudaSuccess && !staged)
			e = cudaStreamSynchronize(g.stream);      /* pageable source larger than the ring */
		free(tmp);
		if (e != cudaSuccess) { fail("cudaMemcpyAsync(frames)", e); return 1; }
		g.h2d += bytes;
		framesDev = (const SrpdFrame*) g.frames.ptr;
	}

	/* scratch sizing: the pools hold the WORST case of this sub-draw (d.maxOutPerInput records
	 * per input primitive), so they cannot overflow whatever the geometry turns out to be; the
	 * host layer splits a draw whose worst case exceeds the budget (srpcuMaxPrimsPerSubDraw). */
	const int nVerts = srpdVertsOfKind(d.kind);
	const uint32_t recStride = srpdRecordStride(st, nVerts);
	uint64_t cap = (uint64_t) d.nInputPrims * d.maxOutPerInput;
	if (cap > 0x3FFFFFF0ull) cap = 0x3FFFFFF0ull;      /* record slots are 30-bit (raster.cu: SRPD_TRI_SLOT_MASK) */
	const uint32_t recCapacity = (uint32_t) cap;
	const uint32_t batchesPerFrame = (d.nInputPrims + SRPD_GEOM_PRIMS - 1) / SRPD_GEOM_PRIMS;

	if (!grow(g.records, (size_t) recCapacity * recStride * nFrames)) return 1;
	if (!grow(g.bboxes, (size_t) recCapacity * sizeof(uint2) * nFrames)) return 1;
	/* one zero-filled block per draw: [header] [per-frame record bump allocators] [tile occupancy bitmap] */
	const uint32_t tilesX = (st.width + SRPD_TILE_W - 1) / SRPD_TILE_W;
	const uint32_t tilesY = (st.height + SRPD_TILE_H - 1) / SRPD_TILE_H;
	const uint32_t occWords = (tilesX * tilesY + 31) / 32;
	const size_t scanStateBytes = (sizeof(uint32_t) * (size_t) nFrames + 7) & ~(size_t) 7;
	const uint32_t chunksPerFrame = (batchesPerFrame + SRPD_SCAN_CHUNK - 1) / SRPD_SCAN_CHUNK;
	const size_t occBytes = (sizeof(uint32_t) * (size_t) occWords * nFrames + 7) & ~(size_t) 7;
	const size_t scanBytes = SRPD_DRAW_HEADER_BYTES + scanStateBytes + occBytes + sizeof(uint2) * (size_t) chunksPerFrame * nFrames;
	if (!grow(g.scan, scanBytes)) return 1;
	if (!grow(g.frameCounts, sizeof(uint32_t) * 2 * nFrames)) return 1;
	CU(cudaMemsetAsync(g.scan.ptr, 0, scanBytes, g.stream));

	mark();
	SrpdGeomArgs ga;
	memset(&ga, 0, sizeof ga);
	ga.d = d;
	ga.frame0 = frame0;
	ga.frames = framesDev;
	if (inlineUniform)
		memcpy(ga.uniformInline, uniforms, uniformBytes);
	ga.records = (unsigned char*) g.records.ptr;
	ga.bboxes = (uint2*) g.bboxes.ptr;
	ga.recCapacity = recCapacity;
	ga.recStride = recStride;
	ga.deferCount = (uint32_t*) g.scan.ptr + 7;
	ga.smCount = (uint32_t) g.smCount;
	ga.abortFlag = (uint32_t*) g.scan.ptr + 1;
	ga.needed = (uint32_t*) g.scan.ptr + 3;
	ga.frameBump = (uint32_t*) ((unsigned char*) g.scan.ptr + SRPD_DRAW_HEADER_BYTES);
	if (!grow(g.ordered, (size_t) recCapacity * sizeof(uint4) * nFrames)) return 1;
	if (!grow(g.idCarry, sizeof(uint32_t) * 2 * (size_t) nFrames)) return 1;
	if (!grow(g.batchInfo, sizeof(uint4) * (size_t) batchesPerFrame * nFrames)) return 1;
	if (!grow(g.deferList, sizeof(uint32_t) * (size_t) batchesPerFrame * nFrames)) return 1;
	ga.deferList = (uint32_t*) g.deferList.ptr;
	ga.ordered = (uint4*) g.ordered.ptr;
	ga.idCarry = (uint32_t*) g.idCarry.ptr + (size_t) ((d.chunkIndex + 1) & 1u) * nFrames;      /* what the previous sub-draw wrote */
	ga.idCarryOut = (uint32_t*) g.idCarry.ptr + (size_t) (d.chunkIndex & 1u) * nFrames;
	ga.batchInfo = (uint4*) g.batchInfo.ptr;
	ga.batchesPerFrame = batchesPerFrame;
	ga.chunksPerFrame = chunksPerFrame;
	ga.chunkSums = (uint2*) ((unsigned char*) g.scan.ptr + SRPD_DRAW_HEADER_BYTES + scanStateBytes + occBytes);
	ga.frameCounts = (uint32_t*) g.frameCounts.ptr;
	ga.occupancy = (uint32_t*) ((unsigned char*) g.scan.ptr + SRPD_DRAW_HEADER_BYTES + scanStateBytes);
	ga.occWordsPerFrame = occWords;
	ga.tilesX = tilesX;
	ga.tilesY = tilesY;
	/* checkpoint table for large triangles: room for 64 framebuffer-sized triangles */
	const uint32_t largeCapacity = 65536;
	uint64_t ckptEntries = 0;
	if (d.kind == SRPD_KIND_TRIANGLE && (st.width > SRPD_LARGE_EXTENT || st.height > SRPD_LARGE_EXTENT))
	{
		ckptEntries = (uint64_t) (tilesX + 1) * st.height * 64;
		if (ckptEntries > (1ull << 27)) ckptEntries = 1ull << 27;     /* 1.5 GiB of float3 at most */
		if (!grow(g.ckptTable, ckptEntries * 3 * sizeof(float))) return 1;
		if (!grow(g.largeList, sizeof(uint2) * largeCapacity)) return 1;
	}
	ga.ckptCursor = (unsigned long long*) ((uint32_t*) g.scan.ptr + 12);
	ga.largeCount = (uint32_t*) g.scan.ptr + 6;
	ga.largeList = (uint2*) g.largeList.ptr;
	ga.largeCapacity = ckptEntries ? largeCapacity : 0;
	ga.ckptCapacity = (uint32_t) ckptEntries;
	ga.stats = g.stats;
	g.launches += (unsigned long long) srpdLaunchGeom(ga, g.stream);
	CU(cudaGetLastError());
	mark();

	/* supertile size from the expected record density: aim at <= ~1.5 k candidates per tile */
	uint32_t superShift = SRPD_SUPER_SHIFT_MAX;
	const uint64_t expected = recCapacity < d.nInputPrims ? recCapacity : d.nInputPrims;
	while (superShift > 1)
	{
		const uint64_t sx = (tilesX + (1u << superShift) - 1) >> superShift, sy = (tilesY + (1u << superShift) - 1) >> superShift;
		if (expected / (sx * sy) <= 2048)
			break;
		const uint64_t nx = (tilesX + (1u << (superShift - 1)) - 1) >> (superShift - 1);
		const uint64_t ny = (tilesY + (1u << (superShift - 1)) - 1) >> (superShift - 1);
		if (nx > 256 || ny > 256 || nx * ny > 8192)
			break;                       /* the 8-bit supertile coordinates / shared-memory cursors would not fit */
		superShift--;
	}
	if (const char* e = getenv("SRP_B200_SUPER_SHIFT")) superShift = (uint32_t) atoi(e);
	if (superShift < 1) superShift = 1;
	if (superShift > SRPD_SUPER_SHIFT_MAX) superShift = SRPD_SUPER_SHIFT_MAX;
	const uint32_t superX = (tilesX + (1u << superShift) - 1) >> superShift;
	const uint32_t superY = (tilesY + (1u << superShift) - 1) >> superShift;
	const uint32_t nSuper = superX * superY;

	/* coarse binning pays once a draw has thousands of primitives: below that every warp tile just
	 * scans all of the draw's records (four launches and ~15 us less on a teapot-sized draw) */
	bool binned = nFrames == 1 && expected > g.binThreshold && nSuper > 1 && nSuper <= 8192
		&& superX <= 256 && superY <= 256;
	if (g.forceBinning == 0) binned = false;
	if (g.forceBinning == 1 && nFrames == 1 && nSuper <= 8192 && superX <= 256 && superY <= 256) binned = true;

	/* barycentric checkpoints of large triangles: needed by the tiles only, so with binning in
	 * between the (usually empty) pre-pass runs beside the binning kernels on the auxiliary stream */
	bool ckptAside = false;
	if (ckptEntries)
	{
		SrpdCkptArgs ca;
		memset(&ca, 0, sizeof ca);
		ca.records = ga.records;
		ca.recCapacity = recCapacity;
		ca.recStride = recStride;
		ca.largeCount = ga.largeCount;
		ca.largeList = ga.largeList;
		ca.largeCapacity = largeCapacity;
		ca.ckptTable = (float*) g.ckptTable.ptr;
		cudaStream_t where = g.stream;
		if (binned && g.ckptAside)
		{
			CU(cudaEventRecord(g.geomDone, g.stream));
			CU(cudaStreamWaitEvent(g.auxStream, g.geomDone, 0));
			where = g.auxStream;
			ckptAside = true;
		}
		srpdLaunchCheckpoints(ca, where);
		g.launches++;
		CU(cudaGetLastError());
		if (ckptAside)
			CU(cudaEventRecord(g.ckptDone, g.auxStream));
	}

	SrpdTileArgs ta;
	memset(&ta, 0, sizeof ta);
	ta.d = d;
	ta.frame0 = frame0;
	ta.frames = framesDev;
	if (inlineUniform)
		memcpy(ta.uniformInline, uniforms, uniformBytes);
	ta.records = ga.records;
	ta.ordered = ga.ordered;
	ta.listOverflow = (uint32_t*) g.scan.ptr + 0;
	ta.recCapacity = recCapacity;
	ta.recStride = recStride;
	ta.frameCounts = ga.frameCounts;
	ta.superX = superX;
	ta.superShift = superShift;
	ta.tilesX = tilesX;
	ta.tilesY = tilesY;
	ta.abortFlag = ga.abortFlag;
	ta.occupancy = ga.occupancy;
	ta.occWordsPerFrame = occWords;
	ta.workCounter = (uint32_t*) g.scan.ptr + 2;
	ta.ckptTable = (const float*) g.ckptTable.ptr;
	ta.smCount = (uint32_t) g.smCount;
	ta.stats = g.stats;
	/* TMA tile stores for the write-back of a single frame into planes of this process */
	if (!framesDev && !(frame0.pad & 1u))
	{
		const bool cd = planeTensorMap(frame0.color, st.width, st.height, 0, &ta.tmColor)
		             && planeTensorMap(frame0.depth, st.width, st.height, 1, &ta.tmDepth);
		if (cd)
			ta.tmaPlanes = 3u | (planeTensorMap(frame0.stencil, st.width, st.height, 2, &ta.tmStencil) ? 4u : 0u);
	}
	if (ta.d.tileRow1 > tilesY) ta.d.tileRow1 = tilesY;
	if (ta.d.tileRow0 > ta.d.tileRow1) ta.d.tileRow0 = ta.d.tileRow1;

	if (binned)
	{
		SrpdBinArgs ba;
		memset(&ba, 0, sizeof ba);
		ba.ordered = ga.ordered;
		ba.frameCounts = ga.frameCounts;
		ba.nChunksMax = (recCapacity + SRPD_BIN_CHUNK - 1) / SRPD_BIN_CHUNK;
		/* draws that store few records use quarter-size chunks (bin.cu): room for those too */
		{
			const uint32_t small = recCapacity < SRPD_BIN_SMALL_RECORDS ? recCapacity : SRPD_BIN_SMALL_RECORDS;
			const uint32_t smallChunks = (small + SRPD_BIN_CHUNK / 4 - 1) / (SRPD_BIN_CHUNK / 4);
			if (ba.nChunksMax < smallChunks) ba.nChunksMax = smallChunks;
		}
		if (nSuper > SRPD_BIN_SMEM_SUPERS)      /* warp-per-chunk fill: chunks of >= 512 records, a bounded number of them */
			ba.nChunksMax = SRPD_BIN_WARP_CHUNKS;
		ba.smCount = (uint32_t) g.smCount;
		ba.superX = superX;
		ba.superY = superY;
		ba.superShift = superShift;
		/* coarse lists: their worst case (every record in every supertile) is not affordable, so
		 * the pool is sized generously and a draw that still exceeds it falls back, on the device,
		 * to tiles that scan all records (bin.cu); the note it leaves raises the pool for later draws */
		if (g.hostNotes[1])
		{
			const uint64_t need = g.hostNotes[1];
			if (need + need / 8 + 1024 > g.listFloor) g.listFloor = need + need / 8 + 1024;
			g.hostNotes[1] = 0;
		}
		uint64_t listCap = (uint64_t) (expected < recCapacity ? expected : recCapacity) * (d.kind == SRPD_KIND_LINE ? 8 : 4) + (uint64_t) nSuper * 64 + 65536;
		if (listCap < g.listFloor) listCap = g.listFloor;
		if (const char* e = getenv("SRP_B200_LIST_CAPACITY")) listCap = (uint64_t) atoll(e);      /* tests: force the fallback */
		if (listCap > 0x7FFFFFF0ull) listCap = 0x7FFFFFF0ull;
		ba.listCapacity = (uint32_t) listCap;
		if (!grow(g.chunkCounts, sizeof(uint32_t) * (size_t) ba.nChunksMax * nSuper)) return 1;
		if (!grow(g.superOffsets, sizeof(uint32_t) * (nSuper + 1))) return 1;
		if (!grow(g.superTotals, sizeof(uint32_t) * nSuper)) return 1;
		if (!grow(g.listIds, sizeof(uint4) * (size_t) ba.listCapacity)) return 1;
		ba.chunkCounts = (uint32_t*) g.chunkCounts.ptr;
		ba.superOffsets = (uint32_t*) g.superOffsets.ptr;
		ba.superTotals = (uint32_t*) g.superTotals.ptr;
		ba.listEntries = (uint4*) g.listIds.ptr;
		ba.listOverflow = (uint32_t*) g.scan.ptr + 0;
		ba.needed = ga.needed;
		ba.scanTicket = (uint32_t*) g.scan.ptr + 5;
		ba.hostNotes = g.hostNotesDev;
		ba.stats = g.stats;
		srpdLaunchBin(ba, g.stream, ckptAside ? g.ckptDone : (cudaEvent_t) nullptr);     /* (ckptAside implies binned) */
		g.launches += 3;
		CU(cudaGetLastError());
		ta.superOffsets = ba.superOffsets;
		ta.listEntries = ba.listEntries;
	}

	{
		/* work-item granularity (items = groups of consecutive 32x8 warp tiles, pulled by warps):
		 * enough items for dynamic load balance (>= ~16 per warp), but not one atomic per tile
		 * when a batch has millions of mostly empty tiles */
		const uint32_t rows = ta.d.tileRow1 - ta.d.tileRow0;
		const uint64_t tiles = (uint64_t) tilesX * rows * 2u * nFrames;
		const uint64_t target = (uint64_t) g.smCount * SRPD_TILE_CTAS_PER_SM * SRPD_TILE_WARPS * 16;
		uint32_t per = 1;
		while (per < 32 && tiles / (per * 2) >= target) per *= 2;
		if (const char* e = getenv("SRP_B200_TILES_PER_ITEM")) per = (uint32_t) atoi(e) > 0 ? (uint32_t) atoi(e) : per;
		ta.tilesPerItem = per;
	}
	/* one-shot mirror request: rasterise in bands and download each band while the next one
	 * is being rasterised (only worth it for big frames; small ones keep the single launch) */
	const SrpcuMirror mirror = g.mirror;
	int* mirrorDone = g.mirrorDone;
	g.mirror = SrpcuMirror{ nullptr, nullptr, nullptr };
	g.mirrorDone = nullptr;
	const uint32_t rows = ta.d.tileRow1 - ta.d.tileRow0;
	uint32_t nBands = 1;
	if ((mirror.color || mirror.depth) && nFrames == 1 && ta.d.tileRow0 == 0 && ta.d.tileRow1 == tilesY
	    && (uint64_t) st.width * st.height >= (1u << 20) && !getenv("SRP_B200_NO_BANDS"))
		nBands = rows >= 64 ? 4 : 1;
	mark();
	if (nBands == 1)
	{
		srpdLaunchTiles(ta, g.stream);
		g.launches++;
		CU(cudaGetLastError());
	}
	else
	{
		const size_t W = (size_t) st.width;
		for (uint32_t b = 0; b < nBands; b++)
		{
			SrpdTileArgs band = ta;
			band.d.tileRow0 = rows * b / nBands;
			band.d.tileRow1 = rows * (b + 1) / nBands;
			band.workCounter = (uint32_t*) g.scan.ptr + 8 + b;
			srpdLaunchTiles(band, g.stream);
			g.launches++;
			CU(cudaGetLastError());
			CU(cudaEventRecord(g.bandEvents[b], g.stream));
			CU(cudaStreamWaitEvent(g.copyStream, g.bandEvents[b], 0));
			const size_t y0 = (size_t) band.d.tileRow0 * SRPD_TILE_H;
			size_t y1 = (size_t) band.d.tileRow1 * SRPD_TILE_H;
			if (y1 > (size_t) st.height) y1 = (size_t) st.height;
			const size_t first = y0 * W, count = (y1 - y0) * W;
			if (mirror.color)
			{
				CU(cudaMemcpyAsync((uint32_t*) mirror.color + first, frame0.color + first, count * 4, cudaMemcpyDeviceToHost, g.copyStream));
				g.d2h += count * 4;
			}
			if (mirror.depth)
			{
				CU(cudaMemcpyAsync((float*) mirror.depth + first, frame0.depth + first, count * 4, cudaMemcpyDeviceToHost, g.copyStream));
				g.d2h += count * 4;
			}
			if (mirror.stencil)
			{
				CU(cudaMemcpyAsync((uint8_t*) mirror.stencil + first, frame0.stencil + first, count, cudaMemcpyDeviceToHost, g.copyStream));
				g.d2h += count;
			}
		}
		/* later work on the main stream (the next draw) must not overtake the downloads */
		CU(cudaEventRecord(g.copyDone, g.copyStream));
		CU(cudaStreamWaitEvent(g.stream, g.copyDone, 0));
		g.copyPending = true;
		if (mirrorDone) *mirrorDone = 1;
	}
	mark();
	return 0;
}

static void addLaneStats(SrpdStats* out, unsigned long long* launches, unsigned long long* h2d, unsigned long long* d2h)
{
	if (launches) *launches += g.launches;
	if (h2d) *h2d += g.h2d;
	if (d2h) *d2h += g.d2h;
	if (!g.ready)
		return;
	cudaMemcpyAsync(g.statsHost, g.stats, sizeof(SrpdStats) * SRPD_STATS_SLOTS, cudaMemcpyDeviceToHost, g.stream);
	cudaStreamSynchronize(g.stream);
	for (int i = 0; i < SRPD_STATS_SLOTS; i++)
	{
		out->primsIn += g.statsHost[i].primsIn;
		out->primsEmitted += g.statsHost[i].primsEmitted;
		out->primsStored += g.statsHost[i].primsStored;
		out->fragsEmitted += g.statsHost[i].fragsEmitted;
		out->fragsShaded += g.statsHost[i].fragsShaded;
		out->overflow += g.statsHost[i].overflow;
	}
}
/* counters summed over the lanes */
void srpcuGetStats(SrpdStats* out, unsigned long long* launches, unsigned long long* h2d, unsigned long long* d2h)
{
	memset(out, 0, sizeof *out);
	if (launches) *launches = 0;
	if (h2d) *h2d = 0;
	if (d2h) *d2h = 0;
	for (int l = 0; l < SRPCU_MAX_LANES; l++)
	{
		LaneScope scope(l);
		addLaneStats(out, launches, h2d, d2h);
	}
}

/* Guard: 1 if a draw hit a record-pool limit since the previous call (synchronises).  The pools
 * are sized for the worst case of every sub-draw, so this is never expected; a draw that does
 * trip it leaves its framebuffer (and a pending clear) untouched, and the host reports it.
 * Only slot 0 of the counter array is used for this accounting. */
static int takeLaneOverflow(void)
{
	unsigned long long& seen = g.guardSeen;
	if (!g.ready)
		return 0;
	cudaMemcpyAsync(&g.statsHost[0].overflow, &g.stats[0].overflow, sizeof(unsigned long long), cudaMemcpyDeviceToHost, g.stream);
	cudaStreamSynchronize(g.stream);
	const unsigned long long now = g.statsHost[0].overflow;
	const int fresh = now > seen;
	seen = now;
	return fresh;
}
int srpcuTakeOverflow(void)
{
	int fresh = 0;
	for (int l = 0; l < SRPCU_MAX_LANES; l++)
	{
		LaneScope scope(l);
		fresh |= takeLaneOverflow();
	}
	return fresh;
}

/* How many input primitives of `d` one sub-draw may take so that the worst-case per-record
 * scratch (record + box + ordered-view entry, for every frame of a batch) stays within the
 * budget (SRP_B200_POOL_BUDGET_MB, default an eighth of the device memory). */
uint32_t srpcuMaxPrimsPerSubDraw(const SrpdDraw* d)
{
	if (srpcuInit()) return d->nInputPrims;
	const uint64_t perRecord = (uint64_t) srpdRecordStride(d->st, srpdVertsOfKind(d->kind)) + sizeof(uint2) + sizeof(uint4);
	const uint64_t perPrim = perRecord * (d->maxOutPerInput ? d->maxOutPerInput : 1) * (d->nFrames ? d->nFrames : 1);
	uint64_t n = g.poolBudget / (perPrim ? perPrim : 1);
	const uint64_t byIndex = 0x3FFFFFF0ull / ((uint64_t) (d->maxOutPerInput ? d->maxOutPerInput : 1) * (d->nFrames ? d->nFrames : 1));      /* 30-bit record slots */
	if (n > byIndex) n = byIndex;
	if (n < (uint64_t) SRPD_GEOM_PRIMS) n = SRPD_GEOM_PRIMS;
	n -= n % SRPD_GEOM_PRIMS;      /* whole batches */
	return n > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t) n;
}

/* ---- peer memory and stream-ordered flags (sort-first strips, multigpu.py) ---- */
__global__ void srpdSignalKernel(uint32_t* flag, uint32_t value)
{
	srpdGridDependencyEnter();
	/* everything this stream ran before has completed (stream order); make it -- peer writes of
	 * the tile kernel included -- visible system-wide before the flag */
	__threadfence_system();
	*(volatile uint32_t*) flag = value;
	__threadfence_system();
}
__global__ void srpdWaitFlagKernel(const uint32_t* flag, uint32_t value, uint32_t* hostNotes)
{
	srpdGridDependencyEnter();
	flag += threadIdx.x;      /* (srpcuStreamWaitFlags: one thread per flag of a consecutive run) */
	/* bounded: a peer that died must not hang this GPU for good (~10 s, then the stream moves on
	 * and the host is told through hostNotes[2]) */
	for (uint32_t spins = 0; *(volatile const uint32_t*) flag < value; spins++)
	{
		if (spins > 20000000u)
		{
			*(volatile uint32_t*) (hostNotes + 2) = value;
			break;
		}
		__nanosleep(500);
	}
	__threadfence_system();
}

int srpcuIpcExport(const void* devicePtr, unsigned char handle[64])
{
	if (srpcuInit()) return 1;
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 bytes");
	cudaIpcMemHandle_t h;
	CU(cudaIpcGetMemHandle(&h, const_cast<void*>(devicePtr)));
	memcpy(handle, &h, 64);
	return 0;
}
void* srpcuIpcOpen(const unsigned char handle[64])
{
	if (srpcuInit()) return nullptr;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, 64);
	void* p = nullptr;
	cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
	if (e != cudaSuccess) { fail("cudaIpcOpenMemHandle", e); return nullptr; }
	return p;
}
int srpcuIpcClose(void* mapped)
{
	if (!g.ready || !mapped) return 0;
	CU(cudaStreamSynchronize(g.stream));
	CU(cudaIpcCloseMemHandle(mapped));
	return 0;
}
int srpcuStreamSignal(uint32_t* flag, uint32_t value)
{
	if (srpcuInit()) return 1;
	srpdSignalKernel<<<1, 1, 0, g.stream>>>(flag, value);
	g.launches++;
	CU(cudaGetLastError());
	return 0;
}
int srpcuStreamWaitFlag(const uint32_t* flag, uint32_t value)
{
	return srpcuStreamWaitFlags(flag, 1u, value);
}
/* one kernel that holds the stream until ALL of flag[0 .. count) are >= value (count <= 1024) */
int srpcuStreamWaitFlags(const uint32_t* flag, uint32_t count, uint32_t value)
{
	if (srpcuInit()) return 1;
	if (count == 0u) return 0;
	if (count > 1024u) { g.lastError = "srp-b200: at most 1024 flags per wait"; return 1; }
	srpdWaitFlagKernel<<<1, count, 0, g.stream>>>(flag, value, g.hostNotesDev);
	g.launches++;
	CU(cudaGetLastError());
	return 0;
}

/* per-stage device time: enable, run draws, collect {geometry, binning, tiles} in ms */
void srpcuSetProfiling(int on)
{
	gProfile = on != 0;
	for (int l = 0; l < SRPCU_MAX_LANES; l++)
		gLanes[l].profile = gProfile;
}
static unsigned long long collectLaneStageTimes(double outMs[3])
{
	if (g.ready)
	{
		cudaStreamSynchronize(g.stream);
		for (size_t i = 0; i + 3 < g.pendingEvents.size(); i += 4)
		{
			for (int k = 0; k < 3; k++)
			{
				float ms = 0.f;
				if (cudaEventElapsedTime(&ms, g.pendingEvents[i + k], g.pendingEvents[i + k + 1]) == cudaSuccess)
					g.stageMs[k] += ms;
			}
			g.stageDraws++;
		}
		for (cudaEvent_t e : g.pendingEvents) g.freeEvents.push_back(e);
		g.pendingEvents.clear();
	}
	for (int k = 0; k < 3; k++) { outMs[k] = g.stageMs[k]; g.stageMs[k] = 0; }
	const unsigned long long n = g.stageDraws;
	g.stageDraws = 0;
	return n;
}
unsigned long long srpcuCollectStageTimes(double outMs[3])
{
	unsigned long long n = 0;
	outMs[0] = outMs[1] = outMs[2] = 0;
	for (int l = 0; l < SRPCU_MAX_LANES; l++)
	{
		LaneScope scope(l);
		double ms[3];
		n += collectLaneStageTimes(ms);
		for (int k = 0; k < 3; k++) outMs[k] += ms[k];
	}
	return n;
}

static void resetLaneStats(void)
{
	g.launches = 0; g.h2d = 0; g.d2h = 0;
	if (g.ready)
	{
		/* keep the overflow counter monotonic (srpcuTakeOverflow compares against the
		 * last value it saw), zero everything else */
		takeLaneOverflow();
		cudaMemsetAsync(g.stats, 0, sizeof(SrpdStats) * SRPD_STATS_SLOTS, g.stream);
		cudaMemcpyAsync(&g.stats[0].overflow, &g.statsHost[0].overflow, sizeof(unsigned long long), cudaMemcpyHostToDevice, g.stream);
		cudaStreamSynchronize(g.stream);
	}
}
void srpcuResetStats(void)
{
	for (int l = 0; l < SRPCU_MAX_LANES; l++)
	{
		LaneScope scope(l);
		resetLaneStats();
	}
}

} // extern "C"
