/* srp-b200 -- the additive extension surface next to the srp C API (include/srp/api.h).
 *
 * Existing srp programs keep using the srpNew..., srp...CopyData and srpDraw...Buffer calls unchanged.  What
 * they add is one CUDA translation unit with the __device__ twins of their shaders
 * (include/srp_b200_device.cuh), which registers the twins here.  Everything else in
 * this header is optional: synchronisation policy, zero-copy / batched entry points
 * used by the benchmarks and the multi-GPU modes, and counters.
 *
 * All functions are plain C ABI (pointers and sizes only). */
#ifndef SRP_B200_H_
#define SRP_B200_H_
#include "srp/api.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- shader registry ---------------------------------------------------------------
 * Replaces the reference's direct calls through sp->vs->shader / sp->fs->shader
 * (src/pipeline/vertex_processing.c:73, src/raster/fragment.c:98): at draw time the
 * host function pointers found in SRPShaderProgram are looked up here and the draw
 * runs device program `deviceProgramId` of the program table linked into the
 * executable.  `uniformSize` is sizeof the user's uniform struct (the API passes it
 * un-sized); it is copied to the device at every draw.  Returns 0 on success.
 * A draw whose program is not registered reports a HIGH error through the message
 * callback and draws nothing -- there is no CPU fallback. */
typedef void (*SRPVertexShaderFunc)(SRPVertexShaderIn*, SRPVertexShaderOut*);
typedef void (*SRPFragmentShaderFunc)(SRPFragmentShaderIn*, SRPFragmentShaderOut*);
int srpB200RegisterProgram(SRPVertexShaderFunc hostVS, SRPFragmentShaderFunc hostFS,
                           int deviceProgramId, size_t uniformSize);
/* The same, one shader at a time: programs that recombine shaders at run time (the reference's
 * tests/scenes/clipping/point.c swaps the fragment shader of a copied SRPShaderProgram) register
 * every host shader function once; a draw then looks its two functions up independently and the
 * uniform block copied is the larger of the two sizes.  `srp_b200/twingen.py` writes these
 * registrations -- and the __device__ twins themselves -- from a program's C source. */
int srpB200RegisterVertexShader(SRPVertexShaderFunc hostVS, int deviceShaderId, size_t uniformSize);
int srpB200RegisterFragmentShader(SRPFragmentShaderFunc hostFS, int deviceShaderId, size_t uniformSize);

/* ---- synchronisation policy --------------------------------------------------------
 * The device planes of a framebuffer are authoritative; fb->color/depth/stencil are a
 * pinned host mirror.
 *   SRP_B200_SYNC_DRAW      (default) every srpDraw*Buffer returns with the mirror up to
 *                           date, like the reference.  srpFramebufferClear is deferred
 *                           and fused into the next draw (or download).
 *   SRP_B200_SYNC_EXPLICIT  draws only enqueue work; the mirror is refreshed by
 *                           srpB200FramebufferDownload() / srpB200Finish(). */
typedef enum { SRP_B200_SYNC_DRAW = 0, SRP_B200_SYNC_EXPLICIT = 1 } SRPB200SyncMode;
void srpB200SetSyncMode(SRPB200SyncMode mode);
SRPB200SyncMode srpB200GetSyncMode(void);
/* Which planes the automatic refresh of the host mirror covers (default: all three).  A
 * program that only ever reads fb->color (like the reference's examples and tests) can drop
 * depth and stencil and halve the PCIe traffic of every draw; explicit downloads still bring
 * everything. */
enum { SRP_B200_MIRROR_COLOR = 1, SRP_B200_MIRROR_DEPTH = 2, SRP_B200_MIRROR_STENCIL = 4, SRP_B200_MIRROR_ALL = 7 };
void srpB200SetMirrorPlanes(int planeMask);
void srpB200Finish(void);                                   /* wait for all enqueued work */
void srpB200FramebufferDownload(const SRPFramebuffer* fb);  /* device planes -> host mirror (synchronous) */
void srpB200FramebufferUpload(const SRPFramebuffer* fb);    /* host mirror -> device planes */
/* Pipelined frames under SRP_B200_SYNC_EXPLICIT (the per-draw wait of the
 * reference's synchronous srpDraw*Buffer contract, examples/03_textured_cube.c:133-160, is what
 * bounds a frame loop once rendering takes microseconds).  DownloadAsync enqueues the
 * device->host copy of the planes selected by srpB200SetMirrorPlanes behind the draws issued
 * so far and returns; Wait blocks until that framebuffer's mirror is complete.  A program that
 * alternates between two framebuffers renders frame i+1 (uploads included) while frame i
 * crosses PCIe; a draw into a framebuffer whose download is still in flight is ordered
 * behind it, so results never tear. */
void srpB200FramebufferDownloadAsync(const SRPFramebuffer* fb);
void srpB200FramebufferWait(const SRPFramebuffer* fb);
/* stream-side counterpart of Wait: everything enqueued after this call is ordered behind the
 * framebuffer's in-flight asynchronous download (the host does not block) */
void srpB200FramebufferFence(const SRPFramebuffer* fb);
/* Frames in flight.  The library has srpB200LaneCount() lanes; a lane is a CUDA stream with its
 * own scratch pools and staging.  Work enqueued on one lane runs in order; work on different lanes
 * is independent and overlaps on the device -- the geometry and binning kernels of one frame,
 * which leave most of the GPU idle, run under the tile kernel of another.  A frame loop under
 * SRP_B200_SYNC_EXPLICIT that renders independent frames selects lane (frame % 2) before the
 * frame's clear and draws, into one framebuffer per lane.  Ordering that still holds across lanes:
 * a framebuffer used on a lane other than the one that touched it last is ordered behind that
 * lane's work so far; srp*BufferCopyData is ordered behind every lane's earlier draws and before
 * every lane's later ones; srpB200Finish() waits for all lanes; counters sum over lanes.
 * Everything else (srpB200Stream, srpB200StreamSignal/Wait) refers to the current lane.
 * Lane 0 is current at start; SetLane returns 0 on success. */
int srpB200SetLane(int lane);
int srpB200GetLane(void);
int srpB200LaneCount(void);

/* ---- device-resident objects -------------------------------------------------------
 * Framebuffer on caller-owned device memory (e.g. planes of a torch tensor that NCCL
 * gathers); color/depth/stencil of the returned struct still point to a host mirror. */
SRPFramebuffer* srpB200NewFramebufferOnDevice(size_t width, size_t height,
                                              void* deviceColor, void* deviceDepth, void* deviceStencil);
/* device pointers of the planes: which = 0 colour (u32), 1 depth (f32), 2 stencil (u8) */
void* srpB200FramebufferDevicePlane(const SRPFramebuffer* fb, int which);
/* texture from RGB8 texels in memory (the file loader srpNewTexture only reads PNG) */
SRPTexture* srpB200NewTextureFromMemory(const uint8_t* rgb, int width, int height,
                                        SRPTextureWrappingMode wrappingModeX, SRPTextureWrappingMode wrappingModeY);

/* ---- I/O neighbours of the draw path ------------------------------------------------
 * What a program does right before it fills its buffers and right after it has its pixels
 * (plain host code).  srpB200LoadOBJ reads a Wavefront OBJ the way the reference's examples do
 * (examples/utility/objparser.c:8-79, `loadOBJMesh`): every corner of an
 * `f p/t/n p/t/n p/t/n` face becomes its own vertex {vec3 position, vec2 uv, vec3 normal}
 * (32 bytes, the reference's OBJVertex) and the indices are 0..n-1 -- the same arrays bit for
 * bit, ready for srpVertexBufferCopyData / srpIndexBufferCopyData(SRP_UINT32) -- without that
 * loader's fixed 65536-element tables.  Returns 0 on success; other face forms are skipped with
 * a warning.  srpB200WritePNG / srpB200SaveFramebufferPNG write the colour plane as an 8-bit
 * RGBA PNG with alpha 255 (tests/utils/save.c:4-33, `saveFramebufferToImage`); the framebuffer
 * variant first brings the host mirror up to date.  Both return 0 on success. */
typedef struct SRPB200Mesh
{
	float* vertices;            /* vertexCount * 8 floats: position.xyz, uv.xy, normal.xyz */
	size_t vertexCount;
	size_t bytesPerVertex;      /* 32 */
	uint32_t* indices;          /* indexCount entries, 0..vertexCount-1 */
	size_t indexCount;
} SRPB200Mesh;
int srpB200LoadOBJ(const char* path, SRPB200Mesh* mesh);
void srpB200FreeMesh(SRPB200Mesh* mesh);
int srpB200WritePNG(const char* path, size_t width, size_t height, const uint32_t* color);
int srpB200SaveFramebufferPNG(const SRPFramebuffer* fb, const char* path);

/* ---- frame-parallel batch ----------------------------------------------------------
 * One call = nFrames independent srpFramebufferClear + srpDrawIndexBuffer pairs that
 * share buffers, program and context state and differ in uniform and target:
 * frame f reads uniform block `(char*) uniforms + f * uniformStride` and renders into
 * fbs[f].  `clearFirst` != 0 applies srpFramebufferClear semantics to every target
 * first.  ib may be NULL (vertex-buffer draw). */
void srpB200DrawBatch(const SRPIndexBuffer* ib, const SRPVertexBuffer* vb,
                      SRPFramebuffer* const* fbs, size_t nFrames,
                      const SRPShaderProgram* sp, const void* uniforms, size_t uniformStride,
                      SRPPrimitive primitive, size_t startIndex, size_t count, int clearFirst);

/* ---- sort-first strips -------------------------------------------------------------
 * Restrict rasterisation (not geometry processing, so primitive ids are unchanged) to
 * framebuffer rows [row0, row1); row0 is rounded down and row1 up to the tile height
 * (srpB200TileHeight()).  Pass (0, SIZE_MAX) to reset. */
void srpB200SetRowRange(size_t row0, size_t row1);
size_t srpB200TileWidth(void);
size_t srpB200TileHeight(void);

/* Multi-GPU form of the strips (one process per GPU): the framebuffer lives on
 * ONE GPU (the root); the other ranks map its three planes through CUDA IPC and their tile
 * kernels write their strips straight into the root's memory over NVLink as part of the tile
 * write-back -- no staging copy and no separate gather step.
 *   root:   srpB200IpcExport() of srpB200FramebufferDevicePlane(fb, 0..2) -> three 64-byte handles
 *           that travel to the other ranks (torch.distributed / any byte transport);
 *   others: srpB200IpcOpen() each handle, srpB200NewFramebufferOnDevice() on the mapped planes,
 *           srpB200SetRowRange(own strip), then the usual srpFramebufferClear + srpDraw*Buffer.
 * Completion is signalled through 32-bit flags in (peer) device memory with two stream-ordered
 * calls: srpB200StreamSignal stores `value` once everything enqueued before it has finished and
 * its writes are visible system-wide; srpB200StreamWait holds the current lane's stream until the
 * flag is >= value.  srpB200DeviceAlloc gives zero-filled device memory that can be exported. */
typedef struct SRPB200IpcHandle { unsigned char bytes[64]; } SRPB200IpcHandle;
void* srpB200DeviceAlloc(size_t bytes);
void srpB200DeviceFree(void* devicePtr);
int srpB200IpcExport(const void* devicePtr, SRPB200IpcHandle* handle);   /* 0 on success */
void* srpB200IpcOpen(const SRPB200IpcHandle* handle);                     /* NULL on failure */
void srpB200IpcClose(void* mappedPtr);
void srpB200StreamSignal(uint32_t* flag, uint32_t value);
void srpB200StreamWait(const uint32_t* flag, uint32_t value);
/* the same for a run of consecutive flags: the stream goes on once ALL of flags[0 .. count) are >= value
 * (one launch instead of count; count <= 1024) */
void srpB200StreamWaitAll(const uint32_t* flags, uint32_t count, uint32_t value);

/* ---- counters ----------------------------------------------------------------------
 * Accumulated since the last reset over all draws (deterministic; equal to the
 * reference's emitFragment / fragment-shader call counts for the same input). */
typedef struct SRPB200Stats
{
	unsigned long long draws;
	unsigned long long primsIn, primsEmitted, primsStored;
	unsigned long long fragsEmitted, fragsShaded;
	unsigned long long kernelLaunches;     /* kernels of this library launched            */
	unsigned long long h2dBytes, d2hBytes; /* bytes this library copied across PCIe       */
	unsigned long long overflow;           /* guard, expected 0: draws that hit a scratch-pool limit (pools hold the worst case) */
	unsigned long long subDraws;           /* kernel chains submitted: > draws when a draw's worst case exceeded the
	                                          scratch budget and it was split into ranges of input primitives */
} SRPB200Stats;
void srpB200GetStats(SRPB200Stats* out);   /* synchronises */
void srpB200ResetStats(void);

/* Per-stage device time.  While enabled every draw records four CUDA events on the
 * submission stream (before geometry, after geometry, after binning, after the tiles);
 * collecting synchronises, returns the number of draws measured and the summed
 * milliseconds of {geometry, binning, tiles} since the previous collection. */
void srpB200SetProfiling(int enable);
unsigned long long srpB200CollectStageTimes(double outMs[3]);

/* device / build identification, e.g. "srp-b200 sm_100a tile 32x16" */
const char* srpB200Version(void);
/* select the CUDA device for the calling process (before any other call); default:
 * environment SRP_B200_DEVICE, else LOCAL_RANK, else 0 */
void srpB200SetDevice(int device);
/* raw CUDA stream (cudaStream_t) the current lane's work is enqueued on, for event timing */
void* srpB200Stream(void);

#ifdef __cplusplus
}
#endif
#endif /* SRP_B200_H_ */
