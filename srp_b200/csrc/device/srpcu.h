/* srp-b200 internal -- the thin C ABI between the host C layer (csrc/host) and the
 * CUDA layer (csrc/device).  Host code never includes CUDA headers; it sees device
 * memory as opaque pointers and submits one SrpdDraw per draw call. */
#ifndef SRPCU_H_
#define SRPCU_H_
#include <stddef.h>
#include <stdint.h>
#include "draw_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Lazily creates the per-process runtime (device selection, stream, scratch pools).
 * Returns 0 on success; on failure srpcuLastError() explains (no GPU / wrong arch /
 * extension built without its program table ...).  There is no CPU fallback. */
int srpcuInit(void);
void srpcuSetDevice(int device);
const char* srpcuLastError(void);
const char* srpcuVersion(void);
void* srpcuStream(void);

void* srpcuMalloc(size_t bytes);                 /* device memory, zero-filled           */
void srpcuFree(void* p);
void* srpcuMallocHost(size_t bytes);             /* pinned host memory, zero-filled      */
void srpcuFreeHost(void* p);
void* srpcuMallocManaged(size_t bytes);          /* managed memory (textures)            */
void srpcuFreeManaged(void* p);
void srpcuPrefetchToDevice(void* p, size_t bytes);
/* copy of caller memory of any kind (pageable, pinned, device) to the device with memcpy
 * semantics: the caller may reuse `src` on return.  Runs on the upload stream behind `lastUse`
 * (event of the last draw that reads `dst`, or NULL); later draws are ordered behind it. */
int srpcuUpload(void* dst, const void* src, size_t bytes, void* lastUse);
int srpcuUploadInStream(void* dst, const void* src, size_t bytes);   /* on the submission stream; enqueue only */
int srpcuRecordEvent(void* event);                                  /* on the submission stream */
int srpcuDownload(void* dstHost, const void* srcDevice, size_t bytes);   /* enqueue only */
int srpcuSynchronize(void);                                         /* every lane */

/* Lanes: SRPCU_MAX_LANES independent submission states (stream + scratch pools + staging).  Work
 * enqueued on different lanes is unordered and overlaps on the device; within a lane it runs in
 * order.  All entry points act on the current lane unless they say otherwise. */
#define SRPCU_MAX_LANES 4
int srpcuSetLane(int lane);                                         /* 0 on success */
int srpcuLane(void);
int srpcuLaneCount(void);
int srpcuOrderBehindLane(int other);     /* what `other` has enqueued so far happens before the current lane's later work */

/* Apply a pending clear to real memory: colour 0, depth -1 (stencil untouched). */
int srpcuClearPlanes(uint32_t* color, float* depth, size_t nPixels);

/* Enqueue one draw (all its kernels) for d->nFrames frames.  `frames` is a host array
 * of nFrames bindings whose `uniform` fields are ignored: frame f uses the uniform
 * block at uniforms + f*uniformStride (uniformBytes each; NULL / 0 = no uniform). */
int srpcuDraw(const SrpdDraw* d, const SrpdFrame* frames,
              const void* uniforms, size_t uniformBytes, size_t uniformStride);

/* Host mirror to refresh as part of the NEXT single-frame srpcuDraw (one-shot; cleared by that
 * draw).  When set, the tile kernel runs in horizontal bands and each band's rows are copied
 * to the host on a second stream while the next band is rasterised; srpcuSynchronize() then
 * waits for both.  Returns through *done whether the draw took care of the download. */
typedef struct SrpcuMirror
{
	void* color; void* depth; void* stencil;   /* pinned host planes; stencil may be NULL (not needed) */
} SrpcuMirror;
void srpcuSetMirrorForNextDraw(const SrpcuMirror* mirror, int* done);

/* Asynchronous download (explicit synchronisation policy): the planes named in `host` are
 * copied to the pinned host mirror on the copy stream once everything enqueued so far on the
 * submission stream has finished; `doneEvent` (srpcuNewEvent) is recorded behind the copies.
 * The submission stream is NOT held up: later draws into OTHER framebuffers overlap with the
 * copies.  Before anything overwrites the source planes the caller makes the submission
 * stream wait for the event (srpcuStreamWaitEvent); the host waits with srpcuHostWaitEvent. */
void* srpcuNewEvent(void);
void srpcuFreeEvent(void* event);
int srpcuDownloadPlanesAsync(const SrpcuMirror* host, const void* dColor, const void* dDepth, const void* dStencil,
                             size_t nPixels, void* doneEvent);
int srpcuHostWaitEvent(void* event);
int srpcuStreamWaitEvent(void* event);

/* Guard: 1 if a draw hit a record-pool limit since the previous call (never expected: the pools
 * hold the worst case of every sub-draw); synchronises. */
int srpcuTakeOverflow(void);
/* input primitives one sub-draw of `d` may take within the scratch budget (whole batches) */
uint32_t srpcuMaxPrimsPerSubDraw(const SrpdDraw* d);

/* ---- multi-GPU plumbing (one process per GPU): peer memory over NVLink ----
 * Export device memory of this library (a framebuffer plane, srpcuMalloc) as a CUDA IPC handle
 * (64 bytes), map another process's export into this one (peer access is enabled lazily), and
 * two stream-ordered primitives on 32-bit flags that may live in peer memory: signal = store
 * `value` once everything enqueued before has finished and is visible system-wide; wait = hold
 * the stream until the flag is >= `value`. */
int srpcuIpcExport(const void* devicePtr, unsigned char handle[64]);
void* srpcuIpcOpen(const unsigned char handle[64]);
int srpcuIpcClose(void* mapped);
int srpcuStreamSignal(uint32_t* flag, uint32_t value);
int srpcuStreamWaitFlag(const uint32_t* flag, uint32_t value);
int srpcuStreamWaitFlags(const uint32_t* flag, uint32_t count, uint32_t value);      /* all of flag[0 .. count) >= value, one kernel */

void srpcuSetProfiling(int on);
unsigned long long srpcuCollectStageTimes(double outMs[3]);

void srpcuGetStats(SrpdStats* out, unsigned long long* launches, unsigned long long* h2d, unsigned long long* d2h);
void srpcuResetStats(void);

int srpcuTileWidth(void);
int srpcuTileHeight(void);

#ifdef __cplusplus
}
#endif
#endif
