/* srp-b200 host layer -- private definitions of the opaque API objects.
 * Host code is C (gcc -std=c2x, i.e. no FP contraction -- the same arithmetic regime as
 * the reference, so host-side matrix constructors produce identical uniforms) and talks
 * to CUDA only through ../device/srpcu.h. */
#ifndef SRP_INTERNAL_H_
#define SRP_INTERNAL_H_
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>
#include "srp/api.h"
#include "srp_b200.h"
#include "../device/srpcu.h"

/* reference: src/core/buffer_p.h:15-30 -- here the payload lives in device memory */
struct SRPVertexBuffer
{
	size_t nBytesPerVertex;
	size_t nVertices;
	size_t nBytesAllocated;
	void* data;                 /* device */
	void* lastUse;              /* event behind the last asynchronous draw that reads `data` (explicit policy) */
};

struct SRPIndexBuffer
{
	SRPType indicesType;
	size_t nBytesPerIndex;
	size_t nIndices;
	size_t nBytesAllocated;
	void* data;                 /* device */
	void* lastUse;              /* as in SRPVertexBuffer */
};

/* The public SRPFramebuffer (host-visible mirror pointers) is the first member, so the
 * pointer handed to the user is also the pointer to this object. */
typedef struct SRPFramebufferImpl
{
	SRPFramebuffer pub;
	uint32_t magic;
	uint32_t* dColor;           /* device planes: authoritative */
	float* dDepth;
	uint8_t* dStencil;
	bool ownsDevicePlanes;
	bool clearPending;          /* srpFramebufferClear deferred into the next draw */
	bool mirrorStale;           /* device planes changed since the last download   */
	bool stencilTouched;        /* a stencil-enabled draw ran since the last download */
	bool downloadInFlight;      /* srpB200FramebufferDownloadAsync not yet waited for */
	void* downloadEvent;        /* recorded behind the asynchronous download's copies */
	int lane;                   /* lane whose stream the last work on the planes was enqueued on */
} SRPFramebufferImpl;
#define SRP_FB_MAGIC 0x53524246u   /* "SRBF" */

/* per-context runtime object behind SRPContext.arena */
struct SRPArena
{
	int reserved;
};

/* message_callback helper, reference src/utils/message_callback.c:18-35 */
void srpMessage(SRPMessageType type, SRPMessageSeverity severity, const char* sourceFunction,
                const char* format, ...);
/* always-loud variant for conditions that would otherwise look like a silent fallback
 * (no GPU, unregistered program): callback if installed AND stderr */
void srpFatalMessage(const char* sourceFunction, const char* format, ...);

size_t srpSizeofType(SRPType type);

/* registry (srp_registry.c): device twins of a (vertex shader, fragment shader) combination */
typedef struct SRPProgramEntry
{
	SRPVertexShaderFunc vs;
	SRPFragmentShaderFunc fs;
	int vsDeviceId, fsDeviceId;
	size_t uniformSize;      /* the larger of what the two shaders read */
} SRPProgramEntry;
bool srpLookupProgram(SRPVertexShaderFunc vs, SRPFragmentShaderFunc fs, SRPProgramEntry* out);

/* framebuffer helpers (srp_framebuffer.c) */
SRPFramebufferImpl* srpFramebufferImpl(const SRPFramebuffer* fb);
/* before anything on the submission stream overwrites the device planes: order it behind an
 * asynchronous download that may still be reading them */
void srpFramebufferBeforeWrite(SRPFramebufferImpl* fb);
void srpFramebufferAfterDraw(SRPFramebufferImpl* const* fbs, size_t n, bool stencilEnabled, bool alreadyMirrored);
/* a draw call that enqueued nothing (count 0, out-of-range, culled away, unregistered program ...):
 * under the default policy the host-visible planes must still show a clear issued before it */
void srpFramebufferAfterSkippedDraw(const SRPFramebuffer* fb);

int srpMirrorPlanes(void);

/* PNG loader (srp_png.c): returns malloc'ed RGB8 or NULL (+ reason) */
uint8_t* srpLoadPngRgb(const char* path, int* width, int* height, const char** reason);

#endif
