/* srp-b200 host layer -- framebuffer objects and the host/device synchronisation policy.
 *
 * Public behaviour of reference src/core/framebuffer.c:17-62: three row-major planes
 * (u32 RGBA8888 with R in the top byte, f32 depth, u8 stencil); srpFramebufferClear sets
 * colour 0 and depth -1 and leaves stencil alone; the stencil plane starts as zeros (the
 * reference never initialises it and relies on fresh pages).
 *
 * Here the planes live in HBM.  The pointers in the public struct address a pinned host
 * mirror that is refreshed by a device-to-host copy: after every draw (default policy)
 * or on request (include/srp_b200.h).  A clear only sets a flag; the next draw's tile
 * kernel starts from the clear values instead of loading the planes, and writes every
 * tile, so clear + draw costs one write of the planes and no read.  If the next draw call
 * turns out to draw nothing, or the planes are requested first, the clear is materialised
 * then (srpFramebufferAfterSkippedDraw, materializeClear). */
#include <stdlib.h>
#include <string.h>
#include "srp_internal.h"

static SRPB200SyncMode gSyncMode = SRP_B200_SYNC_DRAW;
static int gMirrorPlanes = SRP_B200_MIRROR_ALL;

void srpB200SetMirrorPlanes(int planeMask) { gMirrorPlanes = planeMask & SRP_B200_MIRROR_ALL; }
int srpMirrorPlanes(void) { return gMirrorPlanes; }

void srpB200SetSyncMode(SRPB200SyncMode mode) { gSyncMode = mode; }
SRPB200SyncMode srpB200GetSyncMode(void) { return gSyncMode; }

SRPFramebufferImpl* srpFramebufferImpl(const SRPFramebuffer* fb)
{
	SRPFramebufferImpl* impl = (SRPFramebufferImpl*) fb;
	if (impl == NULL || impl->magic != SRP_FB_MAGIC)
		return NULL;
	return impl;
}

static SRPFramebuffer* newFramebuffer(size_t width, size_t height, void* dColor, void* dDepth, void* dStencil)
{
	if (width == 0 || height == 0 || width > 16384 || height > 16384)
	{
		srpFatalMessage("srpNewFramebuffer", "unsupported framebuffer size %zux%zu (1..16384 per side)", width, height);
		return NULL;
	}
	SRPFramebufferImpl* fb = calloc(1, sizeof *fb);
	if (!fb) abort();
	fb->magic = SRP_FB_MAGIC;
	fb->lane = -1;      /* nothing enqueued yet (the planes are zero-filled synchronously) */
	fb->pub.width = width;
	fb->pub.height = height;
	fb->pub.size = width * height;
	const size_t n = fb->pub.size;
	fb->ownsDevicePlanes = dColor == NULL;
	fb->dColor = dColor ? dColor : srpcuMalloc(n * sizeof(uint32_t));
	fb->dDepth = dDepth ? dDepth : srpcuMalloc(n * sizeof(float));
	fb->dStencil = dStencil ? dStencil : srpcuMalloc(n * sizeof(uint8_t));
	fb->pub.color = srpcuMallocHost(n * sizeof(uint32_t));
	fb->pub.depth = srpcuMallocHost(n * sizeof(float));
	fb->pub.stencil = srpcuMallocHost(n * sizeof(uint8_t));
	if (!fb->dColor || !fb->dDepth || !fb->dStencil || !fb->pub.color || !fb->pub.depth || !fb->pub.stencil)
	{
		srpFatalMessage("srpNewFramebuffer", "%s", srpcuLastError());
		srpFreeFramebuffer(&fb->pub);
		return NULL;
	}
	return &fb->pub;
}

SRPFramebuffer* srpNewFramebuffer(size_t width, size_t height)
{
	return newFramebuffer(width, height, NULL, NULL, NULL);
}

SRPFramebuffer* srpB200NewFramebufferOnDevice(size_t width, size_t height,
                                              void* deviceColor, void* deviceDepth, void* deviceStencil)
{
	if (!deviceColor || !deviceDepth || !deviceStencil)
	{
		srpFatalMessage(__func__, "all three device planes must be provided");
		return NULL;
	}
	return newFramebuffer(width, height, deviceColor, deviceDepth, deviceStencil);
}

void srpFreeFramebuffer(SRPFramebuffer* pub)
{
	SRPFramebufferImpl* fb = srpFramebufferImpl(pub);
	if (!fb) return;
	if (fb->downloadEvent)
	{
		srpcuHostWaitEvent(fb->downloadEvent);
		srpcuFreeEvent(fb->downloadEvent);
	}
	if (fb->ownsDevicePlanes)
	{
		srpcuFree(fb->dColor);
		srpcuFree(fb->dDepth);
		srpcuFree(fb->dStencil);
	}
	srpcuFreeHost(fb->pub.color);
	srpcuFreeHost(fb->pub.depth);
	srpcuFreeHost(fb->pub.stencil);
	fb->magic = 0;
	free(fb);
}

void srpFramebufferClear(const SRPFramebuffer* pub)
{
	SRPFramebufferImpl* fb = srpFramebufferImpl(pub);
	if (!fb) return;
	fb->clearPending = true;
	fb->mirrorStale = true;
}

void* srpB200FramebufferDevicePlane(const SRPFramebuffer* pub, int which)
{
	SRPFramebufferImpl* fb = srpFramebufferImpl(pub);
	if (!fb) return NULL;
	return which == 0 ? (void*) fb->dColor : which == 1 ? (void*) fb->dDepth : which == 2 ? (void*) fb->dStencil : NULL;
}

/* make a deferred clear real (needed when somebody wants the planes without a draw) */
static void materializeClear(SRPFramebufferImpl* fb)
{
	if (!fb->clearPending)
		return;
	srpFramebufferBeforeWrite(fb);
	if (srpcuClearPlanes(fb->dColor, fb->dDepth, fb->pub.size))
		srpFatalMessage("srpFramebufferClear", "%s", srpcuLastError());
	fb->clearPending = false;
}

/* A draw call that enqueued nothing (srp_draw.c) after srpFramebufferClear: under the default
 * policy the caller may read fb->color / fb->depth right after the call and must find the clear
 * there, as with the reference's immediate memset (core/framebuffer.c:57-62).  The device planes
 * are cleared by a kernel; the host mirror is written directly (no PCIe round trip). */
void srpFramebufferAfterSkippedDraw(const SRPFramebuffer* pub)
{
	SRPFramebufferImpl* fb = srpFramebufferImpl(pub);
	if (!fb || !fb->clearPending || gSyncMode != SRP_B200_SYNC_DRAW)
		return;
	materializeClear(fb);
	const size_t n = fb->pub.size;
	if (gMirrorPlanes & SRP_B200_MIRROR_COLOR)
		memset(fb->pub.color, 0, n * sizeof(uint32_t));
	if (gMirrorPlanes & SRP_B200_MIRROR_DEPTH)
		for (size_t i = 0; i < n; i++)
			fb->pub.depth[i] = -1.0f;
	/* (mirrorStale stays set: only srp_io.c looks at it, and a redundant download is harmless) */
}

static void enterLane(SRPFramebufferImpl* fb);

static void enqueueDownload(SRPFramebufferImpl* fb, int planes)
{
	enterLane(fb);
	materializeClear(fb);
	const size_t n = fb->pub.size;
	int err = 0;
	if (planes & SRP_B200_MIRROR_COLOR)
		err |= srpcuDownload(fb->pub.color, fb->dColor, n * sizeof(uint32_t));
	if (planes & SRP_B200_MIRROR_DEPTH)
		err |= srpcuDownload(fb->pub.depth, fb->dDepth, n * sizeof(float));
	if ((planes & SRP_B200_MIRROR_STENCIL) && fb->stencilTouched)
		err |= srpcuDownload(fb->pub.stencil, fb->dStencil, n * sizeof(uint8_t));
	if (err)
		srpFatalMessage("srpB200FramebufferDownload", "%s", srpcuLastError());
	if (planes & SRP_B200_MIRROR_STENCIL)
		fb->stencilTouched = false;
	fb->mirrorStale = planes != SRP_B200_MIRROR_ALL;
}

void srpB200FramebufferDownload(const SRPFramebuffer* pub)
{
	SRPFramebufferImpl* fb = srpFramebufferImpl(pub);
	if (!fb) return;
	fb->stencilTouched = true;      /* explicit request: bring everything */
	enqueueDownload(fb, SRP_B200_MIRROR_ALL);
	if (srpcuSynchronize())
		srpFatalMessage(__func__, "%s", srpcuLastError());
}

/* Asynchronous counterpart of srpB200FramebufferDownload for the explicit policy: the planes
 * selected by srpB200SetMirrorPlanes are copied on the copy stream as soon as the draws
 * enqueued so far have finished, and the call returns at once.  Draws into OTHER framebuffers
 * issued afterwards overlap with the copies (a program keeps two framebuffers and alternates);
 * a draw into THIS framebuffer is ordered behind them.  srpB200FramebufferWait() blocks until
 * the mirror is complete. */
void srpB200FramebufferDownloadAsync(const SRPFramebuffer* pub)
{
	SRPFramebufferImpl* fb = srpFramebufferImpl(pub);
	if (!fb) return;
	if (!fb->downloadEvent)
		fb->downloadEvent = srpcuNewEvent();
	if (!fb->downloadEvent)
	{
		srpFatalMessage(__func__, "%s", srpcuLastError());
		return;
	}
	enterLane(fb);
	materializeClear(fb);
	const int planes = gMirrorPlanes;
	const SrpcuMirror m = { (planes & SRP_B200_MIRROR_COLOR) ? fb->pub.color : NULL,
	                        (planes & SRP_B200_MIRROR_DEPTH) ? fb->pub.depth : NULL,
	                        ((planes & SRP_B200_MIRROR_STENCIL) && fb->stencilTouched) ? fb->pub.stencil : NULL };
	if (srpcuDownloadPlanesAsync(&m, fb->dColor, fb->dDepth, fb->dStencil, fb->pub.size, fb->downloadEvent))
	{
		srpFatalMessage(__func__, "%s", srpcuLastError());
		return;
	}
	if (planes & SRP_B200_MIRROR_STENCIL)
		fb->stencilTouched = false;
	fb->mirrorStale = planes != SRP_B200_MIRROR_ALL;
	fb->downloadInFlight = true;
}

void srpB200FramebufferWait(const SRPFramebuffer* pub)
{
	SRPFramebufferImpl* fb = srpFramebufferImpl(pub);
	if (!fb || !fb->downloadInFlight) return;
	if (srpcuHostWaitEvent(fb->downloadEvent))
		srpFatalMessage(__func__, "%s", srpcuLastError());
	fb->downloadInFlight = false;
}

/* order everything enqueued from now on behind this framebuffer's in-flight asynchronous download
 * (the strips' root releases a ring slot to its peers only once the slot's frame has left it) */
void srpB200FramebufferFence(const SRPFramebuffer* pub)
{
	SRPFramebufferImpl* fb = srpFramebufferImpl(pub);
	if (fb)
		srpFramebufferBeforeWrite(fb);
}

/* A framebuffer belongs to the lane that touched it last; when another lane takes it over, that
 * lane is ordered behind everything the previous one has enqueued (srp_b200.h: srpB200SetLane) */
static void enterLane(SRPFramebufferImpl* fb)
{
	const int lane = srpcuLane();
	if (fb->lane != lane)
	{
		if (srpcuOrderBehindLane(fb->lane))
			srpFatalMessage("srpB200SetLane", "%s", srpcuLastError());
		fb->lane = lane;
	}
}

void srpFramebufferBeforeWrite(SRPFramebufferImpl* fb)
{
	enterLane(fb);
	if (fb->downloadInFlight && srpcuStreamWaitEvent(fb->downloadEvent))
		srpFatalMessage("srpDraw", "%s", srpcuLastError());
}

void srpB200FramebufferUpload(const SRPFramebuffer* pub)
{
	SRPFramebufferImpl* fb = srpFramebufferImpl(pub);
	if (!fb) return;
	srpFramebufferBeforeWrite(fb);
	const size_t n = fb->pub.size;
	fb->clearPending = false;
	int err = srpcuUploadInStream(fb->dColor, fb->pub.color, n * sizeof(uint32_t));
	err |= srpcuUploadInStream(fb->dDepth, fb->pub.depth, n * sizeof(float));
	err |= srpcuUploadInStream(fb->dStencil, fb->pub.stencil, n * sizeof(uint8_t));
	if (err || srpcuSynchronize())
		srpFatalMessage(__func__, "%s", srpcuLastError());
}

void srpB200Finish(void)
{
	if (srpcuSynchronize())
		srpFatalMessage(__func__, "%s", srpcuLastError());
}

/* called by the draw entry points once the draw's kernels are enqueued */
void srpFramebufferAfterDraw(SRPFramebufferImpl* const* fbs, size_t n, bool stencilEnabled, bool alreadyMirrored)
{
	for (size_t i = 0; i < n; i++)
	{
		fbs[i]->mirrorStale = true;
		if (stencilEnabled)
			fbs[i]->stencilTouched = true;
	}
	if (gSyncMode != SRP_B200_SYNC_DRAW)
		return;
	if (alreadyMirrored)
	{
		/* the draw copied its bands to the host itself (runtime.cu); only bookkeeping is left */
		if (gMirrorPlanes & SRP_B200_MIRROR_STENCIL)
			fbs[0]->stencilTouched = false;
		fbs[0]->mirrorStale = gMirrorPlanes != SRP_B200_MIRROR_ALL;
	}
	else
		for (size_t i = 0; i < n; i++)
			enqueueDownload(fbs[i], gMirrorPlanes);
	if (srpcuSynchronize())
		srpFatalMessage("srpDraw", "%s", srpcuLastError());
}
