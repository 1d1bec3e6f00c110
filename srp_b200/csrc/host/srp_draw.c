/* srp-b200 host layer -- the draw entry points: validation, dispatch, state snapshot.
 *
 * This is the drop-in seam.  It mirrors reference src/pipeline/draw.c:63-189 (drawBuffer,
 * drawTriangles/Lines/Points, checkOOB) and the warnings of primitive_assembly.c:180-196,
 * then -- instead of assembling and rasterising on the CPU -- snapshots the context, the
 * varyings layout and the uniform into one SrpdDraw and submits it to the CUDA layer.
 *
 * Messages (type, severity, wording) follow the reference so that programs inspecting the
 * callback keep working; conditions that only exist here (unregistered program, varyings
 * larger than the device limit, no GPU) are reported HIGH through the callback and on
 * stderr, and the draw is skipped.  Nothing is ever rendered on the CPU. */
#include <stdlib.h>
#include <string.h>
#include "srp_internal.h"

static size_t gRow0 = 0, gRow1 = SIZE_MAX;
static unsigned long long gDraws = 0, gSubDraws = 0;

void srpB200SetRowRange(size_t row0, size_t row1) { gRow0 = row0; gRow1 = row1; }
size_t srpB200TileWidth(void) { return (size_t) srpcuTileWidth(); }
size_t srpB200TileHeight(void) { return (size_t) srpcuTileHeight(); }
const char* srpB200Version(void) { return srpcuVersion(); }
void srpB200SetDevice(int device) { srpcuSetDevice(device); }
void* srpB200Stream(void) { return srpcuStream(); }
int srpB200SetLane(int lane)
{
	if (srpcuSetLane(lane))
	{
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__, "lane %d does not exist (0..%d)", lane, srpcuLaneCount() - 1);
		return 1;
	}
	return 0;
}
int srpB200GetLane(void) { return srpcuLane(); }
int srpB200LaneCount(void) { return srpcuLaneCount(); }

/* ---- multi-GPU plumbing: peer memory and stream-ordered flags (include/srp_b200.h) ---- */
void* srpB200DeviceAlloc(size_t bytes) { return srpcuMalloc(bytes); }
void srpB200DeviceFree(void* p) { srpcuFree(p); }
int srpB200IpcExport(const void* devicePtr, SRPB200IpcHandle* handle)
{
	if (srpcuIpcExport(devicePtr, handle->bytes))
	{
		srpFatalMessage(__func__, "%s", srpcuLastError());
		return 1;
	}
	return 0;
}
void* srpB200IpcOpen(const SRPB200IpcHandle* handle)
{
	void* p = srpcuIpcOpen(handle->bytes);
	if (!p)
		srpFatalMessage(__func__, "%s", srpcuLastError());
	return p;
}
void srpB200IpcClose(void* mappedPtr)
{
	if (srpcuIpcClose(mappedPtr))
		srpFatalMessage(__func__, "%s", srpcuLastError());
}
void srpB200StreamSignal(uint32_t* flag, uint32_t value)
{
	if (srpcuStreamSignal(flag, value))
		srpFatalMessage(__func__, "%s", srpcuLastError());
}
void srpB200StreamWait(const uint32_t* flag, uint32_t value)
{
	if (srpcuStreamWaitFlag(flag, value))
		srpFatalMessage(__func__, "%s", srpcuLastError());
}
void srpB200StreamWaitAll(const uint32_t* flags, uint32_t count, uint32_t value)
{
	if (srpcuStreamWaitFlags(flags, count, value))
		srpFatalMessage(__func__, "%s", srpcuLastError());
}

void srpB200GetStats(SRPB200Stats* out)
{
	SrpdStats s;
	unsigned long long launches, h2d, d2h;
	srpcuGetStats(&s, &launches, &h2d, &d2h);
	out->draws = gDraws;
	out->primsIn = s.primsIn;
	out->primsEmitted = s.primsEmitted;
	out->primsStored = s.primsStored;
	out->fragsEmitted = s.fragsEmitted;
	out->fragsShaded = s.fragsShaded;
	out->kernelLaunches = launches;
	out->h2dBytes = h2d;
	out->d2hBytes = d2h;
	out->overflow = s.overflow;
	out->subDraws = gSubDraws;
}
void srpB200ResetStats(void) { gDraws = 0; gSubDraws = 0; srpcuResetStats(); }
void srpB200SetProfiling(int enable) { srpcuSetProfiling(enable); }
unsigned long long srpB200CollectStageTimes(double outMs[3]) { return srpcuCollectStageTimes(outMs); }

/* ---- validation, reference draw.c:167-189 and primitive_assembly.c:180-196 ---- */
static bool checkOOB(const SRPIndexBuffer* ib, const SRPVertexBuffer* vb, size_t startIndex, size_t count)
{
	const size_t endIndex = startIndex + count - 1;
	const size_t bufferSize = ib ? ib->nIndices : vb->nVertices;
	if (endIndex >= bufferSize)
	{
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__,
			ib ? "Attempt to OOB access index buffer (read) at indices %zu-%zu (size: %zu)\n"
			   : "Attempt to OOB access vertex buffer (read) at indices %zu-%zu (size: %zu)\n",
			startIndex, endIndex, bufferSize);
		return true;
	}
	return false;
}

static void warnOnExcessVertexCount(SRPPrimitive prim, size_t vertexCount)
{
	if (prim == SRP_PRIM_LINES && vertexCount % 2 != 0)
		srpMessage(SRP_MESSAGE_WARNING, SRP_MESSAGE_SEVERITY_LOW, __func__,
			"Odd vertex count when drawing SRP_PRIM_LINES. The last vertex will be ignored\n");
	if (prim == SRP_PRIM_TRIANGLES && vertexCount % 3 != 0)
		srpMessage(SRP_MESSAGE_WARNING, SRP_MESSAGE_SEVERITY_LOW, __func__,
			"Vertex count not divisible by 3 when drawing SRP_PRIM_TRIANGLES. The last %i vertex/vertices will be ignored\n",
			(int) (vertexCount % 3));
}

/* primitive counts, reference topology.c:14-22,50-60 */
static size_t inputPrimitiveCount(SRPPrimitive prim, size_t n)
{
	switch (prim)
	{
		case SRP_PRIM_TRIANGLES:      return n / 3;
		case SRP_PRIM_TRIANGLE_STRIP:
		case SRP_PRIM_TRIANGLE_FAN:   return n >= 3 ? n - 2 : 0;
		case SRP_PRIM_LINES:          return n / 2;
		case SRP_PRIM_LINE_STRIP:     return n != 1 ? n - 1 : 0;
		case SRP_PRIM_LINE_LOOP:      return n != 1 ? n : 0;
		case SRP_PRIM_POINTS:         return n;
	}
	return 0;
}

static void snapshotStencilFace(SrpdStencilFace* d, const SRPStencilFaceState* s)
{
	d->func = (uint8_t) s->func; d->ref = s->ref; d->mask = s->mask; d->writeMask = s->writeMask;
	d->sfailOp = (uint8_t) s->sfailOp; d->dfailOp = (uint8_t) s->dfailOp; d->passOp = (uint8_t) s->passOp; d->pad = 0;
}

/* varyings layout: attribute after attribute, tightly packed (interpolation.c:104-161).
 * The slot reserved per blob covers both what the program declared and what its
 * attribute list spans (tests/scenes/interpolation/flat.c declares 3 floats over a
 * 1-byte struct). */
static bool snapshotVaryings(SrpdState* st, const SRPVertexShader* vs)
{
	if (vs->nVaryings > SRPD_MAX_VARYINGS)
	{
		srpFatalMessage("srpDraw", "%zu varyings exceed the device limit of %d", vs->nVaryings, SRPD_MAX_VARYINGS);
		return false;
	}
	size_t offset = 0;
	for (size_t i = 0; i < vs->nVaryings; i++)
	{
		const SRPVaryingInfo* info = &vs->varyingsInfo[i];
		size_t elem;
		switch (info->type)
		{
			case SRP_FLOAT: case SRP_INT32: case SRP_UINT32: elem = 4; break;
			case SRP_DOUBLE: case SRP_INT64: case SRP_UINT64: elem = 8; break;
			case SRP_INT16: case SRP_UINT16: elem = 2; break;
			case SRP_INT8: case SRP_UINT8: elem = 1; break;
			default:
				srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, "interpolateAttributes",
					"Unexpected type (%i)", info->type);
				elem = 0;
		}
		st->varyings[i].offset = (uint16_t) offset;
		st->varyings[i].nItems = (uint16_t) (elem ? info->nItems : 0);
		st->varyings[i].type = (uint8_t) info->type;
		st->varyings[i].mode = (uint8_t) info->interpolationMode;
		st->varyings[i].elemSize = (uint8_t) elem;
		st->varyings[i].pad = 0;
		offset += elem * info->nItems;
		if (offset > SRPD_MAX_VARYING_BYTES)
			break;
	}
	size_t slot = vs->varyingsSize > offset ? vs->varyingsSize : offset;
	slot = (slot + 7) & ~(size_t) 7;
	if (slot > SRPD_MAX_VARYING_BYTES)
	{
		srpFatalMessage("srpDraw", "varyings of %zu bytes exceed the device limit of %d bytes", slot, SRPD_MAX_VARYING_BYTES);
		return false;
	}
	st->nVaryings = (int32_t) vs->nVaryings;
	st->varyingsSize = (int32_t) vs->varyingsSize;
	st->slotSize = (int32_t) slot;

	/* all-float layouts (<= 16 floats) take the word-wise interpolation path of the tile kernel */
	st->allFloat = 1;
	st->floatModes = 0;
	size_t nFloats = 0;
	for (size_t i = 0; i < vs->nVaryings && st->allFloat; i++)
	{
		if (st->varyings[i].type != SRP_FLOAT || st->varyings[i].mode > SRP_INTERPOLATION_MODE_FLAT)
			st->allFloat = 0;
		else
			for (unsigned e = 0; e < st->varyings[i].nItems; e++, nFloats++)
				if (nFloats < 16)
					st->floatModes |= (uint32_t) st->varyings[i].mode << (2 * nFloats);
	}
	if (nFloats > 16 || nFloats * 4 > slot)
		st->allFloat = 0;
	st->nFloats = (uint8_t) (st->allFloat ? nFloats : 0);
	return true;
}

static bool buildDraw(
	SrpdDraw* d, const SRPIndexBuffer* ib, const SRPVertexBuffer* vb, const SRPFramebuffer* fb,
	const SRPShaderProgram* sp, SRPPrimitive primitive, size_t startIndex, size_t count,
	SRPProgramEntry* outProgram)
{
	memset(d, 0, sizeof *d);
	const SRPContext* c = &srpContext;

	const bool isTriangle = primitive == SRP_PRIM_TRIANGLES || primitive == SRP_PRIM_TRIANGLE_STRIP || primitive == SRP_PRIM_TRIANGLE_FAN;
	const bool isLine = primitive == SRP_PRIM_LINES || primitive == SRP_PRIM_LINE_STRIP || primitive == SRP_PRIM_LINE_LOOP;
	const bool isPoint = primitive == SRP_PRIM_POINTS;
	if (!isTriangle && !isLine && !isPoint)
	{
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, "drawBuffer", "Unknown primitive type: %i", primitive);
		return false;
	}
	if (isTriangle && c->raster.cullFace == SRP_FACE_FRONT_AND_BACK)
		return false;                                   /* draw.c:89 */
	if (isTriangle && c->raster.polygonMode != SRP_POLYGON_MODE_FILL && c->raster.polygonMode != SRP_POLYGON_MODE_LINE
	    && c->raster.polygonMode != SRP_POLYGON_MODE_POINT)
	{
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, "assembleTrianglesGeneric",
			"Unexpected srpContext.raster.polygonMode (%i)", c->raster.polygonMode);
		return false;
	}
	if (!isPoint)
		warnOnExcessVertexCount(primitive, count);
	if (isPoint && c->raster.pointSize <= 0.)
		return false;                                   /* primitive_assembly.c:227-228 */
	const size_t nPrims = inputPrimitiveCount(primitive, count);
	if (nPrims == 0)
		return false;
	if (nPrims > 100000000u)
	{
		srpFatalMessage("srpDraw", "%zu input primitives exceed the per-draw limit", nPrims);
		return false;
	}

	SRPProgramEntry prog;
	if (!srpLookupProgram(sp->vs->shader, sp->fs->shader, &prog))
	{
		srpFatalMessage("srpDraw",
			"shader program (vs %p, fs %p) has no registered __device__ twins (srpB200RegisterProgram / "
			"srpB200RegisterVertexShader / srpB200RegisterFragmentShader); nothing is drawn -- this library has no CPU path",
			(void*) sp->vs->shader, (void*) sp->fs->shader);
		return false;
	}
	*outProgram = prog;

	SrpdState* st = &d->st;
	st->width = (int32_t) fb->width;
	st->height = (int32_t) fb->height;
	st->frontFaceCW = c->raster.frontFace == SRP_WINDING_CW;
	st->cullFace = (uint8_t) c->raster.cullFace;
	st->polygonMode = isTriangle ? (uint8_t) c->raster.polygonMode : (uint8_t) SRP_POLYGON_MODE_FILL;
	st->provokingFirst = c->provokingVertexMode == SRP_PROVOKING_VERTEX_FIRST;
	st->pointSize = c->raster.pointSize;
	st->scissorEnabled = c->scissor.enabled;
	st->scissorX0 = c->scissor.x;
	st->scissorX1 = c->scissor.x + c->scissor.width;
	st->scissorY0 = c->scissor.y;
	st->scissorY1 = c->scissor.y + c->scissor.height;
	st->stencilEnabled = c->stencil.enabled;
	snapshotStencilFace(&st->stencilFront, &c->stencil.front);
	snapshotStencilFace(&st->stencilBack, &c->stencil.back);
	st->depthTest = c->depth.testEnable;
	st->depthWrite = c->depth.writeEnable;
	st->depthOp = (uint8_t) c->depth.compareOp;
	st->earlyDepth = !sp->fs->mayOverwriteDepth;
	st->vsProgramId = prog.vsDeviceId;
	st->fsProgramId = prog.fsDeviceId;
	if (!snapshotVaryings(st, sp->vs))
		return false;

	d->vb = vb->data;
	d->vbStride = vb->nBytesPerVertex;
	d->ib = ib ? ib->data : NULL;
	d->ibElemSize = ib ? (uint32_t) ib->nBytesPerIndex : 0;
	d->topology = (uint32_t) primitive;
	d->startIndex = startIndex;
	d->count = count;
	d->nInputPrims = (uint32_t) nPrims;
	/* worst-case records per input primitive: the clipper keeps at most 10 polygon vertices, i.e.
	 * a clipped triangle fans into <= 8 (geom.cu: SRPD_CLIP_MAX_VERTS); a line is stored as
	 * segments of 16 DDA fragments and has <= max(W, H) + 2 fragments */
	const uint32_t maxDim = (uint32_t) (fb->width > fb->height ? fb->width : fb->height);
	const uint32_t segmentsPerLine = (maxDim + 2) / 16 + 2;
	if (isTriangle)
	{
		d->kind = st->polygonMode == SRP_POLYGON_MODE_FILL ? SRPD_KIND_TRIANGLE
		        : st->polygonMode == SRP_POLYGON_MODE_LINE ? SRPD_KIND_LINE : SRPD_KIND_POINT;
		d->maxOutPerInput = st->polygonMode == SRP_POLYGON_MODE_FILL ? 8
		                  : st->polygonMode == SRP_POLYGON_MODE_LINE ? 24 * segmentsPerLine : 24;
	}
	else
	{
		d->kind = isLine ? SRPD_KIND_LINE : SRPD_KIND_POINT;
		d->maxOutPerInput = isLine ? segmentsPerLine : 1;
	}
	const size_t th = (size_t) srpcuTileHeight();
	const size_t tilesY = (fb->height + th - 1) / th;
	size_t r0 = gRow0 / th, r1 = gRow1 == SIZE_MAX ? tilesY : (gRow1 + th - 1) / th;
	if (r1 > tilesY) r1 = tilesY;
	if (r0 > r1) r0 = r1;
	d->tileRow0 = (uint32_t) r0;
	d->tileRow1 = (uint32_t) r1;
	st->stripY0 = (int32_t) (r0 * th);
	st->stripY1 = (int32_t) (r1 * th < fb->height ? r1 * th : fb->height);
	return true;
}

/* under the explicit policy the draw is still in flight when the call returns: remember where the
 * stream was, so that a later *CopyData into this buffer waits for exactly this draw */
static void markBufferUse(void** lastUse)
{
	if (*lastUse == NULL)
		*lastUse = srpcuNewEvent();
	if (*lastUse)
		srpcuRecordEvent(*lastUse);
}

static void submit(
	const SRPIndexBuffer* ib, const SRPVertexBuffer* vb, SRPFramebuffer* const* fbs, size_t nFrames,
	const SRPShaderProgram* sp, const void* uniforms, size_t uniformStride,
	SRPPrimitive primitive, size_t startIndex, size_t count, bool clearFirst)
{
	if (nFrames == 0)
		return;
	bool enqueued = false;
	SRPFramebufferImpl** impls = malloc(nFrames * sizeof *impls);
	SrpdFrame* frames = malloc(nFrames * sizeof *frames);
	if (!impls || !frames) abort();
	bool ok = true;
	for (size_t f = 0; f < nFrames; f++)
	{
		impls[f] = srpFramebufferImpl(fbs[f]);
		if (impls[f] == NULL || impls[f]->pub.width != fbs[0]->width || impls[f]->pub.height != fbs[0]->height)
		{
			srpFatalMessage("srpDraw", "framebuffer %zu is not a live srp framebuffer of the batch's size", f);
			ok = false;
			nFrames = f;      /* (only the ones seen so far are touched below) */
			break;
		}
		if (clearFirst)
			srpFramebufferClear(fbs[f]);
	}

	SrpdDraw d;
	SRPProgramEntry prog;
	if (ok && count != 0 && !checkOOB(ib, vb, startIndex, count)
	    && buildDraw(&d, ib, vb, fbs[0], sp, primitive, startIndex, count, &prog))
	{
		for (size_t f = 0; f < nFrames; f++)
			srpFramebufferBeforeWrite(impls[f]);
		d.nFrames = (uint32_t) nFrames;
		/* The scratch pools of a (sub-)draw hold its worst case, so nothing can overflow and nothing
		 * ever has to be repeated -- under either synchronisation policy.  A draw whose worst case
		 * exceeds the scratch budget is submitted as consecutive ranges of its input primitives;
		 * primitive ids continue across the ranges on the device. */
		const uint32_t total = d.nInputPrims;
		const uint32_t perSub = srpcuMaxPrimsPerSubDraw(&d);
		uint32_t chunk = 0;
		for (uint32_t first = 0; first < total; first += perSub, chunk++)
		{
			d.firstPrim = first;
			d.nInputPrims = total - first < perSub ? total - first : perSub;
			d.chunkIndex = chunk;
			const bool last = first + d.nInputPrims >= total;
			for (size_t f = 0; f < nFrames; f++)
			{
				frames[f].uniform = NULL;
				frames[f].color = impls[f]->dColor;
				frames[f].depth = impls[f]->dDepth;
				frames[f].stencil = impls[f]->dStencil;
				frames[f].clearPending = chunk == 0 && impls[f]->clearPending;
				frames[f].pad = impls[f]->ownsDevicePlanes ? 0u : 1u;      /* bit0: planes are caller-provided device memory (maybe a peer GPU's) */
			}
			const size_t ubytes = uniforms ? prog.uniformSize : 0;
			/* default policy: the (last sub-)draw itself may refresh the host mirror band by band,
			 * overlapping the copies with rasterisation (it reports back whether it did) */
			int mirrored = 0;
			if (last && nFrames == 1 && srpB200GetSyncMode() == SRP_B200_SYNC_DRAW)
			{
				const int planes = srpMirrorPlanes();
				const bool wantStencil = (planes & SRP_B200_MIRROR_STENCIL) && (impls[0]->stencilTouched || d.st.stencilEnabled);
				SrpcuMirror m = { (planes & SRP_B200_MIRROR_COLOR) ? impls[0]->pub.color : NULL,
				                  (planes & SRP_B200_MIRROR_DEPTH) ? impls[0]->pub.depth : NULL,
				                  wantStencil ? impls[0]->pub.stencil : NULL };
				srpcuSetMirrorForNextDraw(&m, &mirrored);
			}
			if (srpcuDraw(&d, frames, uniforms, ubytes, uniformStride))
			{
				srpFatalMessage("srpDraw", "%s", srpcuLastError());
				break;
			}
			gDraws += chunk == 0;
			gSubDraws++;
			enqueued = true;
			if (chunk == 0)
				for (size_t f = 0; f < nFrames; f++)
					impls[f]->clearPending = false;      /* consumed by the first sub-draw's tiles */
			if (last)
			{
				if (srpB200GetSyncMode() != SRP_B200_SYNC_DRAW)
				{
					markBufferUse(&((SRPVertexBuffer*) vb)->lastUse);
					if (ib) markBufferUse(&((SRPIndexBuffer*) ib)->lastUse);
				}
				srpFramebufferAfterDraw(impls, nFrames, d.st.stencilEnabled, mirrored != 0);
			}
		}
	}
	if (!enqueued)
		for (size_t f = 0; f < nFrames; f++)
			srpFramebufferAfterSkippedDraw(fbs[f]);
	free(frames);
	free(impls);
}

void srpDrawVertexBuffer(const SRPVertexBuffer* vb, const SRPFramebuffer* fb, const SRPShaderProgram* sp,
                         SRPPrimitive primitive, size_t startIndex, size_t count)
{
	SRPFramebuffer* target = (SRPFramebuffer*) fb;
	submit(NULL, vb, &target, 1, sp, sp->uniform, 0, primitive, startIndex, count, false);
}

void srpDrawIndexBuffer(const SRPIndexBuffer* ib, const SRPVertexBuffer* vb, const SRPFramebuffer* fb,
                        const SRPShaderProgram* sp, SRPPrimitive primitive, size_t startIndex, size_t count)
{
	SRPFramebuffer* target = (SRPFramebuffer*) fb;
	submit(ib, vb, &target, 1, sp, sp->uniform, 0, primitive, startIndex, count, false);
}

void srpB200DrawBatch(const SRPIndexBuffer* ib, const SRPVertexBuffer* vb,
                      SRPFramebuffer* const* fbs, size_t nFrames,
                      const SRPShaderProgram* sp, const void* uniforms, size_t uniformStride,
                      SRPPrimitive primitive, size_t startIndex, size_t count, int clearFirst)
{
	submit(ib, vb, fbs, nFrames, sp, uniforms, uniformStride, primitive, startIndex, count, clearFirst != 0);
}
