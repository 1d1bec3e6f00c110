/* srp-b200 host layer -- vertex / index buffer objects.
 * API behaviour of reference src/core/buffer.c:20-128 (grow-only storage, the element
 * count is nBytesData / element size, index types u8/u16/u32/u64), but the payload is
 * device-resident: *CopyData is an upload with the reference's memcpy semantics -- the caller
 * may reuse its memory when the call returns -- whatever the source is (pageable or pinned
 * host memory, or, through unified addressing, device memory, e.g. a buffer another GPU
 * broadcast over NVLink).  The copy is ordered behind the last draw that reads the buffer
 * (its `lastUse` event under the explicit policy), not behind everything that is queued, so
 * next frame's geometry crosses PCIe while this frame is still being rasterised. */
#include <stdlib.h>
#include "srp_internal.h"

static bool reserveDevice(void** data, size_t* allocated, size_t nBytes, const char* func)
{
	if (nBytes <= *allocated && *data != NULL)
		return true;
	srpcuFree(*data);
	*data = srpcuMalloc(nBytes);
	if (*data == NULL)
	{
		*allocated = 0;
		srpFatalMessage(func, "%s", srpcuLastError());
		return false;
	}
	*allocated = nBytes;
	return true;
}

SRPVertexBuffer* srpNewVertexBuffer(void)
{
	SRPVertexBuffer* vb = calloc(1, sizeof *vb);
	if (!vb) abort();
	return vb;
}

void srpVertexBufferCopyData(SRPVertexBuffer* vb, size_t nBytesPerVertex, size_t nBytesData, const void* data)
{
	if (!reserveDevice(&vb->data, &vb->nBytesAllocated, nBytesData, __func__))
	{
		vb->nVertices = 0;
		return;
	}
	vb->nBytesPerVertex = nBytesPerVertex;
	vb->nVertices = nBytesPerVertex ? nBytesData / nBytesPerVertex : 0;
	if (srpcuUpload(vb->data, data, nBytesData, vb->lastUse))
		srpFatalMessage(__func__, "%s", srpcuLastError());
}

void srpFreeVertexBuffer(SRPVertexBuffer* vb)
{
	if (!vb) return;
	srpcuFree(vb->data);
	srpcuFreeEvent(vb->lastUse);
	free(vb);
}

SRPIndexBuffer* srpNewIndexBuffer(void)
{
	SRPIndexBuffer* ib = calloc(1, sizeof *ib);
	if (!ib) abort();
	ib->indicesType = SRP_UINT8;
	ib->nBytesPerIndex = 1;
	return ib;
}

void srpIndexBufferCopyData(SRPIndexBuffer* ib, SRPType indicesType, size_t nBytesData, const void* data)
{
	const size_t elem = srpSizeofType(indicesType);
	if (!(indicesType == SRP_UINT8 || indicesType == SRP_UINT16 || indicesType == SRP_UINT32 || indicesType == SRP_UINT64))
	{
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__, "Unexpected type (%i)", indicesType);
		ib->nIndices = 0;
		return;
	}
	if (!reserveDevice(&ib->data, &ib->nBytesAllocated, nBytesData, __func__))
	{
		ib->nIndices = 0;
		return;
	}
	ib->indicesType = indicesType;
	ib->nBytesPerIndex = elem;
	ib->nIndices = nBytesData / elem;
	if (srpcuUpload(ib->data, data, nBytesData, ib->lastUse))
		srpFatalMessage(__func__, "%s", srpcuLastError());
}

void srpFreeIndexBuffer(SRPIndexBuffer* ib)
{
	if (!ib) return;
	srpcuFree(ib->data);
	srpcuFreeEvent(ib->lastUse);
	free(ib);
}
