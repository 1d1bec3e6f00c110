/* srp-b200 host layer -- context defaults and state setters.
 * Behaviour of reference src/core/context.c:15-180, including its two quirks that
 * existing programs may depend on: srpStencilTest() enables
 * the test whatever its argument is, and the *Separate setters with
 * SRP_FACE_FRONT_AND_BACK set both faces and then the back face once more. */
#include <stdlib.h>
#include "srp_internal.h"

static void setStencilFace(SRPStencilFaceState* s, SRPCompareOp func, uint8_t ref, uint8_t mask)
{
	s->func = func; s->ref = ref; s->mask = mask;
}
static void setStencilOps(SRPStencilFaceState* s, SRPStencilOp sfail, SRPStencilOp dfail, SRPStencilOp pass)
{
	s->sfailOp = sfail; s->dfailOp = dfail; s->passOp = pass;
}
/* which single face a *Separate call finally writes: FRONT -> front, anything else -> back */
static SRPStencilFaceState* separateTarget(SRPFace face)
{
	return (face == SRP_FACE_FRONT) ? &srpContext.stencil.front : &srpContext.stencil.back;
}

void srpNewContext(SRPContext* c)
{
	c->messageCallback.func = NULL;
	c->messageCallback.userParameter = NULL;
	c->provokingVertexMode = SRP_PROVOKING_VERTEX_LAST;

	c->raster.frontFace = SRP_WINDING_CCW;
	c->raster.cullFace = SRP_FACE_NONE;
	c->raster.polygonMode = SRP_POLYGON_MODE_FILL;
	c->raster.pointSize = 1.f;

	c->scissor.enabled = false;
	c->scissor.x = c->scissor.y = c->scissor.width = c->scissor.height = 0;

	c->depth.testEnable = false;
	c->depth.writeEnable = true;
	c->depth.compareOp = SRP_COMPARE_GREATER;

	c->stencil.enabled = false;
	SRPStencilFaceState* faces[2] = { &c->stencil.front, &c->stencil.back };
	for (int i = 0; i < 2; i++)
	{
		faces[i]->func = SRP_COMPARE_ALWAYS;
		faces[i]->ref = 0;
		faces[i]->mask = 0xFF;
		faces[i]->writeMask = 0xFF;
		faces[i]->sfailOp = faces[i]->dfailOp = faces[i]->passOp = SRP_STENCIL_KEEP;
	}

	/* the slot the reference uses for its bump arena carries the runtime handle here;
	 * the device runtime itself is created lazily by the first call that needs the GPU */
	c->arena = calloc(1, sizeof(struct SRPArena));
}

void srpSetMessageCallback(SRPMessageCallback callback) { srpContext.messageCallback = callback; }
void srpProvokingVertexMode(SRPProvokingVertexMode mode) { srpContext.provokingVertexMode = mode; }
void srpRasterCullFace(SRPFace face) { srpContext.raster.cullFace = face; }
void srpRasterFrontFace(SRPWinding face) { srpContext.raster.frontFace = face; }
void srpRasterPolygonMode(SRPPolygonMode mode) { srpContext.raster.polygonMode = mode; }
void srpRasterPointSize(float size) { srpContext.raster.pointSize = size; }

void srpScissorTest(bool enable) { srpContext.scissor.enabled = enable; }
void srpScissorOptions(size_t x, size_t y, size_t width, size_t height)
{
	srpContext.scissor.x = x; srpContext.scissor.y = y;
	srpContext.scissor.width = width; srpContext.scissor.height = height;
}

void srpStencilTest(bool enable)
{
	(void) enable;                       /* reference behaviour: always enables */
	srpContext.stencil.enabled = true;
}
void srpStencilFunc(SRPCompareOp func, uint8_t ref, uint8_t mask)
{
	setStencilFace(&srpContext.stencil.front, func, ref, mask);
	setStencilFace(&srpContext.stencil.back, func, ref, mask);
}
void srpStencilFuncSeparate(SRPFace face, SRPCompareOp func, uint8_t ref, uint8_t mask)
{
	if (face == SRP_FACE_NONE) return;
	if (face == SRP_FACE_FRONT_AND_BACK) srpStencilFunc(func, ref, mask);
	setStencilFace(separateTarget(face), func, ref, mask);
}
void srpStencilOp(SRPStencilOp sfail, SRPStencilOp dfail, SRPStencilOp pass)
{
	setStencilOps(&srpContext.stencil.front, sfail, dfail, pass);
	setStencilOps(&srpContext.stencil.back, sfail, dfail, pass);
}
void srpStencilOpSeparate(SRPFace face, SRPStencilOp sfail, SRPStencilOp dfail, SRPStencilOp pass)
{
	if (face == SRP_FACE_NONE) return;
	if (face == SRP_FACE_FRONT_AND_BACK) srpStencilOp(sfail, dfail, pass);
	setStencilOps(separateTarget(face), sfail, dfail, pass);
}
void srpStencilWriteMask(uint8_t mask)
{
	srpContext.stencil.front.writeMask = mask;
	srpContext.stencil.back.writeMask = mask;
}
void srpStencilWriteMaskSeparate(SRPFace face, uint8_t mask)
{
	if (face == SRP_FACE_NONE) return;
	if (face == SRP_FACE_FRONT_AND_BACK) srpStencilWriteMask(mask);
	separateTarget(face)->writeMask = mask;
}

void srpDepthTest(bool enable) { srpContext.depth.testEnable = enable; }
void srpDepthWrite(bool enable) { srpContext.depth.writeEnable = enable; }
void srpDepthCompareOp(SRPCompareOp op) { srpContext.depth.compareOp = op; }
