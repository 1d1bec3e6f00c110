/* srp-b200 internal -- plain-C descriptors shared by the host C layer (csrc/host) and
 * the CUDA layer (csrc/device).  Everything a draw needs is snapshotted BY VALUE into
 * one SrpdDraw at srpDraw*Buffer time (the reference reads the global srpContext
 * lazily while it rasterises, src/raster/fragment.c:63-245; user programs mutate the
 * context and the uniform between draws, e.g. tests/scenes/misc/stencil_test.c:126-135). */
#ifndef SRPD_DRAW_TYPES_H_
#define SRPD_DRAW_TYPES_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRPD_MAX_VARYINGS       16    /* attributes per program                          */
#define SRPD_MAX_VARYING_BYTES  64    /* bytes of one varyings blob (rounded up to 8)    */
#define SRPD_MAX_FRAMES_INLINE  1

/* primitive class a draw rasterises (after polygon-mode expansion) */
enum { SRPD_KIND_TRIANGLE = 0, SRPD_KIND_LINE = 1, SRPD_KIND_POINT = 2 };
/* topology of the input stream, same numbering as SRPPrimitive (include/srp/api.h) */
enum { SRPD_TOPO_POINTS = 0, SRPD_TOPO_LINES, SRPD_TOPO_LINE_STRIP, SRPD_TOPO_LINE_LOOP,
       SRPD_TOPO_TRIANGLES, SRPD_TOPO_TRIANGLE_STRIP, SRPD_TOPO_TRIANGLE_FAN };

typedef struct SrpdVarying
{
	uint16_t offset;       /* byte offset inside the blob                         */
	uint16_t nItems;
	uint8_t  type;         /* SRPType                                             */
	uint8_t  mode;         /* SRPInterpolationMode                                */
	uint8_t  elemSize;
	uint8_t  pad;
} SrpdVarying;

typedef struct SrpdStencilFace
{
	uint8_t func, ref, mask, writeMask, sfailOp, dfailOp, passOp, pad;
} SrpdStencilFace;

/* pipeline state snapshot */
typedef struct SrpdState
{
	int32_t width, height;           /* framebuffer, pixels                              */
	int32_t stripY0, stripY1;        /* pixel rows [stripY0, stripY1) this process rasterises (sort-first strips;
	                                    the whole frame otherwise): primitives outside keep their id, store nothing */
	/* raster */
	uint8_t frontFaceCW, cullFace, polygonMode, provokingFirst;
	float   pointSize;
	/* scissor: half-open [x0,x1) x [y0,y1) in size_t arithmetic (wraps like the reference) */
	uint8_t scissorEnabled, stencilEnabled, depthTest, depthWrite;
	uint8_t depthOp, earlyDepth, pad0, pad1;
	uint64_t scissorX0, scissorX1, scissorY0, scissorY1;
	SrpdStencilFace stencilFront, stencilBack;
	/* varyings */
	int32_t nVaryings;
	int32_t varyingsSize;            /* what the program declared                        */
	int32_t slotSize;                /* bytes reserved per blob: max(declared, sum of attributes), rounded up to 8 */
	int16_t vsProgramId, fsProgramId; /* indices into the device shader tables (vertex / fragment) */
	/* fast path: every attribute is SRP_FLOAT and 4-byte aligned (the usual case): the blob
	 * is nFloats consecutive floats, floatModes holds 2 bits of SRPInterpolationMode each */
	uint32_t floatModes;
	uint8_t  allFloat, nFloats, pad2, pad3;
	SrpdVarying varyings[SRPD_MAX_VARYINGS];
} SrpdState;

/* per-frame bindings (frame-parallel batches bind many; a plain draw binds one) */
typedef struct SrpdFrame
{
	const void* uniform;             /* device copy of the uniform block (may be NULL)   */
	uint32_t*   color;
	float*      depth;
	uint8_t*    stencil;
	uint32_t    clearPending;        /* 1: planes logically hold colour 0 / depth -1     */
	uint32_t    pad;                 /* bit0: the planes are caller-provided device memory (srpB200NewFramebufferOnDevice) */
} SrpdFrame;

typedef struct SrpdDraw
{
	SrpdState st;
	/* geometry source (device pointers) */
	const uint8_t* vb;
	uint64_t vbStride;
	const void* ib;                  /* NULL: non-indexed                                */
	uint32_t ibElemSize;             /* 1, 2, 4 or 8                                     */
	uint32_t topology;               /* SRPD_TOPO_*                                      */
	uint64_t startIndex;
	uint64_t count;                  /* stream indices                                   */
	uint32_t nInputPrims;            /* input primitives of THIS sub-draw                */
	uint32_t kind;                   /* SRPD_KIND_* of the emitted primitives            */
	uint32_t maxOutPerInput;         /* worst-case records one input primitive stores    */
	uint32_t nFrames;
	/* A draw whose worst-case record count exceeds the scratch budget is submitted as several
	 * sub-draws over consecutive ranges of its input primitives (srp_draw.c): firstPrim is the
	 * range's first input primitive, chunkIndex > 0 continues the primitive ids of the
	 * previous sub-draw (kept on the device), so no scratch pool can ever overflow. */
	uint32_t firstPrim;
	uint32_t chunkIndex;
	/* sort-first strips: only tile rows [tileRow0, tileRow1) are rasterised          */
	uint32_t tileRow0, tileRow1;
} SrpdDraw;

/* counters the kernels accumulate (all four are deterministic) */
typedef struct SrpdStats
{
	unsigned long long primsIn;      /* input primitives                                 */
	unsigned long long primsEmitted; /* primitive ids handed out (post-clip, post-cull)  */
	unsigned long long primsStored;  /* of those, how many can produce fragments         */
	unsigned long long fragsEmitted; /* coverage-passing fragments (= emitFragment calls)*/
	unsigned long long fragsShaded;  /* fragment shader invocations                      */
	unsigned long long overflow;     /* != 0: a scratch pool was too small, draw dropped */
} SrpdStats;

#ifdef __cplusplus
}
#endif
#endif
