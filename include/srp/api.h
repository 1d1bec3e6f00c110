/* srp-b200 -- public C API of the srp draw path (drop-in contract).
 *
 * This single header carries every type and entry point that kitrofimov/srp exposes
 * through include/srp/{context,buffer,framebuffer,texture,shaders,vertex,type,color,
 * message_callback,arena}.h; the per-topic headers of the same names in this
 * directory simply forward here, so `#include <srp/srp.h>` (and any of the topic
 * headers) keeps working for existing programs.  Names, enumerator order, struct
 * member order and therefore the x86-64 layouts are the reference's -- they are the
 * ABI that already-written programs and shaders were compiled against:
 *
 *   sizeof: SRPVertexShaderIn 24, SRPVertexShaderOut 24, SRPFragmentShaderIn 48,
 *           SRPFragmentShaderOut 20, SRPVertexShader 32, SRPFragmentShader 16,
 *           SRPShaderProgram 24, SRPVaryingInfo 16, SRPFramebuffer 48, SRPContext 144
 *   (checked by static assertions in srp_b200/csrc/host/abi_check.c)
 *
 * Behind these entry points the reference's src/pipeline, src/raster and src/memory
 * are replaced by device-resident buffers and sm_100a kernels; see DESIGN.md and
 * include/srp_b200.h for the small additive extension surface (shader registration,
 * synchronisation policy, batched draws).
 *
 * The header is valid C (C99..C23) and C++/CUDA. */
#ifndef SRP_API_H_
#define SRP_API_H_

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Functions that shaders call are also callable from device code when this header is
 * compiled by nvcc (the definitions live in the library's relocatable device code). */
#if defined(__CUDACC__)
	#define SRP_SHADER_CALLABLE __host__ __device__
#else
	#define SRP_SHADER_CALLABLE
#endif

/* ------------------------------------------------------------------ scalar types
 * reference: include/srp/type.h:17-29 */
typedef enum SRPType
{
	SRP_FLOAT = 0, SRP_DOUBLE,
	SRP_INT8, SRP_INT16, SRP_INT32, SRP_INT64,
	SRP_UINT8, SRP_UINT16, SRP_UINT32, SRP_UINT64
} SRPType;

/* ------------------------------------------------------------------ messages
 * reference: include/srp/message_callback.h:14-45.  Validation problems found at
 * draw time are reported through the callback (if one is installed) and the draw
 * is skipped; there are no return codes. */
typedef enum SRPMessageType { SRP_MESSAGE_ERROR, SRP_MESSAGE_WARNING } SRPMessageType;
typedef enum SRPMessageSeverity
{
	SRP_MESSAGE_SEVERITY_LOW, SRP_MESSAGE_SEVERITY_MODERATE, SRP_MESSAGE_SEVERITY_HIGH
} SRPMessageSeverity;
typedef void (*SRPMessageCallbackFunc)(
	SRPMessageType type, SRPMessageSeverity severity, const char* sourceFunction,
	const char* message, void* userParameter);
typedef struct SRPMessageCallback
{
	SRPMessageCallbackFunc func;
	void* userParameter;
} SRPMessageCallback;

/* ------------------------------------------------------------------ colour
 * reference: include/srp/color.h:17-20 */
typedef struct SRPColor { uint8_t r, g, b, a; } SRPColor;

/* ------------------------------------------------------------------ pipeline state
 * reference: include/srp/context.h:20-143 */
typedef enum SRPProvokingVertexMode
{
	SRP_PROVOKING_VERTEX_FIRST, SRP_PROVOKING_VERTEX_LAST   /* LAST is the default */
} SRPProvokingVertexMode;
typedef enum SRPWinding { SRP_WINDING_CCW, SRP_WINDING_CW } SRPWinding;
typedef enum SRPFace
{
	SRP_FACE_NONE, SRP_FACE_FRONT, SRP_FACE_BACK, SRP_FACE_FRONT_AND_BACK
} SRPFace;
typedef enum SRPPolygonMode
{
	SRP_POLYGON_MODE_FILL, SRP_POLYGON_MODE_LINE, SRP_POLYGON_MODE_POINT
} SRPPolygonMode;
typedef enum SRPCompareOp
{
	SRP_COMPARE_NEVER, SRP_COMPARE_ALWAYS, SRP_COMPARE_LESS, SRP_COMPARE_LEQUAL,
	SRP_COMPARE_GREATER, SRP_COMPARE_GEQUAL, SRP_COMPARE_EQUAL, SRP_COMPARE_NOTEQUAL
} SRPCompareOp;
typedef enum
{
	SRP_STENCIL_KEEP, SRP_STENCIL_ZERO, SRP_STENCIL_REPLACE,
	SRP_STENCIL_INCR, SRP_STENCIL_INCR_WRAP,   /* saturating / wrapping +1 */
	SRP_STENCIL_DECR, SRP_STENCIL_DECR_WRAP,   /* saturating / wrapping -1 */
	SRP_STENCIL_INVERT
} SRPStencilOp;

typedef struct SRPRasterState
{
	SRPWinding frontFace;
	SRPFace cullFace;
	SRPPolygonMode polygonMode;
	float pointSize;               /* pixels */
} SRPRasterState;

typedef struct SRPScissorState
{
	bool enabled;
	size_t x, y;                   /* upper-left corner, y measured from the top */
	size_t width, height;
} SRPScissorState;

typedef struct SRPStencilFaceState
{
	SRPCompareOp func;             /* (ref & mask) func (stored & mask) */
	uint8_t ref, mask, writeMask;
	SRPStencilOp sfailOp, dfailOp, passOp;
} SRPStencilFaceState;

typedef struct SRPStencilState
{
	bool enabled;
	SRPStencilFaceState front, back;
} SRPStencilState;

typedef struct SRPDepthState
{
	bool testEnable;
	bool writeEnable;              /* only honoured while the test is enabled */
	SRPCompareOp compareOp;        /* default GREATER: larger z is nearer */
} SRPDepthState;

/* Opaque per-context runtime object.  The reference keeps its bump arena here
 * (include/srp/arena.h); this build keeps the device runtime (stream, scratch
 * pools) behind the same pointer. */
typedef struct SRPArena SRPArena;

typedef struct SRPContext
{
	SRPMessageCallback messageCallback;
	SRPProvokingVertexMode provokingVertexMode;
	SRPRasterState raster;
	SRPScissorState scissor;
	SRPStencilState stencil;
	SRPDepthState depth;
	SRPArena* arena;
} SRPContext;

/* The user program defines this object (`SRPContext srpContext;`) and initialises it
 * with srpNewContext(&srpContext); all state setters below write to it, and every
 * draw call snapshots it.  reference: include/srp/context.h:194 */
extern SRPContext srpContext;

void srpNewContext(SRPContext* pContext);
void srpSetMessageCallback(SRPMessageCallback callback);
void srpProvokingVertexMode(SRPProvokingVertexMode mode);
void srpRasterCullFace(SRPFace face);
void srpRasterFrontFace(SRPWinding face);
void srpRasterPolygonMode(SRPPolygonMode mode);
void srpRasterPointSize(float size);
void srpScissorTest(bool enable);
void srpScissorOptions(size_t x, size_t y, size_t width, size_t height);
void srpStencilTest(bool enable);   /* NB: like the reference (core/context.c:96-99) this ENABLES regardless of the argument */
void srpStencilFunc(SRPCompareOp func, uint8_t ref, uint8_t mask);
void srpStencilFuncSeparate(SRPFace face, SRPCompareOp func, uint8_t ref, uint8_t mask);
void srpStencilOp(SRPStencilOp sfail, SRPStencilOp dfail, SRPStencilOp pass);
void srpStencilOpSeparate(SRPFace face, SRPStencilOp sfail, SRPStencilOp dfail, SRPStencilOp pass);
void srpStencilWriteMask(uint8_t mask);
void srpStencilWriteMaskSeparate(SRPFace face, uint8_t mask);
void srpDepthTest(bool enable);
void srpDepthWrite(bool enable);
void srpDepthCompareOp(SRPCompareOp op);

/* ------------------------------------------------------------------ vertices & varyings
 * reference: include/srp/vertex.h:24-46 */
typedef struct SRPVertex SRPVertex;              /* user-defined vertex record  */
typedef struct SRPVarying SRPVarying;            /* user-defined VS output blob */
typedef struct SRPInterpolated SRPInterpolated;  /* same blob after interpolation */

typedef enum SRPInterpolationMode
{
	SRP_INTERPOLATION_MODE_PERSPECTIVE,
	SRP_INTERPOLATION_MODE_AFFINE,
	SRP_INTERPOLATION_MODE_FLAT      /* value of the provoking vertex */
} SRPInterpolationMode;

/* One entry per varying; the blob is laid out attribute after attribute, each
 * nItems * sizeof(type) bytes, tightly packed. */
typedef struct
{
	size_t nItems;
	SRPType type;
	SRPInterpolationMode interpolationMode;
} SRPVaryingInfo;

/* ------------------------------------------------------------------ shaders
 * reference: include/srp/shaders.h:19-97 */
typedef struct SRPUniform SRPUniform;            /* opaque, un-sized, may be NULL */

typedef struct SRPVertexShaderIn
{
	SRPUniform* uniform;
	SRPVertex* vertex;
	size_t vertexID;                 /* the vertex index (not the stream index) */
} SRPVertexShaderIn;

typedef struct SRPVertexShaderOut
{
	union {
		float clipPosition[4];
		float ndcPosition[4];
	};
	SRPVarying* varyings;
} SRPVertexShaderOut;

typedef struct SRPVertexShader
{
	void (*shader)(SRPVertexShaderIn* in, SRPVertexShaderOut* out);
	size_t nVaryings;
	SRPVaryingInfo* varyingsInfo;
	size_t varyingsSize;             /* bytes of one varyings blob */
} SRPVertexShader;

typedef struct SRPFragmentShaderIn
{
	SRPUniform* uniform;
	SRPInterpolated* varyings;
	float fragCoord[4];              /* x+.5, y+.5, depth, 1/(interpolated 1/w) */
	bool frontFacing;
	size_t primitiveID;              /* counts emitted (post-clip, post-cull) primitives */
} SRPFragmentShaderIn;

typedef struct SRPFragmentShaderOut
{
	float color[4];
	float fragDepth;                 /* honoured only with mayOverwriteDepth */
} SRPFragmentShaderOut;

typedef struct SRPFragmentShader
{
	void (*shader)(SRPFragmentShaderIn* in, SRPFragmentShaderOut* out);
	bool mayOverwriteDepth;          /* false => early depth test */
} SRPFragmentShader;

typedef struct SRPShaderProgram
{
	SRPUniform* uniform;
	SRPVertexShader* vs;
	SRPFragmentShader* fs;
} SRPShaderProgram;

/* ------------------------------------------------------------------ framebuffer
 * reference: include/srp/framebuffer.h:17-39.  Three row-major planes, y down.
 * `color` is R<<24|G<<16|B<<8|A.  The pointers are host-readable; the device-side
 * planes are authoritative and are mirrored into them according to the
 * synchronisation policy (include/srp_b200.h). */
typedef struct SRPFramebuffer
{
	size_t width, height, size;
	uint32_t* color;
	float* depth;
	uint8_t* stencil;
} SRPFramebuffer;

SRPFramebuffer* srpNewFramebuffer(size_t width, size_t height);
void srpFreeFramebuffer(SRPFramebuffer* fb);
void srpFramebufferClear(const SRPFramebuffer* fb);   /* colour 0, depth -1, stencil untouched */

/* ------------------------------------------------------------------ textures
 * reference: include/srp/texture.h:18-67 */
typedef enum { TW_REPEAT, TW_CLAMP_TO_EDGE } SRPTextureWrappingMode;
typedef enum SRPTextureParameter
{
	SRP_TEXTURE_WRAPPING_MODE_X, SRP_TEXTURE_WRAPPING_MODE_Y
} SRPTextureParameter;
typedef struct SRPTexture SRPTexture;

SRPTexture* srpNewTexture(const char* image, SRPTextureWrappingMode wrappingModeX,
                          SRPTextureWrappingMode wrappingModeY);
void srpFreeTexture(SRPTexture* texture);
SRP_SHADER_CALLABLE void srpTextureGetFilteredColor(const SRPTexture* texture, float u, float v, float out[4]);
int srpTextureGet(SRPTexture* texture, SRPTextureParameter parameter);
void srpTextureSet(SRPTexture* texture, SRPTextureParameter parameter, int data);

/* ------------------------------------------------------------------ buffers & draws
 * reference: include/srp/buffer.h:18-92 */
typedef enum SRPPrimitive
{
	SRP_PRIM_POINTS,
	SRP_PRIM_LINES, SRP_PRIM_LINE_STRIP, SRP_PRIM_LINE_LOOP,
	SRP_PRIM_TRIANGLES, SRP_PRIM_TRIANGLE_STRIP, SRP_PRIM_TRIANGLE_FAN
} SRPPrimitive;
typedef struct SRPVertexBuffer SRPVertexBuffer;
typedef struct SRPIndexBuffer SRPIndexBuffer;

SRPVertexBuffer* srpNewVertexBuffer(void);
void srpFreeVertexBuffer(SRPVertexBuffer* vb);
void srpVertexBufferCopyData(SRPVertexBuffer* vb, size_t nBytesPerVertex, size_t nBytesData, const void* data);
void srpDrawVertexBuffer(const SRPVertexBuffer* vb, const SRPFramebuffer* fb, const SRPShaderProgram* sp,
                         SRPPrimitive primitive, size_t startIndex, size_t count);

SRPIndexBuffer* srpNewIndexBuffer(void);
void srpFreeIndexBuffer(SRPIndexBuffer* ib);
void srpIndexBufferCopyData(SRPIndexBuffer* ib, SRPType indicesType, size_t nBytesData, const void* data);
void srpDrawIndexBuffer(const SRPIndexBuffer* ib, const SRPVertexBuffer* vb, const SRPFramebuffer* fb,
                        const SRPShaderProgram* sp, SRPPrimitive primitive, size_t startIndex, size_t count);

#ifdef __cplusplus
}
#endif
#endif /* SRP_API_H_ */
