/* TEST INFRASTRUCTURE (oracle side) -- allocator handed to the reference through its own
 * customisation point (SRP_MALLOC / SRP_REALLOC / SRP_FREE, reference src/utils/defines.h:
 * 11-21).  The reference's line rasteriser writes fragments whose pixel lies one column /
 * row outside the framebuffer (no bounds check, src/raster/line.c:58-71 with the asserts of
 * fragment.c:68-69 compiled out) and never initialises the stencil
 * plane (App. B-2).  Both are harmless in the reference's own test runs only by luck of the
 * heap layout.  The oracle build therefore gives every allocation a zero-filled guard band
 * on both sides, so that those writes land in padding and the planes that ARE compared stay
 * well defined.  No arithmetic of the reference is touched. */
#ifndef ORACLE_ALLOC_H_
#define ORACLE_ALLOC_H_
#include <stddef.h>
void* oracleMalloc(size_t size);
void* oracleRealloc(void* p, size_t size);
void oracleFree(void* p);
#define SRP_MALLOC(s) oracleMalloc(s)
#define SRP_REALLOC(p, s) oracleRealloc(p, s)
#define SRP_FREE(p) oracleFree(p)
#endif
