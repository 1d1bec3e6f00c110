/* TEST INFRASTRUCTURE (oracle side) -- the few symbols the reference library lacks for
 * being driven through the same C ABI as libsrp_b200.so (srp_b200/host.py loads either).
 * Compiled into oracle/_ref/libref_host.so together with the UNMODIFIED reference objects
 * and the host build of the built-in shader programs (srp_b200/csrc/programs). */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "srp/srp.h"
#include "oracle_alloc.h"
#include "core/texture_p.h"   /* the reference's private texture layout, src/core/texture_p.h:15-24 */

/* texture from RGB8 texels in memory: what stbi_load() would have produced for a file
 * (reference src/core/texture.c:27-48 fills exactly these fields) */
SRPTexture* srpB200NewTextureFromMemory(const uint8_t* rgb, int width, int height,
                                        SRPTextureWrappingMode wrappingModeX, SRPTextureWrappingMode wrappingModeY)
{
	SRPTexture* t = oracleMalloc(sizeof *t);   /* released by the reference's SRP_FREE */
	const size_t n = (size_t) width * height * 3;
	t->data = malloc(n);                       /* released by stbi_image_free() = free() */
	memcpy(t->data, rgb, n);
	t->width = width;
	t->height = height;
	t->widthMinusOne = width - 1;
	t->heightMinusOne = height - 1;
	t->wrappingModeX = wrappingModeX;
	t->wrappingModeY = wrappingModeY;
	return t;
}
