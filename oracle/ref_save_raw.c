/* TEST INFRASTRUCTURE (oracle side) -- not part of the product.
 *
 * Replacement for the reference's tests/utils/save.c (which writes a PNG through
 * stb_image_write, tests/utils/save.c:4-34).  The scene programs under
 * /root/reference/tests/scenes call `saveFramebufferToImage(fb, path)` exactly
 * once at the end of main(); here the three framebuffer planes are written
 * verbatim instead, so that colour (u32), depth (f32 bit pattern) and stencil
 * (u8) can be compared bit-exactly (the PNG +-1 gate cannot
 * see a wrong lambda chain, only the raw depth plane can).
 *
 * File layout ("SRPRAW1\0", u64 width, u64 height, then color[w*h] u32,
 * depth[w*h] f32, stencil[w*h] u8), little endian.
 *
 * Return value: the reference's scene mains do `return ok ? 0 : 1`, and
 * stbi_write_png returns non-zero on success, so 1 means success here too. */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <srp/srp.h>

int saveFramebufferToImage(const SRPFramebuffer* fb, const char* outputPath)
{
	FILE* f = fopen(outputPath, "wb");
	if (!f)
		return 0;
	const char magic[8] = "SRPRAW1";
	uint64_t dims[2] = { fb->width, fb->height };
	size_t n = fb->width * fb->height;
	int ok = fwrite(magic, 1, 8, f) == 8
		&& fwrite(dims, sizeof(uint64_t), 2, f) == 2
		&& fwrite(fb->color, sizeof(uint32_t), n, f) == n
		&& fwrite(fb->depth, sizeof(float), n, f) == n
		&& fwrite(fb->stencil, sizeof(uint8_t), n, f) == n;
	fclose(f);
	return ok;
}
