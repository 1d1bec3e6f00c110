/* srp-b200 host layer -- texture objects.
 * API of reference src/core/texture.c:27-137.  The object and its RGB8 texels live in
 * CUDA managed memory (read-mostly, prefetched to the GPU), because user uniforms carry
 * `SRPTexture*` values that device shaders dereference (examples/03_textured_cube.c:26).
 * The texel fetch itself, srpTextureGetFilteredColor, is defined once for host and
 * device in ../device/texture.cu.  Image decoding: PNG only (srp_png.c); the reference
 * decodes through the vendored stb_image, which is out of scope. */
#include <stdlib.h>
#include <string.h>
#include "srp_internal.h"
#include "srp/detail/texture_layout.h"

static void warnUnknownWrap(const char* func, int data)
{
	if (data != TW_REPEAT && data != TW_CLAMP_TO_EDGE)
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, func,
			"Unknown texture wrapping mode (%i). Falling back to TW_REPEAT", data);
}

SRPTexture* srpB200NewTextureFromMemory(const uint8_t* rgb, int width, int height,
                                        SRPTextureWrappingMode wrappingModeX, SRPTextureWrappingMode wrappingModeY)
{
	if (!rgb || width <= 0 || height <= 0)
		return NULL;
	const size_t nBytes = (size_t) width * height * 3;
	SRPTexture* t = srpcuMallocManaged(sizeof *t);
	uint8_t* texels = t ? srpcuMallocManaged(nBytes) : NULL;
	if (!t || !texels)
	{
		srpFatalMessage(__func__, "%s", srpcuLastError());
		srpcuFreeManaged(t);
		return NULL;
	}
	memcpy(texels, rgb, nBytes);
	t->data = texels;
	t->width = width;
	t->height = height;
	t->widthMinusOne = width - 1;
	t->heightMinusOne = height - 1;
	t->wrappingModeX = wrappingModeX;
	t->wrappingModeY = wrappingModeY;
	srpcuPrefetchToDevice(texels, nBytes);
	return t;
}

SRPTexture* srpNewTexture(const char* image, SRPTextureWrappingMode wrappingModeX, SRPTextureWrappingMode wrappingModeY)
{
	int w = 0, h = 0;
	const char* reason = "unknown";
	uint8_t* rgb = srpLoadPngRgb(image, &w, &h, &reason);
	if (rgb == NULL)
	{
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__, "Failed to load image `%s`: %s", image, reason);
		return NULL;
	}
	SRPTexture* t = srpB200NewTextureFromMemory(rgb, w, h, wrappingModeX, wrappingModeY);
	free(rgb);
	return t;
}

void srpFreeTexture(SRPTexture* t)
{
	if (!t) return;
	srpcuFreeManaged(t->data);
	srpcuFreeManaged(t);
}

int srpTextureGet(SRPTexture* t, SRPTextureParameter parameter)
{
	switch (parameter)
	{
		case SRP_TEXTURE_WRAPPING_MODE_X: return t->wrappingModeX;
		case SRP_TEXTURE_WRAPPING_MODE_Y: return t->wrappingModeY;
		default:
			srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__, "Unknown texture parameter (%i)", parameter);
			return -1;
	}
}

void srpTextureSet(SRPTexture* t, SRPTextureParameter parameter, int data)
{
	/* the object is shared with in-flight kernels: settle them before mutating it */
	srpcuSynchronize();
	switch (parameter)
	{
		case SRP_TEXTURE_WRAPPING_MODE_X:
			t->wrappingModeX = (SRPTextureWrappingMode) data;
			warnUnknownWrap(__func__, data);
			return;
		case SRP_TEXTURE_WRAPPING_MODE_Y:
			t->wrappingModeY = (SRPTextureWrappingMode) data;
			warnUnknownWrap(__func__, data);
			return;
		default:
			srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__, "Unknown texture parameter (%i)", parameter);
	}
}
