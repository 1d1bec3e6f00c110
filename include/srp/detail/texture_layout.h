/* srp-b200 -- layout of the (publicly opaque) SRPTexture object.  Host C code and the
 * device twin of srpTextureGetFilteredColor share it; the object and its texels live
 * in CUDA managed memory so that an `SRPTexture*` stored inside a user uniform is
 * dereferenceable from both sides (reference: src/core/texture_p.h:15-24). */
#pragma once
#include <stdint.h>
#include "srp/api.h"

struct SRPTexture
{
	uint8_t* data;          /* RGB8, row-major, top row first */
	int width, height;
	int widthMinusOne, heightMinusOne;
	SRPTextureWrappingMode wrappingModeX, wrappingModeY;
};
