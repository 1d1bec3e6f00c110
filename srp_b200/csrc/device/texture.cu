/* srp-b200 -- nearest-texel fetch callable from host programs and from device shaders.
 * Follows reference src/core/texture.c:56-81 operation by operation:
 * wrap only when the coordinate is outside [0,1] (REPEAT: u - floor(u) in double; CLAMP:
 * fmax(0, fmin(1, u)) in double), x = (W-1)*u and y = (H-1)*(1-v) in float, round with
 * (size_t)(x + 0.5) in double, channels = byte * (float)(1/255.), alpha = 1. */
#include <math.h>
#include "srp/detail/texture_layout.h"
#include "srp/detail/fpops.h"

extern "C" __host__ __device__ void srpTextureGetFilteredColor(const SRPTexture* t, float u, float v, float out[4])
{
	if (u < 0 || u > 1)
		u = (t->wrappingModeX == TW_REPEAT) ? (float) SRP_DSUB((double) u, floor((double) u))
		                                    : (float) fmax(0.0, fmin(1.0, (double) u));
	if (v < 0 || v > 1)
		v = (t->wrappingModeY == TW_REPEAT) ? (float) SRP_DSUB((double) v, floor((double) v))
		                                    : (float) fmax(0.0, fmin(1.0, (double) v));
	const float x = SRP_FMUL((float) t->widthMinusOne, u);
	const float y = SRP_FMUL((float) t->heightMinusOne, SRP_FSUB(1.0f, v));
	const size_t xi = (size_t) SRP_DADD((double) x, 0.5);
	const size_t yi = (size_t) SRP_DADD((double) y, 0.5);
	const uint8_t* texel = t->data + (xi + yi * (size_t) t->width) * 3;
	const float inv255 = (float) (1. / 255.);
	out[0] = SRP_FMUL((float) texel[0], inv255);
	out[1] = SRP_FMUL((float) texel[1], inv255);
	out[2] = SRP_FMUL((float) texel[2], inv255);
	out[3] = 1.0f;
}
