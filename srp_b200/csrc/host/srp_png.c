/* srp-b200 host layer -- minimal PNG reader for srpNewTexture().
 * Host-side file decode is not part of the draw path (the
 * reference vendors stb_image for this); only what textures need is implemented:
 * non-interlaced PNG, bit depth 8 or 16 (grey, RGB, grey+alpha, RGBA) or 1..8 (palette),
 * converted to tightly packed RGB8 -- alpha is dropped and 16-bit samples keep their
 * high byte, which is what the reference's `stbi_load(..., 3)` yields.  Inflate comes
 * from zlib. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <zlib.h>

static uint32_t be32(const uint8_t* p) { return ((uint32_t) p[0] << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c)
{
	int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
	return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

uint8_t* srpLoadPngRgb(const char* path, int* width, int* height, const char** reason)
{
	static const uint8_t signature[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
	uint8_t *file = NULL, *idat = NULL, *raw = NULL, *rgb = NULL;
	uint8_t palette[256][3];
	memset(palette, 0, sizeof palette);
	*reason = "can't fopen";
	FILE* f = fopen(path, "rb");
	if (!f) return NULL;
	fseek(f, 0, SEEK_END);
	long fileSize = ftell(f);
	fseek(f, 0, SEEK_SET);
	file = malloc(fileSize > 0 ? (size_t) fileSize : 1);
	if (!file || fileSize < 8 || fread(file, 1, (size_t) fileSize, f) != (size_t) fileSize)
	{
		fclose(f); free(file);
		*reason = "short read";
		return NULL;
	}
	fclose(f);
	*reason = "not a PNG (only PNG is supported by this build)";
	if (memcmp(file, signature, 8) != 0) { free(file); return NULL; }

	uint32_t w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
	size_t idatLen = 0;
	idat = malloc((size_t) fileSize);
	for (size_t at = 8; at + 12 <= (size_t) fileSize; )
	{
		const uint32_t len = be32(file + at);
		const uint8_t* type = file + at + 4;
		const uint8_t* data = file + at + 8;
		if (at + 12 + len > (size_t) fileSize) break;
		if (!memcmp(type, "IHDR", 4) && len >= 13)
		{
			w = be32(data); h = be32(data + 4);
			depth = data[8]; ctype = data[9]; interlace = data[12];
		}
		else if (!memcmp(type, "PLTE", 4))
			for (uint32_t i = 0; i < len / 3 && i < 256; i++)
				memcpy(palette[i], data + 3 * i, 3);
		else if (!memcmp(type, "IDAT", 4))
		{
			memcpy(idat + idatLen, data, len);
			idatLen += len;
		}
		else if (!memcmp(type, "IEND", 4))
			break;
		at += 12 + len;
	}
	int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	*reason = "unsupported PNG variant";
	if (w == 0 || h == 0 || w > 65536 || h > 65536 || channels == 0 || interlace != 0
	    || (ctype == 3 ? (depth != 1 && depth != 2 && depth != 4 && depth != 8) : (depth != 8 && depth != 16)))
		goto fail;

	const size_t bitsPerPixel = (size_t) channels * depth;
	const size_t bpp = (bitsPerPixel + 7) / 8;                /* filter distance */
	const size_t stride = (w * bitsPerPixel + 7) / 8;
	uLongf rawLen = (uLongf) ((stride + 1) * h);
	raw = malloc(rawLen);
	*reason = "corrupt PNG (inflate)";
	if (!raw || uncompress(raw, &rawLen, idat, (uLong) idatLen) != Z_OK || rawLen != (stride + 1) * h)
		goto fail;

	/* undo the per-scanline filters in place (rows keep their leading filter byte) */
	for (uint32_t y = 0; y < h; y++)
	{
		uint8_t* row = raw + (size_t) y * (stride + 1) + 1;
		const uint8_t* up = y ? row - (stride + 1) : NULL;
		const int filter = row[-1];
		for (size_t i = 0; i < stride; i++)
		{
			const int a = i >= bpp ? row[i - bpp] : 0;
			const int b = up ? up[i] : 0;
			const int c = (up && i >= bpp) ? up[i - bpp] : 0;
			int v = row[i];
			switch (filter)
			{
				case 0: break;
				case 1: v += a; break;
				case 2: v += b; break;
				case 3: v += (a + b) / 2; break;
				case 4: v += paeth(a, b, c); break;
				default: *reason = "corrupt PNG (filter)"; goto fail;
			}
			row[i] = (uint8_t) v;
		}
	}

	rgb = malloc((size_t) w * h * 3);
	if (!rgb) goto fail;
	for (uint32_t y = 0; y < h; y++)
	{
		const uint8_t* row = raw + (size_t) y * (stride + 1) + 1;
		uint8_t* out = rgb + (size_t) y * w * 3;
		for (uint32_t x = 0; x < w; x++)
		{
			if (ctype == 3)
			{
				const size_t bit = (size_t) x * depth;
				const int idx = (row[bit / 8] >> (8 - depth - (bit % 8))) & ((1 << depth) - 1);
				memcpy(out + 3 * x, palette[idx], 3);
				continue;
			}
			const size_t sampleBytes = depth / 8;
			const uint8_t* px = row + (size_t) x * channels * sampleBytes;
			if (channels <= 2)
				out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = px[0];
			else
			{
				out[3 * x + 0] = px[0];
				out[3 * x + 1] = px[sampleBytes];
				out[3 * x + 2] = px[2 * sampleBytes];
			}
		}
	}
	free(file); free(idat); free(raw);
	*width = (int) w; *height = (int) h;
	return rgb;
fail:
	free(file); free(idat); free(raw); free(rgb);
	return NULL;
}
