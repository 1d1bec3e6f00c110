/* srp-b200 host layer -- the I/O neighbours of the draw path: what a
 * program does right before it fills its buffers and right after it has its pixels.
 *
 *   srpB200LoadOBJ            counterpart of the reference's examples/utility/objparser.c:8-79
 *                             (`loadOBJMesh`): every corner of an `f a/b/c a/b/c a/b/c` face becomes
 *                             its own vertex {vec3 position, vec2 uv, vec3 normal} = 32 bytes (the
 *                             reference's OBJVertex, objparser.h:6-10), indices are 0..n-1 (u32).
 *                             Same output bit for bit (numbers go through strtof, which is what the
 *                             reference's sscanf("%f") uses), without its fixed 65536-element tables.
 *   srpB200WritePNG /         counterpart of tests/utils/save.c:4-33 (`saveFramebufferToImage`):
 *   srpB200SaveFramebufferPNG the colour plane (R in the top byte) as an 8-bit RGBA PNG with alpha 255.
 *
 * Both are plain host code; neither touches the device except that saving a framebuffer first
 * brings its host mirror up to date. */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include "srp_internal.h"

/* ---- OBJ ------------------------------------------------------------------------------ */
typedef struct FloatTable { float* v; size_t n, cap; } FloatTable;

static bool pushFloats(FloatTable* t, const float* src, size_t k)
{
	if (t->n + k > t->cap)
	{
		size_t cap = t->cap ? t->cap * 2 : 4096;
		while (cap < t->n + k) cap *= 2;
		float* v = realloc(t->v, cap * sizeof *v);
		if (!v) return false;
		t->v = v; t->cap = cap;
	}
	memcpy(t->v + t->n, src, k * sizeof *src);
	t->n += k;
	return true;
}

/* up to `want` numbers after the keyword; missing ones stay 0 (sscanf leaves them unset in the
 * reference, i.e. undefined -- files that matter have them all) */
static void parseFloats(const char* p, float* out, int want)
{
	for (int i = 0; i < want; i++)
	{
		char* end;
		out[i] = strtof(p, &end);
		if (end == p) break;
		p = end;
	}
}

int srpB200LoadOBJ(const char* path, SRPB200Mesh* mesh)
{
	if (!mesh) return 1;
	memset(mesh, 0, sizeof *mesh);
	FILE* f = path ? fopen(path, "rb") : NULL;
	if (!f)
	{
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__, "can't open `%s`", path ? path : "(null)");
		return 1;
	}
	fseek(f, 0, SEEK_END);
	const long size = ftell(f);
	fseek(f, 0, SEEK_SET);
	char* text = malloc((size_t) (size > 0 ? size : 0) + 1);
	if (!text || (size > 0 && fread(text, 1, (size_t) size, f) != (size_t) size))
	{
		fclose(f); free(text);
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__, "short read of `%s`", path);
		return 1;
	}
	fclose(f);
	text[size > 0 ? size : 0] = 0;

	FloatTable pos = { 0 }, uv = { 0 }, nrm = { 0 }, out = { 0 };
	size_t badFaces = 0;
	bool ok = true;
	for (char* line = text; ok && *line; )
	{
		char* eol = strchr(line, '\n');
		if (eol) *eol = 0;
		float v[3] = { 0, 0, 0 };
		if (line[0] == 'v' && line[1] == ' ')
		{
			parseFloats(line + 2, v, 3);
			ok = pushFloats(&pos, v, 3);
		}
		else if (line[0] == 'v' && line[1] == 't')
		{
			parseFloats(line + 2, v, 2);
			ok = pushFloats(&uv, v, 2);
		}
		else if (line[0] == 'v' && line[1] == 'n')
		{
			parseFloats(line + 2, v, 3);
			ok = pushFloats(&nrm, v, 3);
		}
		else if (line[0] == 'f')
		{
			/* exactly the reference's accepted form: three corners `p/t/n` */
			long idx[9];
			int got = 0;
			const char* p = line + 1;
			for (; got < 9; got++)
			{
				char* end;
				idx[got] = strtol(p, &end, 10);
				if (end == p) break;
				p = end;
				if (got % 3 != 2)
				{
					if (*p != '/') { got++; break; }
					p++;
				}
			}
			bool valid = got == 9;
			for (int c = 0; valid && c < 3; c++)
				valid = idx[3 * c] >= 1 && (size_t) idx[3 * c] <= pos.n / 3
				     && idx[3 * c + 1] >= 1 && (size_t) idx[3 * c + 1] <= uv.n / 2
				     && idx[3 * c + 2] >= 1 && (size_t) idx[3 * c + 2] <= nrm.n / 3;
			if (!valid)
				badFaces++;
			else
				for (int c = 0; ok && c < 3; c++)
				{
					float vert[8];
					memcpy(vert + 0, pos.v + 3 * (idx[3 * c] - 1), 3 * sizeof(float));
					memcpy(vert + 3, uv.v + 2 * (idx[3 * c + 1] - 1), 2 * sizeof(float));
					memcpy(vert + 5, nrm.v + 3 * (idx[3 * c + 2] - 1), 3 * sizeof(float));
					ok = pushFloats(&out, vert, 8);
				}
		}
		if (!eol) break;
		line = eol + 1;
	}
	free(text); free(pos.v); free(uv.v); free(nrm.v);
	const size_t nVerts = out.n / 8;
	uint32_t* indices = ok ? malloc((nVerts ? nVerts : 1) * sizeof *indices) : NULL;
	if (!ok || !indices || nVerts > 0xFFFFFFFEull)
	{
		free(out.v); free(indices);
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__, "out of memory while reading `%s`", path);
		return 1;
	}
	for (size_t i = 0; i < nVerts; i++)
		indices[i] = (uint32_t) i;
	if (badFaces)
		srpMessage(SRP_MESSAGE_WARNING, SRP_MESSAGE_SEVERITY_LOW, __func__,
			"%zu face(s) of `%s` are not of the form `f p/t/n p/t/n p/t/n` (or index out of range) and were skipped", badFaces, path);
	mesh->vertices = out.v;
	mesh->vertexCount = nVerts;
	mesh->bytesPerVertex = 8 * sizeof(float);
	mesh->indices = indices;
	mesh->indexCount = nVerts;
	return 0;
}

void srpB200FreeMesh(SRPB200Mesh* mesh)
{
	if (!mesh) return;
	free(mesh->vertices);
	free(mesh->indices);
	memset(mesh, 0, sizeof *mesh);
}

/* ---- PNG ------------------------------------------------------------------------------ */
static void putBe32(uint8_t* p, uint32_t v) { p[0] = (uint8_t) (v >> 24); p[1] = (uint8_t) (v >> 16); p[2] = (uint8_t) (v >> 8); p[3] = (uint8_t) v; }

static bool writeChunk(FILE* f, const char type[4], const uint8_t* data, uint32_t len)
{
	uint8_t head[8], tail[4];
	putBe32(head, len);
	memcpy(head + 4, type, 4);
	uLong crc = crc32(0L, head + 4, 4);
	if (len) crc = crc32(crc, data, len);
	putBe32(tail, (uint32_t) crc);
	return fwrite(head, 1, 8, f) == 8 && (len == 0 || fwrite(data, 1, len, f) == len) && fwrite(tail, 1, 4, f) == 4;
}

int srpB200WritePNG(const char* path, size_t width, size_t height, const uint32_t* color)
{
	if (!path || !color || width == 0 || height == 0 || width > 0x7FFFFFFF / 4 || height > 0x7FFFFFFF)
		return 1;
	/* scanlines: filter byte 0 + RGBA, alpha forced to 255 like the reference's writer */
	const size_t rowBytes = 1 + width * 4;
	uint8_t* raw = malloc(rowBytes * height);
	if (!raw) return 1;
	for (size_t y = 0; y < height; y++)
	{
		uint8_t* row = raw + y * rowBytes;
		row[0] = 0;
		for (size_t x = 0; x < width; x++)
		{
			const uint32_t c = color[y * width + x];
			row[1 + 4 * x + 0] = (uint8_t) (c >> 24);
			row[1 + 4 * x + 1] = (uint8_t) (c >> 16);
			row[1 + 4 * x + 2] = (uint8_t) (c >> 8);
			row[1 + 4 * x + 3] = 0xFF;
		}
	}
	uLongf zlen = compressBound((uLong) (rowBytes * height));
	uint8_t* z = malloc(zlen);
	int rc = 1;
	if (z && compress2(z, &zlen, raw, (uLong) (rowBytes * height), 6) == Z_OK && zlen <= 0x7FFFFFFFul)
	{
		FILE* f = fopen(path, "wb");
		if (f)
		{
			static const uint8_t signature[8] = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
			uint8_t ihdr[13];
			putBe32(ihdr, (uint32_t) width);
			putBe32(ihdr + 4, (uint32_t) height);
			ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;   /* 8-bit RGBA, no interlace */
			const bool ok = fwrite(signature, 1, 8, f) == 8 && writeChunk(f, "IHDR", ihdr, 13)
				&& writeChunk(f, "IDAT", z, (uint32_t) zlen) && writeChunk(f, "IEND", NULL, 0);
			rc = (fclose(f) == 0 && ok) ? 0 : 1;
		}
	}
	free(z); free(raw);
	return rc;
}

int srpB200SaveFramebufferPNG(const SRPFramebuffer* pub, const char* path)
{
	SRPFramebufferImpl* fb = srpFramebufferImpl(pub);
	if (!fb)
		return 1;
	if (fb->mirrorStale || fb->clearPending || fb->downloadInFlight)
		srpB200FramebufferDownload(pub);
	const int rc = srpB200WritePNG(path, fb->pub.width, fb->pub.height, fb->pub.color);
	if (rc)
		srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__, "can't write `%s`", path ? path : "(null)");
	return rc;
}
