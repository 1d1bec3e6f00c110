/* TEST INFRASTRUCTURE -- see oracle_alloc.h */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define GUARD ((size_t) 1 << 16)   /* 64 KiB either side: > one 16384-px row of u32 */

void* oracleMalloc(size_t size)
{
	uint8_t* base = calloc(1, size + 2 * GUARD);
	if (!base) abort();
	memcpy(base, &size, sizeof size);
	return base + GUARD;
}
void oracleFree(void* p)
{
	if (p) free((uint8_t*) p - GUARD);
}
void* oracleRealloc(void* p, size_t size)
{
	if (!p) return oracleMalloc(size);
	size_t old;
	memcpy(&old, (uint8_t*) p - GUARD, sizeof old);
	void* q = oracleMalloc(size);
	memcpy(q, p, old < size ? old : size);
	oracleFree(p);
	return q;
}
